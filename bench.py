"""Benchmark of the D-LSG hot path on B200 (contract: one JSON line on rank 0).

Workload (BASELINE.json configs[1]): D-LSG training step, batch 64 per GPU, MSR-VTT-shaped synthetic features
(26 frames, 1536-d 2D + 2048-d 3D features, 36 x 2048 regions, V=10547, captions <= 26 tokens), bf16 tensor-core
GEMMs with fp32 accumulation / fp32 master weights.  One step = zero_grad -> CapGnnModel forward -> fused masked
cross-entropy -> backward -> Adam (lr 1.6e-4, betas (0.5,0.9), run_gun.py:91).  Metric: train clips/s.

  value : device-resident inputs (regions are 490 MB/step > 126 MB L2, so every step streams from HBM)
  e2e   : same step through the public module API with HOST (pinned) inputs, H2D inside the timed region and a
          D2H read of the loss every step
  --impl reference : the reference algorithm's CPU port (oracle/) timed on the host cores on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'd-lsg-video-caption_b200')
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = 'train clips/s (fwd+bwd+Adam, B=64/GPU, MSR-VTT-shaped)'
V_MSR = 10547


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='b200')
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--no-decode', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--graph', type=int, default=int(os.environ.get('DLSG_GRAPH', '1')))
    ap.add_argument('--profile-step', action='store_true', help='run W warm-up steps, then ONE eager step inside cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)')
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_port_step_time(batch, steps, warmup, threads):
    """Reference algorithm (oracle port, torch CPU fp32 autograd) fwd+CE+bwd+Adam on `batch` MSR-shaped clips."""
    from dlsg import synth
    from oracle import dlsg_oracle as O
    import contextlib
    import io
    import models.model as M
    torch.set_num_threads(threads)
    args = synth.msr_args()
    with contextlib.redirect_stdout(io.StringIO()):
        net = M.CapGnnModel(args, synth.Vocab(V_MSR))          # parameter container only (CPU); math = oracle
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and not k.endswith('pe.pe'))
          for k, v in net.state_dict().items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1.6e-4, betas=(0.5, 0.9))
    frames, regions, caps, lens = synth.make_inputs(batch, args, V_MSR, seed=12)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        out = O.cap_gnn_forward(sd, frames, regions, caps, 26, 1.0, args.a_feature_size)[0]
        loss = O.packed_ce_loss(out, caps, lens)
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_b = 2
    t = cpu_port_step_time(sample_b, a.steps, a.warmup, threads)
    v = sample_b / t
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'clips/s', 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'D-LSG training step, MSR-VTT-shaped synthetic features (bounded sample: %d clips/step)' % sample_b},
            'cpu_baseline': {'value': v, 'unit': 'clips/s', 'cores': threads, 'kind': 'port',
                             'sample': '%d clips per step, fwd+CE+bwd+Adam, oracle/dlsg_oracle.py on torch CPU fp32' % sample_b},
            'e2e': {'value': v, 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


_STAGE = ['start']
_T0 = time.time()


def stage(name):
    """Progress marker on stderr (one line per stage, rank-tagged): a stalled multi-rank run can be located from the log."""
    _STAGE[0] = name
    sys.stderr.write('[bench rank %s +%.1fs] %s\n' % (os.environ.get('RANK', '0'), time.time() - _T0, name))
    sys.stderr.flush()


def arm_watchdog(a):
    """A run that has not printed its line after DLSG_BENCH_TIMEOUT seconds (default 420) reports where it stalled and exits,
    instead of sitting in a collective until the caller's own limit kills it."""
    limit = float(os.environ.get('DLSG_BENCH_TIMEOUT', '420'))

    def fire():
        if int(os.environ.get('RANK', '0')) == 0:
            print(json.dumps({'metric': METRIC, 'value': None, 'unit': 'clips/s', 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
                              'error': 'no result after %.0f s; last stage: %s' % (limit, _STAGE[0])}), flush=True)
        sys.stderr.write('[bench rank %s] watchdog: stalled in stage %r\n' % (os.environ.get('RANK', '0'), _STAGE[0]))
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(limit, fire)
    t.daemon = True
    t.start()
    return t


# ---- multi-rank supervisor -------------------------------------------------------------------------------------------
# The only 8-GPU attempt of round 1 stalled after NCCL's start-up line and could not be investigated.  For
# world >= 4 every rank therefore runs the measurement in a child process and, if the children have not finished inside a
# fixed wall-clock window, kills them and tries the next, more conservative configuration.  The windows are absolute
# (measured from process start, which torchrun makes simultaneous on all ranks), so every rank switches at the same time.
TIERS = [('cuda-graph step, gradient all-reduce inside the graph', {}, []),
         ('cuda-graph step, gradient all-reduce inside the graph, NCCL_NVLS_ENABLE=0', {'NCCL_NVLS_ENABLE': '0'}, []),
         ('eager DistributedDataParallel step, NCCL_NVLS_ENABLE=0', {'NCCL_NVLS_ENABLE': '0'}, ['--graph', '0'])]


def supervise(a):
    win = [float(x) for x in os.environ.get('DLSG_BENCH_TIER_SECONDS', '120,120,150').split(',')]
    base_port = int(os.environ.get('MASTER_PORT', '29500'))
    rank0 = int(os.environ.get('RANK', '0')) == 0
    t_open = 0.0
    for i, (name, env_add, extra) in enumerate(TIERS):
        t_close = t_open + win[min(i, len(win) - 1)]
        delay = _T0 + t_open - time.time()
        if delay > 0:
            time.sleep(delay)                         # every rank opens tier i at the same wall-clock time
        env = dict(os.environ)
        env.update(env_add)
        env.update(DLSG_BENCH_WORKER='1', DLSG_BENCH_TIER=name, DLSG_BENCH_TIER_INDEX=str(i))
        if i > 0:                                     # fresh rendezvous: rank 0's worker hosts a new store on another port
            env['MASTER_PORT'] = str(base_port + 17 * i)
            env['TORCHELASTIC_USE_AGENT_STORE'] = 'False'
        stage('supervisor: tier %d (%s)' % (i, name))
        child = subprocess.Popen([sys.executable, os.path.abspath(__file__)] + sys.argv[1:] + extra, env=env)
        rc = None
        try:
            rc = child.wait(timeout=max(1.0, _T0 + t_close - time.time()))
        except subprocess.TimeoutExpired:
            child.kill()
            child.wait()
        if rc == 0:
            return 0
        stage('supervisor: tier %d %s' % (i, 'timed out' if rc is None else 'exited with %s' % rc))
        t_open = t_close + float(os.environ.get('DLSG_BENCH_TIER_GAP', '8'))   # let the killed workers' GPU contexts disappear
    if rank0:
        print(json.dumps({'metric': METRIC, 'value': None, 'unit': 'clips/s', 'n_gpus': a.gpus, 'steps': a.steps, 'warmup': a.warmup,
                          'error': 'no multi-rank configuration finished (see the stage log on stderr)'}), flush=True)
    return 3


def main():
    a = parse()
    if a.impl == 'reference':
        return run_reference(a)
    if int(os.environ.get('WORLD_SIZE', '1')) >= 4 and os.environ.get('DLSG_BENCH_WORKER') != '1':
        sys.exit(supervise(a))
    fake = os.environ.get('DLSG_BENCH_FAKE')          # supervisor self-test hook (tests/test_bench_supervisor_cpu.py)
    if fake is not None:
        if os.environ.get('DLSG_BENCH_TIER_INDEX', '0') in fake.split(','):
            time.sleep(3600)
        print(json.dumps({'fake': True, 'tier': os.environ.get('DLSG_BENCH_TIER'), 'port': os.environ.get('MASTER_PORT'),
                          'nvls': os.environ.get('NCCL_NVLS_ENABLE'), 'argv': sys.argv[1:]}), flush=True)
        return
    watchdog = arm_watchdog(a)
    import contextlib
    import io
    from dlsg import synth, ops, losses, linalg as la
    import models.model as M
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        # user-buffer registration for captured collectives only applies to VMM allocations (not the caching allocator's):
        # switch the attempt off; a collective that cannot complete aborts after 3 minutes instead of hanging
        os.environ.setdefault('NCCL_GRAPH_REGISTER', '0')
        stage('init_process_group(nccl) world=%d' % world)
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
        stage('process group ready')
    la.set_precision('bf16')
    args = synth.msr_args(train_batch_size=a.batch)
    B = a.batch
    torch.manual_seed(12)
    with contextlib.redirect_stdout(io.StringIO()):
        net = M.CapGnnModel(args, synth.Vocab(V_MSR)).to(dev)
    net.train()
    model = net
    use_graph = bool(a.graph) and not a.profile_step
    if world > 1 and not use_graph:
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True,
                                                          gradient_as_bucket_view=True)
    opt = torch.optim.Adam(net.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=use_graph)
    stage('model built; generating synthetic inputs')
    frames, regions, caps, lens = synth.make_inputs(B, args, V_MSR, seed=12 + rank)
    h_fr, h_rg, h_cp = frames.pin_memory(), regions.pin_memory(), caps.pin_memory()
    d_fr, d_rg, d_cp = h_fr.to(dev), h_rg.to(dev), h_cp.to(dev)
    be = ops.backend()

    def step(fr, rg, cp):
        opt.zero_grad(set_to_none=True)
        out = model(fr, rg, cp, 26, 1.0)[0]
        loss = losses.packed_cross_entropy(out, cp, lens)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n

    if a.profile_step:
        for _ in range(a.warmup):
            step(d_fr, d_rg, d_cp)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(d_fr, d_rg, d_cp)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # ---- device-resident timing
    eager_ms = None
    if use_graph:
        from dlsg.graphs import GraphedTrainStep
        if world == 1:
            for _ in range(2):
                step(d_fr, d_rg, d_cp)
            eager_ms = timed(lambda: step(d_fr, d_rg, d_cp), 3)
        l0 = be.launches
        stage('capturing the training step (eager warm-up steps first when world > 1)')
        gs = GraphedTrainStep(net, opt, d_fr, d_rg, d_cp, lens, 26, 1.0,
                              process_group=(dist.group.WORLD if dist is not None else None), warmup=(0 if world == 1 else 3))
        launches = gs.launches
        run_dev = lambda: gs()

        def e2e_step():
            gs.load(h_fr, h_rg, h_cp)
            return gs().item()
    else:
        launches = None
        run_dev = lambda: step(d_fr, d_rg, d_cp)

        def e2e_step():
            fr = h_fr.to(dev, non_blocking=True)
            rg = h_rg.to(dev, non_blocking=True)
            cp = h_cp.to(dev, non_blocking=True)
            return step(fr, rg, cp).item()
    stage('warm-up replays')
    for _ in range(a.warmup):
        run_dev()
    stage('timed region: %d steps' % a.steps)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = be.launches
    ms = timed(run_dev, a.steps)
    if launches is None:
        launches = (be.launches - l0) // a.steps
    clocks = sampler.stop() if rank == 0 else None
    # ---- end-to-end: host pinned inputs -> H2D -> step -> loss.item()
    stage('end-to-end (host inputs) timing')
    for _ in range(2):
        e2e_step()
    ms_e2e_serial = timed(e2e_step, a.steps)
    ms_e2e = ms_e2e_serial
    h2d = h_fr.numel() * 4 + h_rg.numel() * 4 + h_cp.numel() * 8
    if use_graph:
        # Same work with the loader-style prefetch any trainer uses (DataLoader(pin_memory) + non_blocking copies): the H2D
        # copy of step k+1 runs on a copy stream while step k computes; every step still copies its own 515 MB from pinned
        # host memory inside the timed region and reads its loss back.
        copy_stream = torch.cuda.Stream()
        stage_bufs = [torch.empty_like(d_fr), torch.empty_like(d_rg), torch.empty_like(d_cp)]
        ready = torch.cuda.Event()
        consumed = torch.cuda.Event()

        def prefetch():
            copy_stream.wait_event(consumed)
            with torch.cuda.stream(copy_stream):
                stage_bufs[0].copy_(h_fr, non_blocking=True)
                stage_bufs[1].copy_(h_rg, non_blocking=True)
                stage_bufs[2].copy_(h_cp, non_blocking=True)
                ready.record(copy_stream)

        def e2e_pipe():
            cur = torch.cuda.current_stream()
            cur.wait_event(ready)
            gs.load(stage_bufs[0], stage_bufs[1], stage_bufs[2])          # device-to-device into the graph's static inputs
            consumed.record(cur)
            prefetch()                                     # next step's H2D overlaps this step's compute
            return gs().item()
        consumed.record(torch.cuda.current_stream())
        prefetch()
        for _ in range(2):
            e2e_pipe()
        ms_e2e = timed(e2e_pipe, a.steps)

    # ---- dominant kernel: region-projection GEMM (both encoders fused: M=B*936, N=2048, K=2048) timed alone
    roof = None
    extra = {}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        Mr, Nr, Kr = B * 26 * 36, 2048, 2048
        A = torch.randn(Mr, Kr, device=dev).to(torch.bfloat16)
        Wt = torch.randn(Nr, Kr, device=dev).to(torch.bfloat16)
        bias = torch.randn(Nr, device=dev)
        O_ = torch.empty(Mr, Nr, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            be.gemm(A, Wt, O_, bias=bias, tanh=True)
        n = 10
        k_ms = timed(lambda: be.gemm(A, Wt, O_, bias=bias, tanh=True), n) if dist is None else None
        if k_ms is None:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                be.gemm(A, Wt, O_, bias=bias, tanh=True)
            e1.record()
            torch.cuda.synchronize()
            k_ms = e0.elapsed_time(e1) / n
        flops = 2.0 * Mr * Nr * Kr
        ach = flops / (k_ms * 1e-3) / 1e12
        peak = peaks.get('bf16_tflops', 1590.0)
        roof = {'bound': 'tensor', 'kernel': 'gemm_tc_kernel<256> region projection %dx%dx%d bf16 (+bias+tanh, bf16 out)' % (Mr, Nr, Kr),
                'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the committed ncu --set full
                # capture profiles/r01_ncu_full_top_kernels_v2.json (algorithmic: A 245 MB + W 8 MB read, 245 MB written)
                'traffic': 253.84e6 + 208.14e6, 'traffic_unit': 'bytes/launch',
                'peak_source': 'measured (MEASURED_PEAKS.json bf16_tflops, burst)' if 'bf16_tflops' in peaks else 'fallback',
                'ms_per_launch': k_ms, 'launches_per_step': 1}
        del A, Wt, O_
        # ---- decoding throughput (secondary metrics of BASELINE.json: greedy B=256, beam-5 B=128)
        stage('secondary metrics (roofline kernel, decode, GAN iteration, CPU baseline)')
        if not a.no_decode and world == 1:
            net.eval()
            with torch.no_grad():
                for name, Bd, beam in (('greedy_captions_per_s_B256', 256, 1), ('beam5_captions_per_s_B128', 128, 5)):
                    f2, r2, _, _ = synth.make_inputs(Bd, args, V_MSR, seed=7)
                    f2, r2 = f2.to(dev), r2.to(dev)
                    net.update_beam_size(beam)
                    for _ in range(3):
                        net(f2, r2, None)
                    dms = timed(lambda: net(f2, r2, None), 5)
                    extra[name + '_eager'] = Bd / (dms * 1e-3)
                    if use_graph:
                        from dlsg.graphs import GraphedDecode
                        gd = GraphedDecode(net, f2, r2, beam)
                        for _ in range(2):
                            gd()
                        dms = timed(lambda: gd(), 5)
                        del gd
                    extra[name] = Bd / (dms * 1e-3)
            net.train()
            # ---- full GAN iteration of the live trainer (BASELINE.json configs[4] at one GPU; run_gun.py:147-234 + :339-398):
            # G forward, 5 critic steps with the WGAN-GP double backward, G step with the critic term, both Adams
            try:
                from dlsg.gan import GanIteration
                with contextlib.redirect_stdout(io.StringIO()):
                    Dnet = M.DiscV2(args, V_MSR).to(dev).train()
                og = torch.optim.Adam(net.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=True)
                od = torch.optim.Adam(Dnet.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=True)
                gi = GanIteration(net, Dnet, og, od, d_fr, d_rg, d_cp, lens, 26, 0.6, 5, 0.01, graph=use_graph)
                for _ in range(2):
                    gi()
                gms = timed(lambda: gi(), 3)
                extra['gan_iteration_ms_B%d' % B] = gms
                extra['gan_iteration_clips_per_s'] = B / (gms * 1e-3)
                del gi, Dnet, og, od
            except Exception as e:                  # secondary metric: never take the headline line down with it
                extra['gan_iteration_error'] = repr(e)[:200]
    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sb = 2
        t = cpu_port_step_time(sb, 2, 1, threads)
        cpu = {'value': sb / t, 'unit': 'clips/s', 'cores': threads, 'kind': 'port',
               'sample': '%d clips per step x 2 steps, fwd+CE+bwd+Adam, oracle/dlsg_oracle.py (torch CPU fp32)' % sb}
    stage('done')
    watchdog.cancel()
    if rank == 0:
        line = {'metric': METRIC, 'value': world * B / (ms * 1e-3), 'unit': 'clips/s', 'n_gpus': world, 'steps': a.steps,
                'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16', 'data': 'synthetic',
                'config': {'workload': 'D-LSG training step (CapGnnModel fwd + masked CE + bwd + Adam), batch %d/GPU, MSR-VTT-shaped '
                                       'synthetic features (26 frames, 1536+2048-d, 36x2048 regions, V=%d), bf16 GEMMs fp32 accum' % (B, V_MSR),
                           'global_batch': world * B, 'parallelism': 'dp%d' % world,
                           'multi_rank_path': os.environ.get('DLSG_BENCH_TIER', TIERS[0][0] if use_graph else TIERS[2][0]) if world > 1 else None,
                           'l2': 'inputs (490 MB regions/step) exceed the 126 MB L2; no explicit flush'},
                'e2e': {'value': world * B / (ms_e2e * 1e-3), 'unit': 'clips/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                        'ms_per_step': ms_e2e, 'mode': 'H2D of step k+1 prefetched on a copy stream during step k' if use_graph else 'serial',
                        'serial_ms_per_step': ms_e2e_serial},
                'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu,
                'cuda_graph': use_graph, 'eager_ms_per_step': eager_ms}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        # A captured CUDA graph that contains NCCL kernels keeps the communicator busy at interpreter shutdown
        # (destroy_process_group was observed to hang): make sure all ranks are done, then leave without teardown.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
