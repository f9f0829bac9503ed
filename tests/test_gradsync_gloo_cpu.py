"""The graph-capturable data-parallel path (dlsg.functional.GradSync: per-block flat buckets all-reduced as each
block's backward finishes) on CPU: world_size 2, gloo, kernels emulated.  Gradients must equal the mean of the
per-shard oracle gradients, exactly like the DDP path (tests/test_ddp_gloo_cpu.py)."""
import contextlib
import io
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT, os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dlsg import synth, ops, linalg as la, functional as DF
    from oracle import dlsg_oracle as O
    from cpu_emul import CpuEmulBackend
    import models.model as M
    ops.set_backend(CpuEmulBackend())
    la.set_precision('fp32')
    torch.set_num_threads(2)
    args, V, B = synth.small_args(), 37, 2
    with contextlib.redirect_stdout(io.StringIO()):
        net = M.CapGnnModel(args, synth.Vocab(V))
    synth.fill_state_dict(net)
    net.eval()
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=200 + rank)
    params = [p for p in net.parameters()]
    local = {}
    worst16 = 0.0
    for dtype in (torch.bfloat16, torch.float32):          # the product default (bf16 buckets) first, then exact fp32 buckets
        net.zero_grad(set_to_none=True)
        out = net(frames, regions, caps, args.max_words, 1.0)[0]
        sync = DF.GradSync(None, dtype=dtype)               # the REAL GradSync: gloo + CPU tensors take its stream-less branch
        sync.begin_step()
        DF.GRAD_SYNC = sync
        try:
            O.packed_ce_loss(out, caps, lens).backward()
        finally:
            DF.GRAD_SYNC = None
        sync.wait()
        assert len(sync._plans) >= 5, 'expected the early (mid-backward) reductions on top of one per block'
        for p_ in params:                                    # p.grad keeps the LOCAL gradient; every one has an averaged bucket view
            assert (p_.grad is None) == (sync.grad_of(p_) is None)
        if dtype == torch.bfloat16:
            local = {id(p_): p_.grad.clone() for p_ in params if p_.grad is not None}
            red16 = {id(p_): sync.grad_of(p_).float().clone() for p_ in params if p_.grad is not None}
        sync.write_back(params)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    for r in range(world):
        f, g, c, l = synth.make_inputs(B, args, V, seed=200 + r)
        ro = O.cap_gnn_forward(sd, f, g, c, args.max_words, 1.0, args.a_feature_size)[0]
        (O.packed_ce_loss(ro, c, l) / world).backward()
    worst = 0.0
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        ref = sd[k].grad
        if float((p.grad - ref).abs().max()) > 1e-7:
            worst = max(worst, float((p.grad - ref).norm() / (ref.norm() + 1e-12)))
        # bf16 buckets: one rounding of each rank's gradient + one of the sum (2^-9 relative each)
        worst16 = max(worst16, float((red16[id(p)] - ref).norm() / (ref.norm() + 1e-12)))
    q.put((rank, worst, worst16))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_block_bucket_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, worst, worst16 in res:
        assert worst < 1e-4, (rank, worst)
        assert worst16 < 8e-3, (rank, worst16)
