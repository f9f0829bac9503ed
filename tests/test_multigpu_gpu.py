"""Data-parallel step on real GPUs (needs >= 2 devices, skipped otherwise): world_size 2 over NCCL.

  * GraphedTrainStep(process_group=...): the averaged gradients Adam consumes (GradSync bucket views) equal the mean of the
    per-shard gradients computed on ONE GPU, i.e. DistributedDataParallel's arithmetic (run_gun.py:63-72); tolerance is the
    bf16 rounding of the buckets (2^-9 per value, twice) - exact-ish with DLSG_GRAD_REDUCE=fp32;
  * after several replays every rank holds bit-identical weights (same reduced gradients, same Adam);
  * GanIteration(process_group=...) (G + D losses, BASELINE config 5) runs captured and keeps G and D weights identical.
"""
import contextlib
import io
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, mode, q):
    try:
        for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), NCCL_GRAPH_REGISTER='0')
        if mode == 'fp32':
            os.environ['DLSG_GRAD_REDUCE'] = 'fp32'
        import datetime
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dev = torch.device('cuda', rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=120))
        from dlsg import synth, linalg as la, losses
        from dlsg.graphs import GraphedTrainStep
        import models.model as M
        la.set_precision('bf16')
        args, V, B = synth.msr_args(), 10547, 4          # full widths: the fused / hoisted kernel paths of the benchmarked step
        torch.manual_seed(5)
        with contextlib.redirect_stdout(io.StringIO()):
            net = M.CapGnnModel(args, synth.Vocab(V))
        synth.fill_state_dict(net)
        net = net.to(dev).eval()                                  # eval(): no dropout, so one GPU can reproduce the shards
        sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
        shards = [synth.make_inputs(B, args, V, seed=300 + r) for r in range(world)]
        fr, rg, cp, lens = shards[rank]
        res = {}
        if mode in ('bf16', 'fp32'):
            # reference on this GPU alone: mean over ranks of the per-shard gradients (eager path, same kernels)
            def shard_mean_grads():
                # (in a function: no reference to `out` / the loss may survive, or their AccumulateGrad nodes - created on the
                # default stream - would be synchronised with the capture stream and invalidate the capture)
                acc = {}
                for r in range(world):
                    net.zero_grad(set_to_none=True)
                    f, g, c, l = shards[r]
                    out = net(f.to(dev), g.to(dev), c.to(dev), args.max_words, 1.0)[0]
                    losses.packed_cross_entropy(out, c.to(dev), l).backward()
                    for k, p in net.named_parameters():
                        if p.grad is not None:
                            acc[k] = acc.get(k, 0) + p.grad.detach().clone() / world
                return acc
            ref = shard_mean_grads()
            torch.cuda.synchronize()
            net.zero_grad(set_to_none=True)
            opt = torch.optim.Adam(net.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=True)
            gs = GraphedTrainStep(net, opt, fr.to(dev), rg.to(dev), cp.to(dev), lens, args.max_words, 1.0,
                                  process_group=dist.group.WORLD, warmup=1)
            net.load_state_dict(sd0)                              # undo the warm-up / capture updates: replay on the initial weights
            gs.refresh_weights()
            gs()
            torch.cuda.synchronize()
            # Norm floor: the K / Q weights of the second attention head have gradients that are a cancellation residue
            # (|g| ~ 7e-5 against 0.1 .. 6 for every other tensor; the softmax Jacobian terms sum to zero over the nodes), so two
            # EAGER runs of the same step already differ by 14 % there (fp32 atomics order; tools/diag_graph_grads.py).  Errors
            # are therefore measured against max(|ref_k|, 1e-3 * largest gradient norm of the model).
            floor = 1e-3 * max(float(v.norm()) for v in ref.values())
            errs = []
            for k, p in net.named_parameters():
                if k not in ref:
                    continue
                red = gs.sync.grad_of(p)
                assert red is not None, k
                errs.append((float((red.float() - ref[k]).norm()) / max(float(ref[k].norm()), floor), k, float(ref[k].norm())))
            errs.sort(reverse=True)
            res['grad_rel'] = errs[0][0]
            res['worst'] = errs[:6]
            res['bytes'] = gs.sync.bytes
            for _ in range(3):
                gs()
            torch.cuda.synchronize()
            flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
            other = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(other, flat)
            res['weights_equal'] = all(torch.equal(other[0], o) for o in other)
            res['moved'] = float((flat - torch.cat([sd0[k].reshape(-1) for k, _ in net.named_parameters()])).abs().max())
        else:
            from dlsg.gan import GanIteration
            with contextlib.redirect_stdout(io.StringIO()):
                D = M.DiscV2(args, V).to(dev).train()
            for p in D.parameters():                              # same critic on every rank
                dist.broadcast(p.data, 0)
            og = torch.optim.Adam(net.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=True)
            od = torch.optim.Adam(D.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=True)
            gi = GanIteration(net, D, og, od, fr.to(dev), rg.to(dev), cp.to(dev), lens, args.max_words, 1.0, 2, 0.01,
                              process_group=dist.group.WORLD, graph=True, warmup=1)
            for _ in range(2):
                out = gi()
            torch.cuda.synchronize()
            res['finite'] = all(bool(torch.isfinite(o).all()) for o in out)
            for name, m in (('G', net), ('D', D)):
                flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
                other = [torch.empty_like(flat) for _ in range(world)]
                dist.all_gather(other, flat)
                res['%s_equal' % name] = all(torch.equal(other[0], o) for o in other)
        q.put((rank, res))
        dist.barrier()
        torch.cuda.synchronize()
    except Exception as e:                                         # noqa: BLE001 - reported to the parent
        import traceback
        q.put((rank, {'error': traceback.format_exc()[-4000:]}))
    finally:
        sys.stdout.flush()
        os._exit(0)        # a captured graph with NCCL nodes keeps the communicator busy at teardown (see bench.py)


def _run(mode):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    for r, res in out.items():
        assert 'error' not in res, res['error']
    return out


@pytest.mark.timeout(400)
def test_graphed_step_world2_reduced_gradients_bf16_buckets():
    for r, res in _run('bf16').items():
        # two bf16 roundings of 2^-9 (pack, sum) on top of the run-to-run noise below
        assert res['grad_rel'] < 1.5e-2, (r, res)
        assert res['weights_equal'] and res['moved'] > 0, (r, res)


@pytest.mark.timeout(400)
def test_graphed_step_world2_reduced_gradients_fp32_buckets():
    for r, res in _run('fp32').items():
        # Two EAGER runs of one step already differ by 4.8e-3 on the worst tensor (decoder.context_att.K.weight, a small
        # gradient downstream of the atomic split-K data-gradient GEMMs; tools/diag_graph_grads.py, third session of round 2:
        # captured-vs-eager 5.2e-3, eager-vs-eager 4.8e-3, eager-vs-oracle 2.5e-2), so the bound is 2.5x that noise.
        assert res['grad_rel'] < 1.2e-2, (r, res)
        assert res['weights_equal'] and res['moved'] > 0, (r, res)


@pytest.mark.timeout(400)
def test_gan_iteration_world2_generator_and_critic_stay_identical():
    for r, res in _run('gan').items():
        assert res['finite'] and res['G_equal'] and res['D_equal'], (r, res)
