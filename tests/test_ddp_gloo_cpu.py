"""N>1 path on CPU: world_size-2 gloo DDP over the drop-in CapGnnModel (kernels emulated by tests/cpu_emul.py).

Checks the data-parallel contract of run_gun.py:62-72 - DistributedDataParallel(find_unused_parameters=True)
wraps the model unchanged, the 6 never-used parameters keep grad=None, and the all-reduced gradients equal the
mean of the per-shard oracle gradients (clips are independent; the gradient all-reduce is the only collective).
"""
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
    from dlsg import synth, ops, linalg as la
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
    ddp = torch.nn.parallel.DistributedDataParallel(net, find_unused_parameters=True)
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=100 + rank)
    out = ddp(frames, regions, caps, args.max_words, 1.0)[0]
    O.packed_ce_loss(out, caps, lens).backward()
    # oracle gradients of BOTH shards, averaged (what the all-reduce must produce)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    for r in range(world):
        f, g, c, l = synth.make_inputs(B, args, V, seed=100 + r)
        ro = O.cap_gnn_forward(sd, f, g, c, args.max_words, 1.0, args.a_feature_size)[0]
        (O.packed_ce_loss(ro, c, l) / world).backward()
    worst, n_none = 0.0, 0
    for k, p in net.named_parameters():
        if p.grad is None:
            n_none += 1
            continue
        ref = sd[k].grad
        err = float((p.grad - ref).abs().max())
        rel = float((p.grad - ref).norm() / (ref.norm() + 1e-12))
        if err > 1e-7:
            worst = max(worst, rel)
    q.put((rank, worst, n_none))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ddp_world2_gradients_match_oracle_mean():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, worst, n_none in res:
        assert worst < 1e-4, (rank, worst)
        assert n_none == 6, n_none
