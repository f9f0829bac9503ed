"""Caller-side glue of the hot path on CPU: batched id->caption conversion, the any-world-size merge
of per-rank caption dicts (world_size-2 gloo, kernels emulated), and the input pipeline's staging / ordering logic."""
import collections
import contextlib
import io
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup_paths():
    for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT, os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)


class _Opt:
    num_obj = 6


def _loader(args, V, vids, bs, seed):
    from dlsg import synth
    batches = []
    for i in range(0, len(vids), bs):
        ids = vids[i:i + bs]
        fr, rg, _, _ = synth.make_inputs(len(ids), args, V, seed=seed + i)
        batches.append((fr, rg, None, list(ids)))
    return batches


def _make_net():
    from dlsg import synth, ops, linalg as la
    from cpu_emul import CpuEmulBackend
    import models.model as M
    ops.set_backend(CpuEmulBackend())
    la.set_precision('fp32')
    args, V = synth.small_args(), 37
    with contextlib.redirect_stdout(io.StringIO()):
        net = M.CapGnnModel(args, synth.Vocab(V))
    synth.fill_state_dict(net)
    net.eval()
    return net, args, V


def _worker(rank, world, port, q):
    _setup_paths()
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from dlsg import evalglue
    net, args, V = _make_net()
    opt = _Opt()
    opt.num_obj = args.num_obj
    vids = list(range(100, 107))                    # 7 clips: the sampler pads the last shard with a duplicate
    shard = vids[rank::world]
    if len(shard) < (len(vids) + world - 1) // world:
        shard = shard + [vids[0]]
    loaders = {r: None for r in range(world)}
    # every clip's features depend only on its id, so a duplicate decodes to the same caption on any rank
    def loader_for(ids):
        from dlsg import synth
        out = []
        for v in ids:
            fr, rg, _, _ = synth.make_inputs(1, args, V, seed=v)
            out.append((fr, rg, None, [v]))
        return out
    merged, _ = evalglue.gather_results_all_ranks(net, opt, loader_for(shard), multi_gpu=False)
    single, _ = evalglue.gather_results(net, opt, loader_for(vids))
    ok = (dict(merged) == dict(single)) and len(merged) == len(vids)
    q.put((rank, ok, len(merged)))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_merge_of_rank_shards_equals_single_process_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res


def test_tokens_to_captions_equals_decode_tokens_rows():
    _setup_paths()
    from dlsg import evalglue, ops
    old = ops._backend
    try:
        net, args, V = _make_net()
        g = torch.Generator().manual_seed(3)
        table = torch.randint(0, V, (9, args.max_words), generator=g)
        table[0, 0] = net.decoder.vocab('<end>')          # empty caption
        table[1, :] = 5                                   # never ends
        got = evalglue.tokens_to_captions(net.decoder, table)
        want = [net.decoder.decode_tokens(row) for row in table]
        assert got == want
    finally:
        ops.set_backend(old)


class _FakeStep:
    def __init__(self):
        self.loaded = []
        self.lens = []

    def load(self, fr, rg, cp, cap_lens=None):
        self.loaded.append((fr.clone(), rg.clone(), cp.clone()))
        self.lens.append(cap_lens)

    def __call__(self):
        fr, rg, cp = self.loaded[-1]
        return fr.float().sum() + rg.float().sum() + cp.sum()


def test_feature_pipe_stages_bf16_and_keeps_batches_in_order():
    _setup_paths()
    from dlsg import pipeline
    fr = torch.randn(2, 26, 48)
    rg = torch.randn(2, 26, 6, 64)
    cp = torch.randint(0, 30, (2, 26))
    h = pipeline.stage(fr, rg, cp, pin=False)
    assert h[0].dtype == torch.bfloat16 and h[1].dtype == torch.bfloat16 and h[2].dtype == torch.int64
    assert torch.equal(h[1], rg.to(torch.bfloat16))
    step = _FakeStep()
    pipe = pipeline.FeaturePipe(step, *h, device='cpu')
    assert pipe.bytes_per_batch == fr.numel() * 2 + rg.numel() * 2 + cp.numel() * 8
    with pytest.raises(RuntimeError):
        pipe.run()                                       # nothing staged
    outs = []
    for k in range(3):
        hk = pipeline.stage(fr + k, rg - k, cp, pin=False)
        pipe.put(*hk, cap_lens=(3 + k, 4))
        with pytest.raises(RuntimeError):
            pipe.put(*hk)                                # the staged batch must be consumed first
        outs.append(float(pipe.run()))
        assert torch.equal(step.loaded[-1][0], hk[0]) and step.lens[-1] == (3 + k, 4)
    assert len(set(outs)) == 3
    with pytest.raises(ValueError):
        pipe.put(h[0][:1], h[1][:1], h[2][:1])           # wrong batch shape
