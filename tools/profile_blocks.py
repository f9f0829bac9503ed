"""Per-block GPU time of one D-LSG training step at the bench configuration (B=64, MSR-VTT shapes, bf16).

Each block's forward and backward (TunBlock, EncoderVisualBlock, DecoderTrainBlock), the fused CE (+ its backward) and
the Adam step are captured into their own CUDA graph and replayed, so the numbers are device times without host launch
cost - the same regime as the graphed train step of bench.py.  One JSON line per segment + their sum.

  python tools/profile_blocks.py [--single-stream]
"""
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
from dlsg import synth, ops, losses, linalg as la, functional as DF, decoder as DD  # noqa: E402
import models.model as M  # noqa: E402

dev = torch.device('cuda')
la.set_precision('bf16')
if '--single-stream' in sys.argv:
    DF.two_streams = lambda like, f0, f1: (f0(), f1())
B, V = 64, 10547
args = synth.msr_args(train_batch_size=B)
torch.manual_seed(12)
with contextlib.redirect_stdout(io.StringIO()):
    net = M.CapGnnModel(args, synth.Vocab(V)).to(dev)
net.train()
opt = torch.optim.Adam(net.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=True)
frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=12)
fr, rg, cp = frames.to(dev), regions.to(dev), caps.to(dev)
be = ops.backend()

rec = []


def wrap(cls):
    of, ob = cls.forward, cls.backward

    def f(self, t):
        outs, sv = of(self, t)
        rec.append(['fwd', self, t, sv, None])
        return outs, sv

    def b(self, sv, gouts):
        rec.append(['bwd', self, None, sv, gouts])
        return ob(self, sv, gouts)
    cls.forward, cls.backward = f, b
    return of, ob


orig = {c: wrap(c) for c in (DF.TunBlock, DF.EncoderVisualBlock, DD.DecoderTrainBlock)}


def step():
    opt.zero_grad(set_to_none=True)
    out = net(fr, rg, cp, 26, 1.0)[0]
    loss = losses.packed_cross_entropy(out, cp, lens)
    loss.backward()
    opt.step()
    return out


for _ in range(2):
    rec.clear()
    out = step()
torch.cuda.synchronize()
calls = list(rec)
for c, (of, ob) in orig.items():
    c.forward, c.backward = of, ob


def time_graph(name, fn, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    DF.WC.force = True
    l0 = be.launches
    try:
        with torch.cuda.graph(g):
            keep = fn()
    finally:
        DF.WC.force = False
    n = be.launches - l0
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    if name is not None:
        print(json.dumps({'segment': name, 'ms': round(best, 3), 'dlsg_calls': n}), flush=True)
    del keep
    return best


ABLATE = ('gemm', 'convert', 'norm_fwd', 'norm_bwd', 'lstm_cell_fwd', 'lstm_cell_bwd', 'lstm_cell_norm_fwd', 'norm_lstm_cell_bwd',
          'attn2_fwd', 'attn2_bwd', 'colsum', 'embedding_gather', 'embedding_scatter_add', 'latent_psl_fwd', 'latent_psl_bwd',
          'softmax_fwd', 'softmax_bwd', 'dropout', 'axpby', 'add_rowbcast', 'mean_nodes_fwd', 'mean_nodes_bwd', 'cast')
ablate = '--ablate' in sys.argv


class CountCalls:
    def __init__(self, fn):
        self.fn, self.n = fn, 0

    def __call__(self, *a, **k):
        self.n += 1


def ablation(name, fn, base_ms):
    """Marginal cost of each backend method inside this segment: replace it by a no-op (results are garbage, the launch
    sequence has no data-dependent control flow) and re-time the segment."""
    for m in ABLATE:
        if not hasattr(be, m):
            continue
        cc = CountCalls(getattr(be, m))
        setattr(be, m, cc)
        try:
            ms = time_graph(None, fn, reps=3)
        finally:
            delattr(be, m)
        if cc.n:
            # capture ran fn twice (warm-up + capture)
            print(json.dumps({'segment': name, 'without': m, 'calls': cc.n // 2, 'ms': round(ms, 3), 'delta_us': round((base_ms - ms) * 1e3, 1),
                              'us_per_call': round((base_ms - ms) * 1e3 / max(1, cc.n // 2), 2)}), flush=True)


total = 0.0
with torch.no_grad():
    for kind, blk, t, sv, gouts in calls:
        name = '%s %s' % (type(blk).__name__, kind)
        if kind == 'fwd':
            def fn(blk=blk, t=t):
                DF.WC.begin_train_block()
                return blk.forward(t)
        else:
            core = sv.get('core') if isinstance(sv, dict) else None

            def fn(blk=blk, sv=sv, gouts=gouts, core=core):
                if core is not None:                      # re-do the transposed weight packs like a fresh step does
                    for k in [k for k in core.pk if k.endswith('T')]:
                        del core.pk[k]
                return blk.backward(sv, gouts)
        base = time_graph(name, fn)
        total += base
        if ablate:
            ablation(name, fn, base)

# loss (fused masked CE) forward + backward
logits = out.detach().clone().requires_grad_(True)
lens_d = torch.as_tensor(list(lens), dtype=torch.int32, device=dev)
inv = torch.tensor([1.0 / max(1, int(sum(lens)))], dtype=torch.float32, device=dev)


def ce():
    logits.grad = None
    loss = losses.packed_cross_entropy(logits, cp, lens_d, inv)
    loss.backward()
    return loss


with torch.enable_grad():
    total += time_graph('packed_cross_entropy fwd+bwd', ce)
total += time_graph('Adam (torch fused, capturable)', lambda: opt.step())
print(json.dumps({'segment': 'SUM', 'ms': round(total, 3)}))
