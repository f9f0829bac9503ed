"""Warm, in-graph timing of every launch of ONE decoder time step (forward) at the bench shapes (B=64, MSR-VTT, bf16),
each kernel alone (48 back-to-back launches in a CUDA graph) and the whole step as the real dependent chain."""
import contextlib
import io
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
from dlsg import synth, ops, linalg as la, functional as DF, decoder as DD  # noqa: E402
import models.model as M  # noqa: E402

dev = torch.device('cuda')
la.set_precision('bf16')
B, V, T = 64, 10547, 26
args = synth.msr_args(train_batch_size=B)
with contextlib.redirect_stdout(io.StringIO()):
    net = M.CapGnnModel(args, synth.Vocab(V)).to(dev)
be = ops.backend()
t = {k: v.detach() for k, v in net.decoder._used().items()}
DF.WC.begin_train_block()
core = DD.DecoderCore(t, '', True, 0.3)
P = 5
nodes = torch.randn(core.nh, B, P, core.H, device=dev)
Kp, Vp, glob, Gq, n_op = core.precompute(nodes)
b = core.alloc(T, B, P, nodes)
Dall = la.op_zeros((B, T), core.Hd, nodes)
pk, pf = core.pk, ''
nh, H, Hq, Hd = core.nh, core.H, core.Hq, core.Hd
oq, ol, oQ = core.oq, core.ol, core.oQ
print(json.dumps({'Hq': Hq, 'Hd': Hd, 'H': H, 'nh': nh, 'Kq': core.Kq, 'Kl': core.Kl, 'Sq': b.Sq, 'Sl': b.Sl, 'hoist': core.hoist,
                  'fused': core.fused}), flush=True)


def timeit(name, fn, reps=48):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    print(json.dumps({'case': name, 'us': round(best, 2)}), flush=True)


i, j = 3, 4
lnq_w, lnq_b = t['query_lstm_layernorm.weight'], t['query_lstm_layernorm.bias']
lnl_w, lnl_b = t['lang_lstm_layernorm.weight'], t['lang_lstm_layernorm.bias']
for drop in (None, (0.3, 1234, 0)):
    tag = 'dropout' if drop else 'no-drop'
    timeit('lstm_cell_norm_fwd query (S=%d, H=%d) %s' % (b.Sq, Hq, tag),
           lambda: be.lstm_cell_norm_fwd(b.gq[:, i], b.cq[i], b.cq[j], lnq_w, lnq_b, b.q32[i], h_out=b.qh[i], row_bias=Gq,
                                         h2=b.Xq[j][:, oQ:oQ + Hq], y2=b.Xl[i][:, oq:oq + Hq], stats=b.statq[i], ydrop=drop))
    timeit('lstm_cell_norm_fwd lang  (S=%d, H=%d) %s' % (b.Sl, Hd, tag),
           lambda: be.lstm_cell_norm_fwd(b.gl[:, i], b.cl[i], b.cl[j], lnl_w, lnl_b, Dall[:, i], h_out=b.lh[j], bias=pk['bl'],
                                         h2=b.Xq[j][:, :Hd], h3=b.Xl[j][:, ol:ol + Hd], drop=drop, stats=b.statl[i], post_tanh=True))
    timeit('norm_fwd ctx head (rows=64, H=%d, pre_tanh) %s' % (H, tag),
           lambda: be.norm_fwd(b.co[i][:, :H], t['context_att.output_layer.2.weight'], t['context_att.output_layer.2.bias'],
                               y=b.Xl[i][:, :H], stats=b.statc[i, 0], pre_tanh=True, drop=drop))
timeit('gemm Xq.Wq split-K partials', lambda: be.gemm(b.Xq[i], pk['Wq'], b.gq[:, i] if b.Sq > 1 else b.gq[0, i], splitk=b.Sq))
timeit('gemm Xl.Wl split-K partials', lambda: be.gemm(b.Xl[i], pk['Wl'], b.gl[:, i] if b.Sl > 1 else b.gl[0, i], splitk=b.Sl))
timeit('attn2_fwd', lambda: be.attn2_fwd(Kp, Vp, b.q32[i], b.alpha[i], b.co[i], 1.0 / math.sqrt(H), 1))
drops = ((0.3, 11, 0), (0.1, 11, 1 << 32), (0.3, 11, 2 << 32), None)
timeit('whole step (dependent chain, dropout)', lambda: core.step(b, i, j, Kp, Vp, Gq, 1, drops, lang_y=Dall[:, i]))
timeit('whole step (dependent chain, no dropout)', lambda: core.step(b, i, j, Kp, Vp, Gq, 1, lang_y=Dall[:, i]))

# ---- marginal cost of each launch inside the dependent chain (what the 26-step loop actually pays): the chain with one member
# removed; and the pure launch-to-launch period of a trivial dependent kernel
scale = 1.0 / math.sqrt(H)
g_q = lambda: be.gemm(b.Xq[i], pk['Wq'], b.gq[:, i] if b.Sq > 1 else b.gq[0, i], splitk=b.Sq)
c_q = lambda: be.lstm_cell_norm_fwd(b.gq[:, i], b.cq[i], b.cq[j], lnq_w, lnq_b, b.q32[i], h_out=b.qh[i], row_bias=Gq,
                                    h2=b.Xq[j][:, oQ:oQ + Hq], y2=b.Xl[i][:, oq:oq + Hq], stats=b.statq[i])
a_2 = lambda: be.attn2_fwd(Kp, Vp, b.q32[i], b.alpha[i], b.co[i], scale, 1)
g_l = lambda: be.gemm(b.Xl[i], pk['Wl'], b.gl[:, i] if b.Sl > 1 else b.gl[0, i], splitk=b.Sl)
c_l = lambda: be.lstm_cell_norm_fwd(b.gl[:, i], b.cl[i], b.cl[j], lnl_w, lnl_b, Dall[:, i], h_out=b.lh[j], bias=pk['bl'],
                                    h2=b.Xq[j][:, :Hd], h3=b.Xl[j][:, ol:ol + Hd], stats=b.statl[i], post_tanh=True)
tiny = torch.zeros(256, device=dev)
chains = {'chain: gemm_q, cell_q, attn2, gemm_l, cell_l (separate launches)': [g_q, c_q, a_2, g_l, c_l],
          'chain without cell_q': [g_q, a_2, g_l, c_l],
          'chain without cell_l': [g_q, c_q, a_2, g_l],
          'chain without attn2': [g_q, c_q, g_l, c_l],
          'chain: the two GEMMs only': [g_q, g_l],
          'chain: gemm_q only': [g_q],
          'chain: gemm_l only': [g_l],
          'chain: 5 trivial dependent kernels (axpby on 256 floats)': [lambda: be.axpby(tiny, 1.0, tiny, 0.5)] * 5}
for name, fns in chains.items():
    timeit(name, lambda: [f() for f in fns], reps=24)
