"""Warm per-kernel timings (CUDA events, back-to-back launches inside one CUDA graph so host launch cost is excluded).
Complements the ncu launch list, whose per-launch numbers are cold-cache / serialised.  Prints one JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
from dlsg import ops  # noqa: E402

dev = 'cuda'
be = ops.backend()


def timeit(name, fn, reps=50, bytes_=None, flops=None):
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
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    out = {'case': name, 'us': round(us, 2)}
    if bytes_:
        out['GB/s'] = round(bytes_ / us / 1e3, 1)
    if flops:
        out['TFLOP/s'] = round(flops / us / 1e6, 1)
    print(json.dumps(out), flush=True)


def R(*s, dtype=torch.float32):
    return torch.randn(*s, device=dev).to(dtype)


bf = torch.bfloat16
B = 64
# ---- skinny recurrent GEMMs (weights streamed; rotate over 4 weight copies so L2 does not hide HBM entirely)
for (N, K, tag) in ((4096, 2880, 'Wq'), (6144, 4608, 'Wl'), (2048, 1024, 'Wqp'), (2880, 4096, 'dXq'), (4608, 6144, 'dXl')):
    x = R(B, K, dtype=bf)
    ws = [R(N, K, dtype=bf) for _ in range(4)]
    out = torch.empty(B, N, device=dev)
    cnt = [0]

    def f():
        cnt[0] += 1
        be.gemm(x, ws[cnt[0] % 4], out)
    timeit('gemm_tc skinny auto-splitk %s M=64 N=%d K=%d' % (tag, N, K), f, bytes_=N * K * 2)
    # direct (no split) path: call with workspace disabled by passing splitk through a 1-split buffer
    g1 = torch.empty(B, N, device=dev)
    old = be._workspace
    be._workspace = lambda d: torch.empty(0, dtype=torch.uint8, device=d)

    def f3():
        cnt[0] += 1
        be.gemm(x, ws[cnt[0] % 4], g1)
    timeit('gemm_tc skinny no-split   %s M=64 N=%d K=%d' % (tag, N, K), f3, bytes_=N * K * 2)
    be._workspace = old

# ---- big GEMMs
for (M, N, K, tag) in ((59904, 2048, 2048, 'region fwd'), (2048, 2048, 59904, 'region wgrad'), (1664, 10547, 1536, 'vocab'),
                       (1664, 8192, 1024, 'bilstm in')):
    a, b = R(M, K, dtype=bf), R(N, K, dtype=bf)
    o = torch.empty(M, N, device=dev, dtype=bf if tag == 'region fwd' else torch.float32)
    timeit('gemm_tc %s %dx%dx%d' % (tag, M, N, K), lambda: be.gemm(a, b, o), reps=10, flops=2.0 * M * N * K)

# ---- row kernels at per-step size and at region size
for rows, D, dt in ((64, 1024, torch.float32), (64, 1536, torch.float32), (59904, 1024, bf), (1664, 2048, torch.float32)):
    x = R(rows, D, dtype=dt)
    g_, b_ = R(D), R(D)
    y = torch.empty(rows, D, device=dev, dtype=dt)
    st = torch.empty(rows, 2, device=dev)
    es = 2 if dt == bf else 4
    timeit('norm_fwd rows=%d D=%d %s' % (rows, D, dt), lambda: be.norm_fwd(x, g_, b_, y=y, stats=st, pre_tanh=True), reps=20,
           bytes_=rows * D * es * 2)
    be.norm_fwd(x, g_, b_, y=y, stats=st)
    dy = R(rows, D, dtype=dt)
    dx = torch.empty(rows, D, device=dev, dtype=dt)
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    timeit('norm_bwd rows=%d D=%d %s' % (rows, D, dt), lambda: be.norm_bwd(dy, x, g_, b_, st, dx=dx, dgamma=dg, dbeta=db, pre_tanh=True),
           reps=20, bytes_=rows * D * es * 3)
    timeit('norm_bwd(+dropout 0.3) rows=%d D=%d' % (rows, D), lambda: be.norm_bwd(dy, x, g_, b_, st, dx=dx, dgamma=dg, dbeta=db, drop=(0.3, 5, 0)),
           reps=20, bytes_=rows * D * es * 3)

# ---- LSTM cell / attention at per-step size
H = 1024
gates = R(B, 4 * H)
cprev, cout, hout = R(B, H), torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)
rb = R(B, 4 * H)
timeit('lstm_cell_fwd B=64 H=1024', lambda: be.lstm_cell_fwd(gates, cprev, cout, h_out=hout, row_bias=rb))
dgt = torch.empty(B, 4 * H, device=dev)
timeit('lstm_cell_bwd B=64 H=1024', lambda: be.lstm_cell_bwd(gates, cprev, cout, hout, None, cprev, dgates=dgt))
Kp, Vp, qp = R(2, B, 5, H), R(2, B, 5, H), R(B, 2 * H)
al, ctx = torch.empty(B, 10, device=dev), torch.empty(B, 2 * H, device=dev, dtype=bf)
timeit('node_attn_fwd rows=64 nh=2 P=5', lambda: be.node_attn_fwd(Kp, Vp, qp, al, ctx, 1))
dK, dV, dq = torch.zeros_like(Kp), torch.zeros_like(Vp), torch.empty(B, 2 * H, device=dev, dtype=bf)
dctx = R(B, 2 * H)
timeit('node_attn_bwd rows=64 nh=2 P=5', lambda: be.node_attn_bwd(Kp, Vp, qp, al, dctx, dq, dK, dV))
# ---- conversions
x = R(59904, 2048)
d1, d2 = torch.empty(59904, 2048, device=dev, dtype=bf), torch.empty(2048, 59904, device=dev, dtype=bf)
timeit('convert2d cast+transpose regions 59904x2048', lambda: be.convert(x, dst=d1, dstT=d2), reps=5, bytes_=59904 * 2048 * 8)
timeit('convert2d cast only 59904x2048', lambda: be.convert(x, dst=d1), reps=5, bytes_=59904 * 2048 * 6)
# ---- vocab rows
lg = R(640, 10547)
tl, ti = torch.empty(640, 5, device=dev), torch.empty(640, 5, device=dev, dtype=torch.int64)
last = torch.zeros(640, device=dev, dtype=torch.int64)
timeit('beam_topk rows=640 V=10547 k=5', lambda: be.beam_topk(lg, last, 2, 5, tl, ti), bytes_=640 * 10547 * 4)
