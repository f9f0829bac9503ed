"""Per-kernel roofline table of the hot path at the bench shapes (B=64, MSR-VTT-shaped): every kernel class timed warm
(back-to-back launches replayed inside one CUDA graph, CUDA events), algorithmic bytes / flops per launch divided by
the time, against the measured peaks (MEASURED_PEAKS.json, else the B200_PROFILING.md fallback: 6.65 TB/s, 1590 TFLOP/s).
HBM-bound cases use working sets larger than the 126 MB L2 (or rotate over several buffers).  One JSON line per kernel.

  python tools/roofline_table.py > gpurun_out/roofline.jsonl
"""
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
bf = torch.bfloat16
try:
    PEAKS = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    SRC = 'measured'
except Exception:
    PEAKS, SRC = {}, 'fallback'
HBM = float(PEAKS.get('hbm_gbs', 6650.0))
TF = float(PEAKS.get('bf16_tflops', 1590.0))


def timeit(fn, reps):
    for _ in range(2):
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
    return e0.elapsed_time(e1) * 1e3 / reps


def row(kernel, used_for, bound, us, bytes_=None, flops=None, note=None):
    o = {'kernel': kernel, 'used_for': used_for, 'bound': bound, 'us_per_launch': round(us, 2)}
    if bound == 'tensor':
        ach = flops / us / 1e6
        o.update(algorithmic_gflop=round(flops / 1e9, 2), achieved=round(ach, 1), peak=TF, unit='TFLOP/s', frac=round(ach / TF, 3))
    elif bound in ('hbm', 'l2'):
        ach = bytes_ / us / 1e3
        o.update(algorithmic_mb=round(bytes_ / 1e6, 2), achieved=round(ach, 1), peak=HBM, unit='GB/s', frac=round(ach / HBM, 3))
    else:
        o.update(algorithmic_mb=round((bytes_ or 0) / 1e6, 3))
    o['peak_source'] = SRC
    if note:
        o['note'] = note
    print(json.dumps(o), flush=True)


def R(*s, dtype=torch.float32):
    return torch.randn(*s, device=dev).to(dtype)


B = 64
M = B * 26 * 36
# ---- tensor-bound GEMMs
a, w, bias = R(M, 2048, dtype=bf), R(2048, 2048, dtype=bf), R(2048)
o16 = torch.empty(M, 2048, device=dev, dtype=bf)
row('gemm_tc_kernel<256>', 'region projection fwd (both encoders), bias+tanh, bf16 out', 'tensor',
    timeit(lambda: be.gemm(a, w, o16, bias=bias, tanh=True), 10), flops=2.0 * M * 2048 * 2048)
dO = R(M, 2048, dtype=bf)
dW = torch.empty(2048, 2048, device=dev)
row('gemm_tc_kernel<256> (both operands MN-major)', 'region projection weight gradient, K=59904', 'tensor',
    timeit(lambda: be.gemm(dO.t(), a.t(), dW), 10), flops=2.0 * M * 2048 * 2048, note='128 tiles on 148 SMs')
x, wv = R(1664, 1536, dtype=bf), R(10547, 1536, dtype=bf)
lg = torch.empty(1664, 10547, device=dev)
row('gemm_tc_kernel<256>', 'vocabulary projection (teacher forced, all 26 steps)', 'tensor',
    timeit(lambda: be.gemm(x, wv, lg), 10), flops=2.0 * 1664 * 10547 * 1536, note='output 70 MB fp32, pitch not 16-byte aligned')
del dO, dW
# ---- HBM-bound streaming kernels (working sets > L2)
r32 = R(M, 2048)
row('cast_f32_bf16_flat', 'regions fp32 -> bf16', 'hbm', timeit(lambda: be.convert(r32, dst=a), 5), bytes_=M * 2048 * 6)
del r32
t16 = R(M, 1024, dtype=bf)
g_, b_ = R(1024), R(1024)
y16 = torch.empty(M, 1024, device=dev, dtype=bf)
st = torch.empty(M, 2, device=dev)
row('norm_fwd_bf16_kernel<4>', 'obj_norm fwd on the region activations', 'hbm',
    timeit(lambda: be.norm_fwd(t16, g_, b_, y=y16, stats=st), 10), bytes_=M * 1024 * 4)
dy16, dx16 = R(M, 1024, dtype=bf), torch.empty(M, 1024, device=dev, dtype=bf)
dg, db, dbias = torch.zeros(1024, device=dev), torch.zeros(1024, device=dev), torch.zeros(1024, device=dev)
row('norm_bwd_bf16_kernel<4>', 'obj_norm bwd (+tanh derivative, +bias-gradient column sums)', 'hbm',
    timeit(lambda: be.norm_bwd(dy16, t16, g_, b_, st, dx=dx16, dgamma=dg, dbeta=db, in_is_tanh=True, dxsum=dbias), 10),
    bytes_=M * 1024 * 6)
del t16, y16, dy16, dx16, a, o16
n = 117_000_000 // 1024
P_, G_, M_, V_ = (R(n, 1024) for _ in range(4))
V_.abs_()
S16 = torch.empty(n, 1024, device=dev, dtype=bf)
step = torch.ones((), device=dev)
plan = be.make_adam_plan([dict(p=P_, g=G_, m=M_, v=V_, dst=S16)])
row('adam_multi_kernel', 'Adam over 117 M parameters + bf16 operand copies', 'hbm',
    timeit(lambda: be.adam_multi(plan, step, 1.6e-4, 0.5, 0.9, 1e-8), 3), bytes_=n * 1024 * 30)
del P_, G_, M_, V_, S16
logits = R(1664, 10547)
tg = torch.randint(4, 10547, (64, 26), device=dev)
lens = torch.full((64,), 20, dtype=torch.int32, device=dev)
loss, dl = torch.zeros(1, device=dev), torch.empty(64, 26, 10547, device=dev)
row('ce_masked_kernel', 'fused masked cross-entropy fwd + d(logits)', 'hbm',
    timeit(lambda: be.ce_masked(logits.view(64, 26, 10547), tg, lens, loss, dl, 1.0 / 1280), 10), bytes_=1664 * 10547 * 8,
    note='140 MB per launch barely exceeds L2: partly L2-resident')
bl = [R(640, 10547) for _ in range(6)]
tl, ti = torch.empty(640, 5, device=dev), torch.empty(640, 5, device=dev, dtype=torch.int64)
last = torch.zeros(640, device=dev, dtype=torch.int64)
cnt = [0]


def topk():
    cnt[0] += 1
    be.beam_topk(bl[cnt[0] % 6], last, 2, 5, tl, ti)


row('beam_topk_kernel<5>', 'beam step: log-softmax + after-<end> forcing + top-5 of V=10547, 640 rows', 'hbm', timeit(topk, 12),
    bytes_=640 * 10547 * 4)
del bl, logits, dl
# ---- weight-streaming skinny GEMMs of the recurrent loops (weights stay in L2 across steps in the real loop: rotate 4 copies)
for (N, K, tag, atomic) in ((4096, 2864, 'query-LSTM gates Xq.Wq^T', False), (6144, 4608, 'lang-LSTM gates Xl.Wl^T', False),
                            (4608, 6144, 'decoder dgrad dgl.Wl (atomic split-K)', True), (4096, 1024, 'BiLSTM h.Whh^T', False)):
    xs = R(B, K, dtype=bf)
    ws = [R(N, K, dtype=bf) for _ in range(4)]
    out = torch.zeros(B, N, device=dev)
    cnt = [0]

    def f():
        cnt[0] += 1
        be.gemm(xs, ws[cnt[0] % 4], out, atomic=atomic)
    row('gemm_tc_kernel<64> swap-AB split-K', tag + ' (M=64, N=%d, K=%d)' % (N, K), 'l2', timeit(f, 48), bytes_=N * K * 2,
        note='weight streaming; reported against the HBM peak')
    del ws
# ---- latency-bound per-step kernels (64 rows)
H = 1536
gates = R(3, B, 4 * H)
cprev, cout, hout = R(B, H), torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)
gam, bet = R(H), R(H)
y = torch.empty(B, H, device=dev, dtype=bf)
row('lstm_cell_norm_fwd_kernel', 'lang-LSTM cell + LayerNorm + tanh (split-K partials in), 64 rows', 'latency',
    timeit(lambda: be.lstm_cell_norm_fwd(gates, cprev, cout, gam, bet, y, h_out=hout, post_tanh=True), 48), bytes_=B * H * 4 * 16)
