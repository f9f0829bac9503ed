"""Where does a skinny tcgen05 GEMM launch spend its time?  Uses dlsg_debug_gemm_trace: per-CTA phase timestamps of the
LAST launch of a CUDA graph of back-to-back launches, next to the measured launch-to-launch period of that graph.

phases (ns relative to the earliest CTA entry of the launch, median / max over CTAs):
 0 CTA entry  1 barriers+TMEM ready  2 first TMA issued  3 all TMA issued  4 first operands landed
 5 all MMAs issued  6 last accumulator complete  7 epilogue stored (CTA exit)
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
trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
B = 64
for (N, K, S, tag) in ((4096, 2880, 4, 'Wq'), (6144, 4608, 3, 'Wl'), (4096, 1024, 4, 'bilstm Whh'), (4096, 1024, 1, 'bilstm Whh'),
                       (4608, 6144, 4, 'dXl')):
    x = torch.randn(B, K, device=dev).to(bf)
    w = torch.randn(N, K, device=dev).to(bf)
    part = torch.empty(S, B, N, device=dev)

    def f():
        if S == 1:
            old = be._workspace
            be._workspace = lambda d: torch.empty(0, dtype=torch.uint8, device=d)
            be.gemm(x, w, part[0])
            be._workspace = old
        else:
            be.gemm(x, w, part, splitk=S)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    reps = 32
    g = torch.cuda.CUDAGraph()
    be.lib.dlsg_debug_gemm_trace(trace.data_ptr())
    with torch.cuda.graph(g):
        for _ in range(reps):
            f()
    be.lib.dlsg_debug_gemm_trace(None)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    trace.zero_()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    period = e0.elapsed_time(e1) * 1e3 / reps
    t = trace.view(148, 8, 2).cpu()
    used = t[:, 0, 0] > 0
    gt = t[used][:, :, 0].double()
    ck = t[used][:, :, 1].double()
    t0 = gt[:, 0].min()
    rel = gt - t0
    med = rel.median(dim=0).values.tolist()
    mx = rel.max(dim=0).values.tolist()
    cyc = (ck - ck[:, :1]).median(dim=0).values.tolist()
    print(json.dumps({'case': '%s N=%d K=%d S=%d' % (tag, N, K, S), 'ctas': int(used.sum()), 'period_us': round(period, 2),
                      'busy_us(max exit - min entry)': round(mx[7] / 1e3, 2),
                      'phase_ns_median': [int(v) for v in med], 'phase_ns_max': [int(v) for v in mx],
                      'phase_cycles_median(clock64)': [int(v) for v in cyc]}), flush=True)
