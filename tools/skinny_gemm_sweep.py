"""Warm timing sweep of the skinny (batch-64) recurrent GEMMs: auto split-K with the in-kernel fix-up, explicit split-K
partial outputs for S = 1..6, HBM-streamed (8 rotating weight copies) vs L2-resident (one copy).  CUDA events around a
CUDA graph of back-to-back launches; one JSON line per case."""
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


def timeit(name, fn, reps=48, bytes_=None):
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
    out = {'case': name, 'us': round(best, 2)}
    if bytes_:
        out['GB/s'] = round(bytes_ / best / 1e3, 1)
    print(json.dumps(out), flush=True)


B = 64
for (N, K, tag) in ((4096, 2880, 'Wq'), (6144, 4608, 'Wl'), (4096, 1024, 'bilstm Whh'), (2880, 4096, 'dXq'), (4608, 6144, 'dXl')):
    for ncopies in (8, 1):
        x = torch.randn(B, K, device=dev).to(bf)
        ws = [torch.randn(N, K, device=dev).to(bf) for _ in range(ncopies)]
        out = torch.empty(B, N, device=dev)
        cnt = [0]

        def f():
            cnt[0] += 1
            be.gemm(x, ws[cnt[0] % ncopies], out)
        where = 'hbm' if ncopies > 1 else 'l2'
        timeit('auto   %-10s N=%d K=%d %s' % (tag, N, K, where), f, bytes_=N * K * 2)
        kb = (K + 63) // 64
        for S in (1, 2, 3, 4, 6):
            per = (kb + S - 1) // S
            if (kb + per - 1) // per != S:
                continue
            part = torch.empty(S, B, N, device=dev)

            def f2():
                cnt[0] += 1
                if S == 1:
                    old = be._workspace
                    be._workspace = lambda d: torch.empty(0, dtype=torch.uint8, device=d)
                    be.gemm(x, ws[cnt[0] % ncopies], part[0])
                    be._workspace = old
                else:
                    be.gemm(x, ws[cnt[0] % ncopies], part, splitk=S)
            timeit('part S=%d %-8s N=%d K=%d %s' % (S, tag, N, K, where), f2, bytes_=N * K * 2)
        del ws
# floor: a device-to-device copy of the same number of bytes (what "touch the weights once" costs)
for mb in (8, 24, 57):
    src = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    dst = torch.empty_like(src)
    timeit('memcpy d2d %d MB (read+write)' % mb, lambda: dst.copy_(src), bytes_=2 * (mb << 20))
