"""Microbenchmark of the fused region aggregation kernels (csrc/region_agg.cu) at the benchmarked shape (B=64, T=26, R=36,
H=1024, both encoders): warm launches back to back (inputs 245 MB per pass > L2), CUDA events, algorithmic bytes / time
against the measured HBM peak.  Prints one JSON line per kernel."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'd-lsg-video-caption_b200'))
from dlsg import ops  # noqa: E402


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    be = ops.CudaBackend()
    B, T, R, H, E = 64, 26, 36, 1024, 2
    TR = T * R
    dev = 'cuda'
    peak = 6543.1
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    g = torch.Generator(device=dev).manual_seed(1)
    Ybuf = torch.tanh(torch.randn(B * TR, E * H, device=dev, generator=g)).to(torch.bfloat16)
    Y = [Ybuf[:, e * H:(e + 1) * H] for e in range(E)]
    mk = lambda *s: [torch.randn(*s, device=dev, generator=g) for _ in range(E)]
    z = lambda *s: [torch.zeros(*s, device=dev) for _ in range(E)]
    F, dA = mk(B * T, H), mk(B * T, H)
    gamma = [1 + 0.1 * x for x in mk(H)]
    beta = [0.1 * x for x in mk(H)]
    scale = 1 / math.sqrt(2048)
    o = dict(agg=z(B * T, H), U=z(B * T, H), stats=z(B * TR, 2), St=z(B, T, TR), tconst=z(B * T, 4))
    work = [torch.empty(be.region_aggregate_bwd_workspace(B, T, TR), dtype=torch.uint8, device=dev) for _ in range(E)]
    dSm, tcA = z(B, T, TR), z(B * T, 4)
    dbuf = torch.zeros(B * TR, E * H, dtype=torch.bfloat16, device=dev)
    dpre = [dbuf[:, e * H:(e + 1) * H] for e in range(E)]
    dF, dg, db, dbias = z(B * T, H), z(H), z(H), z(H)
    ybytes = B * TR * H * 2 * E
    runs = [
        ('region_aggregate_fwd_kernel (training: + stats, scores, U)', ybytes,
         lambda: be.region_aggregate_fwd(Y, F, gamma, beta, scale, T, **o)),
        ('region_aggregate_fwd_kernel (inference: aggregate only)', ybytes,
         lambda: be.region_aggregate_fwd(Y, F, gamma, beta, scale, T, agg=o['agg'])),
        ('region_aggregate_fwd_kernel (scores pass of the backward)', ybytes,
         lambda: be.region_aggregate_fwd(Y, dA, gamma, beta, scale, T, St=dSm, tconst=tcA, scores_only=True)),
        ('region_aggregate_fwd_kernel (loads only: the bulk-copy ring without arithmetic)', ybytes,
         lambda: be.region_aggregate_fwd(Y, dA, gamma, beta, scale, T, St=dSm, tconst=tcA, scores_only=2)),
        ('region_aggregate_prep_kernel + region_aggregate_bwd_kernel', 2 * ybytes,
         lambda: be.region_aggregate_bwd(Y, o['stats'], o['St'], dSm, F, dA, o['U'], o['tconst'], tcA, gamma, beta, scale, T,
                                         dpre=dpre, dF=dF, dgamma=dg, dbeta=db, dbias=dbias, work=work)),
    ]
    for name, nbytes, fn in runs:
        us = timeit(fn)
        gbs = nbytes / us / 1e3
        print(json.dumps({'kernel': name, 'us_per_launch': round(us, 2), 'algorithmic_mb': round(nbytes / 1e6, 1),
                          'achieved': round(gbs, 1), 'peak': peak, 'unit': 'GB/s', 'frac': round(gbs / peak, 3), 'bound': 'hbm'}))


if __name__ == '__main__':
    main()
