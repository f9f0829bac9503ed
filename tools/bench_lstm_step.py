"""Warm timing of the BiLSTM step chain (B=64, H=1024, 26 steps, both directions) captured in a CUDA graph:
the one-launch step (csrc/lstm_step.cu) with and without its recurrent product, against the split-K tcgen05 GEMM + cell
kernel pair on two streams.  One JSON line per variant (us per time step)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'd-lsg-video-caption_b200'))
from dlsg import ops, linalg as la, functional as DF  # noqa: E402


def main():
    be = ops.CudaBackend()
    dev = 'cuda'
    B, T, H = 64, 26, 1024
    H4 = 4 * H
    g = torch.Generator(device=dev).manual_seed(1)
    W = [(torch.randn(H4, H, device=dev, generator=g) * 0.03).to(torch.bfloat16) for _ in range(2)]
    Gin = torch.randn(B, T, 2 * H4, device=dev, generator=g)
    lstm_out = torch.empty(B, T, 2 * H, device=dev)
    gates = torch.zeros(4, 2, T, B, H4, device=dev)
    cs = torch.zeros(2, T + 1, B, H, device=dev)
    hprev = torch.zeros(2, B, T, H, device=dev, dtype=torch.bfloat16)

    def fused(product=True):
        for k in range(T):
            tt = (k, T - 1 - k)
            nx = (k + 1, T - 2 - k) if k + 1 < T else None
            be.lstm_step_fwd(W, [hprev[d, :, tt[d]] for d in (0, 1)] if (k > 0 and product) else None,
                             [Gin[:, tt[d], d * H4:(d + 1) * H4] for d in (0, 1)], [cs[d, k] for d in (0, 1)],
                             [cs[d, k + 1] for d in (0, 1)], [gates[0, d, tt[d]] for d in (0, 1)],
                             [lstm_out[:, tt[d], d * H:(d + 1) * H] for d in (0, 1)],
                             [hprev[d, :, nx[d]] for d in (0, 1)] if nx is not None else None)

    Sg = la.splitk_for(B, H4, H, sms=74)

    def pair():
        def run_dir(d):
            order = list(range(T)) if d == 0 else list(range(T - 1, -1, -1))
            for k, tt in enumerate(order):
                if k > 0:
                    be.gemm(hprev[d, :, tt], W[d], gates[:Sg, d, tt] if Sg > 1 else gates[0, d, tt], splitk=Sg)
                nxt = order[k + 1] if k + 1 < T else None
                be.lstm_cell_fwd(gates[:Sg, d, tt], cs[d, k], cs[d, k + 1], row_bias=Gin[:, tt, d * H4:(d + 1) * H4],
                                 h2=lstm_out[:, tt, d * H:(d + 1) * H], h3=(hprev[d, :, nxt] if nxt is not None else None))
        DF.two_streams(Gin, lambda: run_dir(0), lambda: run_dir(1))

    for name, fn in (('one-launch step', lambda: fused(True)), ('one-launch step, cell only (no recurrent product)', lambda: fused(False)),
                     ('split-K tcgen05 GEMM + cell kernel, two streams', pair)):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                fn()
            for _ in range(5):
                gr.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                gr.replay()
            b.record()
            torch.cuda.synchronize()
        print(json.dumps({'variant': name, 'us_per_time_step': round(a.elapsed_time(b) / 20 / T * 1e3, 2)}))


if __name__ == '__main__':
    main()
