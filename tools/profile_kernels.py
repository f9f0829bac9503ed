"""Launch the dominant kernels of the D-LSG training step in isolation at the benched shapes, for
`ncu --set full` captures (one GPU, a handful of launches):

  ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|norm_bwd_vec|norm_fwd_vec|convert2d" \
      -o gpurun_out/prof python tools/profile_kernels.py
"""
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
B = 64
M, N, K = B * 26 * 36, 2048, 2048


def R(*s, dtype=torch.float32):
    return torch.randn(*s, device=dev).to(dtype)


which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('all', 'gemm'):
    a, w, bias = R(M, K, dtype=bf), R(N, K, dtype=bf), R(N)
    o = torch.empty(M, N, device=dev, dtype=bf)
    for _ in range(2):
        be.gemm(a, w, o, bias=bias, tanh=True)              # region projection (both encoders), bf16 out, bias+tanh
    dO = R(M, N, dtype=bf)
    dW = torch.empty(N, K, device=dev)
    be.gemm(dO.t(), a.t(), dW)                               # its weight gradient (K = 59904), both operands read transposed in place
    x = R(B, 2880, dtype=bf)
    wq = R(4096, 2880, dtype=bf)
    g = torch.empty(B, 4096, device=dev)
    be.gemm(x, wq, g)                                        # per-step query-LSTM gate GEMM (swap-AB, auto split-K)
if which in ('all', 'rows'):
    x = R(M, 1024, dtype=bf)
    g_, b_ = R(1024), R(1024)
    y = torch.empty(M, 1024, device=dev, dtype=bf)
    st = torch.empty(M, 2, device=dev)
    for _ in range(2):
        be.norm_fwd(x, g_, b_, y=y, stats=st)                # obj_norm forward (LN on the GEMM's tanh output)
    dy = R(M, 1024, dtype=bf)
    dx = torch.empty(M, 1024, device=dev, dtype=bf)
    dg, db = torch.zeros(1024, device=dev), torch.zeros(1024, device=dev)
    for _ in range(2):
        be.norm_bwd(dy, x, g_, b_, st, dx=dx, dgamma=dg, dbeta=db, in_is_tanh=True)
    r = R(M, 2048)
    d1 = torch.empty(M, 2048, device=dev, dtype=bf)
    be.convert(r, dst=d1)                                    # regions fp32 -> bf16
torch.cuda.synchronize()
