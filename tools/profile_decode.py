"""Greedy (B=256) / beam-5 (B=128) decode at MSR-VTT shapes: CUDA-event timing, or one decode inside
cudaProfilerStart/Stop for `ncu --profile-from-start off` launch lists.

  python tools/profile_decode.py time
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
      python tools/profile_decode.py profile greedy
"""
import contextlib
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
from dlsg import synth, ops, linalg as la  # noqa: E402
import models.model as M  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'time'
which = sys.argv[2] if len(sys.argv) > 2 else 'both'
dev = torch.device('cuda')
la.set_precision('bf16')
args = synth.msr_args()
V = 10547
with contextlib.redirect_stdout(io.StringIO()):
    net = M.CapGnnModel(args, synth.Vocab(V)).to(dev).eval()
be = ops.backend()
cases = [('greedy', 256, 1), ('beam5', 128, 5)]
with torch.no_grad():
    for name, B, beam in cases:
        if which not in ('both', name):
            continue
        fr, rg, _, _ = synth.make_inputs(B, args, V, seed=7)
        fr, rg = fr.to(dev), rg.to(dev)
        net.update_beam_size(beam)
        for _ in range(2):
            net(fr, rg, None)
        torch.cuda.synchronize()
        if mode == 'profile':
            torch.cuda.profiler.start()
            net(fr, rg, None)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            continue
        l0 = be.launches
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            out = net(fr, rg, None)[0]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        wall = (time.perf_counter() - t0) / n * 1e3
        print('%s B=%d: %.2f ms/batch (wall %.2f ms) -> %.0f captions/s, %d launches/batch, out %s' %
              (name, B, ms, wall, B / ms * 1e3, (be.launches - l0) // n, tuple(out.shape)), flush=True)
