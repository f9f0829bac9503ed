"""Warm timings of the vocabulary-row kernels (beam top-k, arg-max, log-softmax, masked CE) at the decode / train shapes,
rotating over buffers larger than L2.  One JSON line per case."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT):
    sys.path.insert(0, p)
import torch
from dlsg import ops
be = ops.backend()
dev = 'cuda'
try:
    HBM = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    HBM = 6650.0


def timeit(fn, reps=12):
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
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


rows, V = 640, 10547
bufs = [torch.randn(rows, V, device=dev) for _ in range(6)]
cnt = [0]
def nxt():
    cnt[0] += 1
    return bufs[cnt[0] % 6]
last_live = torch.full((rows,), 7, dtype=torch.int64, device=dev)
last_mixed = torch.randint(0, 6, (rows,), device=dev)
for k in (1, 5):
    tl, ti = torch.zeros(rows, k, device=dev), torch.zeros(rows, k, dtype=torch.int64, device=dev)
    for name, last, norm in (('all rows live', last_live, True), ('1/6 rows finished', last_mixed, True), ('no normalisation', last_live, False)):
        us = timeit(lambda: be.beam_topk(nxt(), last, 2, k, tl, ti, normalize=norm))
        print(json.dumps({'kernel': 'beam_topk_kernel<%d>' % k, 'case': name, 'rows': rows, 'V': V, 'us': round(us, 2),
                          'GBps': round(rows * V * 4 / us / 1e3, 1), 'frac_of_hbm': round(rows * V * 4 / us / 1e3 / HBM, 3)}))
ids = torch.zeros(rows, dtype=torch.int64, device=dev)
us = timeit(lambda: be.row_argmax(nxt(), ids))
print(json.dumps({'kernel': 'row_argmax_kernel', 'rows': rows, 'V': V, 'us': round(us, 2), 'GBps': round(rows * V * 4 / us / 1e3, 1),
                  'frac_of_hbm': round(rows * V * 4 / us / 1e3 / HBM, 3)}))
out = torch.empty(rows, V, device=dev)
us = timeit(lambda: be.log_softmax(nxt(), out))
print(json.dumps({'kernel': 'log_softmax_kernel', 'rows': rows, 'V': V, 'us': round(us, 2), 'GBps': round(rows * V * 8 / us / 1e3, 1),
                  'frac_of_hbm': round(rows * V * 8 / us / 1e3 / HBM, 3)}))
us = timeit(lambda: out.copy_(nxt()))
print(json.dumps({'kernel': 'torch copy (reference)', 'rows': rows, 'V': V, 'us': round(us, 2), 'GBps': round(rows * V * 8 / us / 1e3, 1)}))
B, L = 64, 26
lg = [torch.randn(B, L, V, device=dev) for _ in range(3)]
tg = torch.randint(0, V, (B, L), device=dev)
lens = torch.randint(4, 27, (B,), dtype=torch.int32, device=dev)
full = torch.full((B,), 26, dtype=torch.int32, device=dev)
loss, dl = torch.zeros(1, device=dev), torch.empty(B, L, V, device=dev)
c2 = [0]
def nl():
    c2[0] += 1
    return lg[c2[0] % 3]
for name, ln in (('ragged captions (len ~ U{4..26})', lens), ('all 26 tokens counted', full)):
    us = timeit(lambda: be.ce_masked(nl(), tg, ln, loss, dl, 1.0 / 1000))
    cnt_rows = int(ln.sum())
    byts = cnt_rows * V * 8 + (B * L - cnt_rows) * V * 4
    print(json.dumps({'kernel': 'ce_masked_kernel', 'case': name, 'us': round(us, 2), 'algorithmic_mb': round(byts / 1e6, 1),
                      'GBps': round(byts / us / 1e3, 1), 'frac_of_hbm': round(byts / us / 1e3 / HBM, 3)}))
