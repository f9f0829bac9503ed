"""Diagnostic (2 ranks, torchrun): replay time of the critic's gradient averaging (GradSync.reduce_params: pack ->
all-reduce -> unpack) captured in a CUDA graph, and of its parts."""
import contextlib, io, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
from dlsg import synth, functional as DF, linalg as la
import models.model as M

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
dev = torch.device('cuda')
la.set_precision('bf16')
args = synth.msr_args()
with contextlib.redirect_stdout(io.StringIO()):
    D = M.DiscV2(args, 10547).to(dev)
params = [p for p in D.parameters()]
for p in params:
    p.grad = torch.randn_like(p)
sync = DF.GradSync(dist.group.WORLD)
flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.bfloat16, device=dev)


def timeit(fn, name):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(5):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            g.replay()
        b.record()
        torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({'case': name, 'ms_per_call': round(a.elapsed_time(b) / 50, 4)}), flush=True)


timeit(lambda: sync.reduce_params(params), 'reduce_params (pack + all-reduce + unpack), %d params, %.1f MB bf16' % (len(params), flat.numel() * 2 / 1e6))
timeit(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG), 'all_reduce(flat bf16, AVG) alone')
timeit(lambda: dist.all_reduce(flat, op=dist.ReduceOp.SUM), 'all_reduce(flat bf16, SUM) alone')
f32 = flat.float()
timeit(lambda: dist.all_reduce(f32, op=dist.ReduceOp.SUM), 'all_reduce(flat fp32, SUM) alone')
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
