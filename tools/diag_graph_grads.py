"""Diagnostic: per-parameter gradient agreement of (captured step | eager CUDA step | CPU oracle) at MSR widths, B=4."""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT):
    sys.path.insert(0, p)
import torch
from dlsg import synth, linalg as la, losses
from dlsg.graphs import GraphedTrainStep
from oracle import dlsg_oracle as O
import models.model as M

dev = 'cuda'
la.set_precision('bf16')
args, V, B = synth.msr_args(), 10547, 4
with contextlib.redirect_stdout(io.StringIO()):
    net = M.CapGnnModel(args, synth.Vocab(V))
synth.fill_state_dict(net)
sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and not k.endswith('pe.pe')) for k, v in net.state_dict().items()}
fr, rg, cp, lens = synth.make_inputs(B, args, V, seed=300)
out = O.cap_gnn_forward(sd, fr, rg, cp, 26, 1.0, args.a_feature_size)[0]
O.packed_ce_loss(out, cp, lens).backward()
orc = {k: v.grad for k, v in sd.items() if v.grad is not None}
net = net.to(dev).eval()
sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
fr, rg, cp = fr.to(dev), rg.to(dev), cp.to(dev)


def eager():
    net.zero_grad(set_to_none=True)
    o = net(fr, rg, cp, 26, 1.0)[0]
    losses.packed_cross_entropy(o, cp, lens).backward()
    return {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
e1 = eager()
e2 = eager()
torch.cuda.synchronize()
net.zero_grad(set_to_none=True)
opt = torch.optim.Adam(net.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=True)
gs = GraphedTrainStep(net, opt, fr, rg, cp, lens, 26, 1.0, warmup=0)
net.load_state_dict(sd0)
gs.refresh_weights()
gs()
torch.cuda.synchronize()
g1 = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}


def rel(a, b):
    return float((a.float().cpu() - b.float().cpu()).norm() / (b.float().norm().cpu() + 1e-20))
rows = []
for k in orc:
    if k in e1:
        rows.append((rel(g1[k], e1[k]), rel(e2[k], e1[k]), rel(e1[k], orc[k]), float(orc[k].norm()), k))
rows.sort(reverse=True)
print('graph-vs-eager  eager-vs-eager  eager-vs-oracle  |oracle grad|  name')
for r in rows[:12]:
    print('%.3e  %.3e  %.3e  %.3e  %s' % r)
rows.sort(key=lambda r: -r[2])
print('--- worst eager-vs-oracle')
for r in rows[:12]:
    print('%.3e  %.3e  %.3e  %.3e  %s' % r)
