"""Full GAN training iteration of the live trainer (run_gun.py:147-234 + train_disc :339-398) on our modules, timed
on the GPU (BASELINE.json configs[4] at one GPU: generator + discriminator losses, batch 64, MSR-VTT-shaped).

One iteration = G forward #1 (fake sample) -> num_D_visual=5 x [3 D forwards (real one-hot, fake logits, mixed) +
WGAN-GP double backward + Adam(D)] -> G forward #2 -> packed CE -> D(raw logits) -> total.backward() -> Adam(G).
Eager (the double backward goes through torch autograd over dlsg.generic primitives), CUDA events, one JSON line.

  python tools/bench_gan.py [--batch 64] [--steps 5] [--warmup 2] [--phases]
"""
import argparse
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'd-lsg-video-caption_b200'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--num-d', type=int, default=5)          # opt.py:36 num_D_visual
    ap.add_argument('--vocab', type=int, default=10547)
    ap.add_argument('--phases', action='store_true', help='also time G-fwd / D-loop / G-step separately (extra syncs)')
    ap.add_argument('--graph', type=int, default=0, help='1: dlsg.gan.GanIteration captured as one CUDA graph')
    ap.add_argument('--profile-dstep', action='store_true', help='warm up, then run ONE eager discriminator step inside '
                    'cudaProfilerStart/Stop and exit (ncu --profile-from-start off)')
    a = ap.parse_args()
    from dlsg import synth, losses, ops, linalg as la
    import models.model as M
    dev = torch.device('cuda', 0)
    la.set_precision('bf16')
    args = synth.msr_args(train_batch_size=a.batch)
    B, V, L = a.batch, a.vocab, 26
    torch.manual_seed(12)
    with contextlib.redirect_stdout(io.StringIO()):
        G = M.CapGnnModel(args, synth.Vocab(V)).to(dev).train()
        D = M.DiscV2(args, V).to(dev).train()
    cap = bool(a.graph)
    opt_g = torch.optim.Adam(G.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=cap)     # run_gun.py:91
    opt_d = torch.optim.Adam(D.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=cap)     # run_gun.py:100
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=12)
    frames, regions, caps = frames.to(dev), regions.to(dev), caps.to(dev)
    att_mask = synth.att_mask_from_captions(caps).to(dev)
    eps_tf, lam = 0.6, 0.01                                                              # opt.py:37 lambda_D_visual
    be = ops.backend()
    marks = []

    def mark(name):
        if a.phases:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    def train_disc(real, fake, obj, mot, alpha):
        for _ in range(a.num_d):
            opt_d.zero_grad(set_to_none=True)
            r_logit = D(real, obj, mot, att_mask, alpha)
            f_logit = D(fake, obj, mot, att_mask, alpha)
            e_gp = torch.rand(B, 1, 1, device=dev, requires_grad=True)
            mixed = real.detach() * e_gp + fake.detach() * (1 - e_gp)
            m_logit = D(mixed, obj, mot, att_mask, alpha)
            g = torch.autograd.grad(inputs=mixed, outputs=m_logit, grad_outputs=torch.ones_like(m_logit),
                                    create_graph=True, retain_graph=True)[0]
            gn = g.contiguous().view(B, -1).norm(2, dim=1)
            gp = ((gn - 1) * (gn - 1)).mean()
            loss_d = f_logit.mean() - r_logit.mean() + 10 * gp
            loss_d.backward(retain_graph=True)
            opt_d.step()
        return loss_d.detach()

    def iteration():
        mark('start')
        f_cap, obj, mot, alpha = G(frames, regions, caps, L, eps_tf)
        mark('g_fwd1')
        real = torch.zeros(B, L, V, device=dev).scatter_(2, caps.unsqueeze(2), 1)            # run_gun.py:449-453 to_onehot
        train_disc(real, f_cap.detach(), obj.detach(), mot.detach(), alpha.detach())
        mark('d_loop')
        opt_g.zero_grad(set_to_none=True)
        out, obj, mot, alpha = G(frames, regions, caps, L, eps_tf)
        cap_loss = losses.packed_cross_entropy(out, caps, lens)
        f_logit = D(out, obj.detach(), mot.detach(), att_mask=att_mask, alpha_all=alpha.detach())
        total = cap_loss + (-f_logit.mean()) * lam
        total.backward()
        opt_g.step()
        mark('g_step')
        return total.detach()

    import random
    random.seed(12)
    if a.profile_dstep:
        iteration()
        with torch.no_grad():
            f_cap, obj, mot, alpha = G(frames, regions, caps, L, eps_tf)
        real = torch.zeros(B, L, V, device=dev).scatter_(2, caps.unsqueeze(2), 1)
        a.num_d = 1
        train_disc(real, f_cap, obj, mot, alpha)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        train_disc(real, f_cap, obj, mot, alpha)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    if a.graph:
        from dlsg.gan import GanIteration
        it = GanIteration(G, D, opt_g, opt_d, frames, regions, caps, lens, L, eps_tf, a.num_d, lam, graph=True)
        iteration = lambda: it()[0]
    for _ in range(a.warmup):
        iteration()
    torch.cuda.synchronize()
    marks.clear()
    l0 = be.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    line = {'metric': 'GAN iteration (G fwd + %d x D step with WGAN-GP + G step), B=%d MSR-VTT-shaped' % (a.num_d, B),
            'ms_per_iteration': ms, 'clips_per_s': B / (ms * 1e-3), 'steps': a.steps, 'warmup': a.warmup,
            'libdlsg_launches_per_iteration': (it.launches if a.graph else (be.launches - l0) // a.steps), 'loss': float(loss), 'mode': 'cuda_graph' if a.graph else 'eager'}
    if a.phases:
        ph = {}
        for (n0, ev0), (n1, ev1) in zip(marks[:-1], marks[1:]):
            if n1 != 'start':
                ph[n1] = ph.get(n1, 0.0) + ev0.elapsed_time(ev1) / a.steps
        line['phase_ms'] = ph
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
