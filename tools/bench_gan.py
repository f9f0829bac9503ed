"""Full GAN training iteration of the live trainer (run_gun.py:147-234 + train_disc :339-398) on our modules, timed
on the GPU (BASELINE.json configs[4] at one GPU: generator + discriminator losses, batch 64, MSR-VTT-shaped).

One iteration = G forward #1 (fake sample) -> num_D_visual=5 x [D(real one-hot), D(fake logits), D(mixed) + WGAN-GP
double backward + Adam(D)] -> G forward #2 -> packed CE -> D(raw logits) -> total.backward() -> Adam(G)
(dlsg.gan.GanIteration).  CUDA events, one JSON line.

  python tools/bench_gan.py [--graph 1] [--batched 1] [--batch 64] [--steps 5] [--warmup 2]
  ncu ... --profile-from-start off python tools/bench_gan.py --profile-dstep      # launch list of ONE critic step
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
    ap.add_argument('--graph', type=int, default=1, help='1: the iteration captured as one CUDA graph; 0: eager')
    ap.add_argument('--overlap-g', type=int, default=1, help='1: G forward #2 on a side stream under the critic steps')
    ap.add_argument('--batched', type=int, default=1, help='1: the three critic calls of a step as one stacked forward')
    ap.add_argument('--profile-dstep', action='store_true', help='warm up, then run ONE eager critic step inside '
                    'cudaProfilerStart/Stop and exit (ncu --profile-from-start off)')
    a = ap.parse_args()
    from dlsg import synth, ops, linalg as la
    from dlsg.gan import GanIteration
    import models.model as M
    dev = torch.device('cuda', 0)
    la.set_precision('bf16')
    args = synth.msr_args(train_batch_size=a.batch)
    B, V, L = a.batch, a.vocab, 26
    torch.manual_seed(12)
    with contextlib.redirect_stdout(io.StringIO()):
        G = M.CapGnnModel(args, synth.Vocab(V)).to(dev).train()
        D = M.DiscV2(args, V).to(dev).train()
    cap = bool(a.graph) and not a.profile_dstep
    opt_g = torch.optim.Adam(G.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=cap)     # run_gun.py:91
    opt_d = torch.optim.Adam(D.parameters(), lr=1.6e-4, betas=(0.5, 0.9), fused=True, capturable=cap)     # run_gun.py:100
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=12)
    frames, regions, caps = frames.to(dev), regions.to(dev), caps.to(dev)
    eps_tf, lam = 0.6, 0.01                                                              # opt.py:37 lambda_D_visual
    be = ops.backend()
    import random
    random.seed(12)
    if a.profile_dstep:
        it = GanIteration(G, D, opt_g, opt_d, frames, regions, caps, lens, L, eps_tf, 1, lam, graph=False, batched=bool(a.batched))
        it()
        with torch.no_grad():
            f_cap, obj, mot, alpha = G(frames, regions, caps, L, eps_tf)
        # token form (batched critic calls) takes the caption ids; the literal form the (B,L,V) one-hot (run_gun.py:449-453)
        real = caps if a.batched else torch.zeros(B, L, V, device=dev).scatter_(2, caps.unsqueeze(2), 1)
        att_mask = synth.att_mask_from_captions(caps).to(dev)
        it._disc_steps(real, f_cap, obj, mot, att_mask, alpha)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        it._disc_steps(real, f_cap, obj, mot, att_mask, alpha)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    it = GanIteration(G, D, opt_g, opt_d, frames, regions, caps, lens, L, eps_tf, a.num_d, lam, graph=bool(a.graph), overlap_g=bool(a.overlap_g),
                      batched=bool(a.batched))
    for _ in range(a.warmup):
        it()
    torch.cuda.synchronize()
    l0 = be.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = it()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    line = {'metric': 'GAN iteration (G fwd + %d x D step with WGAN-GP + G step), B=%d MSR-VTT-shaped' % (a.num_d, B),
            'ms_per_iteration': ms, 'clips_per_s': B / (ms * 1e-3), 'steps': a.steps, 'warmup': a.warmup,
            'libdlsg_launches_per_iteration': (it.launches if a.graph else (be.launches - l0) // a.steps),
            'cap_loss': float(out[0]), 'loss_G': float(out[1]), 'loss_D': float(out[2]), 'wasserstein': float(out[3]),
            'mode': 'cuda_graph' if a.graph else 'eager', 'batched_critic_calls': bool(a.batched)}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    main()
