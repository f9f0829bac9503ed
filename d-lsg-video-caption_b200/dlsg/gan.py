"""One full iteration of the live GAN trainer (run_gun.py:147-234 + train_disc :339-398) on the drop-in modules,
optionally captured as ONE CUDA graph.

  G forward #1 (fake sample)                                         run_gun.py:167
  num_D x [ D(real one-hot), D(fake logits), D(mixed), WGAN-GP double backward, Adam(D) ]     :339-383
  G forward #2 -> packed CE -> D(raw logits) -> total.backward() -> Adam(G)                  :183-234

The reference's own loop (run_gun.py, unchanged) runs on the same modules eagerly; this entry exists because an
iteration is ~10^4 small kernels and eager Python launches them at ~20 us each.  As with GraphedTrainStep the
teacher-forcing coin flips (layer.py:432) and the dropout seeds are frozen at capture time; the WGAN-GP epsilon
(torch.rand, run_gun.py:355) is re-drawn on every replay (graph-safe CUDA RNG).
"""
import torch

from . import functional as DF
from . import generic as GN
from . import linalg as la
from . import losses


class GanIteration:
    def __init__(self, G, D, opt_g, opt_d, frames, regions, captions, cap_lens, max_words=26, tf_ratio=0.6, num_d=5,
                 gan_lambda=0.01, process_group=None, graph=True, warmup=2, batched=True, overlap_g=True):
        dev = frames.device
        self.G, self.D, self.opt_g, self.opt_d = G, D, opt_g, opt_d
        self.frames, self.regions, self.captions = frames.clone(), regions.clone(), captions.clone()
        self.lens = torch.as_tensor(list(cap_lens), dtype=torch.int32, device=dev)
        self.inv = torch.tensor([1.0 / max(1, sum(min(int(c), max_words) for c in cap_lens))], dtype=torch.float32, device=dev)
        self.lam = torch.tensor(float(gan_lambda), dtype=torch.float32, device=dev)
        self.max_words, self.tf, self.num_d, self.batched = max_words, tf_ratio, num_d, batched
        self.V = D.conv1d.weight.shape[1]
        self.pg, self.world, self.sync = process_group, 1, None
        if process_group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.world = dist.get_world_size(process_group)
            self.sync = DF.GradSync(process_group)
            self.sync_d = DF.GradSync(process_group)
        self.d_params = [p for p in D.parameters() if p.requires_grad]
        # The generator's second forward (run_gun.py:183) does not depend on the critic steps (they update D only): it runs on
        # a side stream next to them - the critic steps are ~10^3 small dependent launches that leave most SMs idle - and
        # joins before D scores its output.  (Same program order on the host; only the stream differs.)
        self.side = torch.cuda.Stream() if (overlap_g and dev.type == 'cuda') else None
        self.graph = None
        self.out = None
        if graph:
            from . import graphs as GR
            GR.check_capturable(opt_g)
            GR.check_capturable(opt_d)
            snap = GR.snapshot([G, D], [opt_g, opt_d]) if warmup > 0 else None
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    self._body()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            if snap is not None:
                GR.restore(snap)                  # warm-up iterations leave no trace in weights, Adam state or RNG streams
                la.new_param_epoch()
            GR.ensure_adam_state(opt_g, [p for p in G.parameters()])
            GR.ensure_adam_state(opt_d, self.d_params)
            self.opt_g.zero_grad(set_to_none=True)
            self.opt_d.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            from . import ops
            l0 = ops.backend().launches
            DF.WC.force = True
            try:
                with torch.cuda.graph(self.graph):
                    self.out = self._body()
            finally:
                DF.WC.force = False
            self.launches = ops.backend().launches - l0
            torch.cuda.synchronize()

    # ------------------------------------------------------------------ run_gun.py:339-383
    def _disc_steps(self, real, fake, obj, mot, att_mask, alpha):
        """real: (B,L,V) one-hot (literal form) or the (B,L) token ids (token form, batched=True)."""
        D, B = self.D, fake.shape[0]
        loss_d = wass = torch.zeros((), device=fake.device)
        if self.batched:
            obj3, mot3, mask3, alpha3 = obj.repeat(3, 1, 1), mot.repeat(3, 1, 1), att_mask.repeat(3, 1, 1), alpha.repeat(3, 1, 1)
        for _ in range(self.num_d):
            self.opt_d.zero_grad(set_to_none=True)
            la.begin_pool(fake.device)                # one zero-filled slab for this critic step's ~70 small gradient accumulators
            eps = torch.rand(B, 1, 1, device=fake.device, requires_grad=not self.batched)
            if self.batched:
                # Token form of the three critic calls (same function of the parameters, re-associated):
                #  * D's first layer is linear, so the embedding of the ONE-HOT real caption is a column gather of its
                #    weight (no (B,L,V) one-hot, no V-wide GEMM), and the embedding of the interpolate
                #    eps * real + (1 - eps) * fake is the same mix of the two embeddings (the bias passes through: the
                #    coefficients sum to 1) - only the fake logits pay the (B*L, V) x (V, 512) product;
                #  * the penalty's gradient wrt the V-wide interpolate is g_tok @ W (W = that layer's (512,V) weight), so its
                #    squared norm per sample is sum_t g_t (W W^T) g_t^T: a 512 x 512 Gram matrix once per critic step instead
                #    of a (B*L, V) gradient and its double backward.
                # One stacked D forward over [real | fake | mixed] tokens (per-group batch means inside DiscV2).
                L = fake.shape[1]
                W, bias = D.conv1d.weight[:, :, 0], D.conv1d.bias
                tok_real = (GN.column_gather(W, real) + bias).view(B, L, -1)
                tok_fake = GN.linear(fake, W, bias)
                tok_mixed = GN.lerp_rows(tok_real, tok_fake, eps)
                logits = D(None, obj3, mot3, mask3, alpha3, _groups=3, _tokens=torch.cat([tok_real, tok_fake, tok_mixed], 0))
                r_logit, f_logit, m_logit = logits[:B], logits[B:2 * B], logits[2 * B:]
                with GN.input_grads_only():           # tok_mixed is a non-leaf: weight / bias gradients of this pass are not computed
                    g_tok = torch.autograd.grad(inputs=tok_mixed, outputs=m_logit, grad_outputs=torch.ones_like(m_logit),
                                                create_graph=True, retain_graph=True)[0]
                g2 = g_tok.reshape(B * L, -1)
                gram = GN.bmm_nt(W, W)                                  # (512,512) = W W^T
                gn = (GN.bmm_nt(g2, gram) * g2).view(B, -1).sum(1).sqrt()
            else:
                mixed = real * eps + fake * (1 - eps)
                r_logit = D(real, obj, mot, att_mask, alpha)
                f_logit = D(fake, obj, mot, att_mask, alpha)
                m_logit = D(mixed, obj, mot, att_mask, alpha)
                g = torch.autograd.grad(inputs=mixed, outputs=m_logit, grad_outputs=torch.ones_like(m_logit),
                                        create_graph=True, retain_graph=True)[0]
                gn = g.contiguous().view(B, -1).norm(2, dim=1)
            gp = ((gn - 1) * (gn - 1)).mean()
            r_loss, f_loss = r_logit.mean(), f_logit.mean()
            loss_d = f_loss - r_loss + 10 * gp
            loss_d.backward()
            la.end_pool()
            if self.world > 1 and 'nodsync' not in self.sync_d.debug:      # (measurement switch: DLSG_SYNC_DEBUG=nodsync)
                self.sync_d.reduce_params(self.d_params)
            self.opt_d.step()
            la.new_param_epoch()                      # critic weights changed: refresh their bf16 copies on next use
            wass = (r_loss - f_loss).detach()
        return loss_d.detach(), wass

    def _body(self):
        la.set_manual_param_epochs(True)              # one weight conversion per optimizer step, not per D forward
        la.new_param_epoch()
        try:
            return self._iteration()
        finally:
            la.set_manual_param_epochs(False)
            la.new_param_epoch()

    def _iteration(self):
        G, D = self.G, self.D
        caps = self.captions
        B, L = caps.shape[0], self.max_words
        seq = (caps[:, :L] > 0).to(torch.float32)
        att_mask = seq.unsqueeze(2) * seq.unsqueeze(1)                                   # run_gun.py:164-166
        with torch.no_grad():                                                            # :167 (detached right after, :170-174)
            f_cap, obj, mot, alpha = G(self.frames, self.regions, caps, L, self.tf)
        if self.batched:
            real = caps[:, :L].contiguous()                                              # token ids: the one-hot is never built
        else:
            real = torch.zeros(B, L, self.V, device=caps.device).scatter_(2, caps[:, :L].unsqueeze(2), 1)  # :449-453
        self.opt_g.zero_grad(set_to_none=True)
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                g_out = G(self.frames, self.regions, caps, L, self.tf)                   # :183, under the critic steps
        loss_d, wass = self._disc_steps(real, f_cap, obj, mot, att_mask, alpha)
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        else:
            g_out = G(self.frames, self.regions, caps, L, self.tf)                       # :183
        out, obj, mot, alpha = g_out
        cap_loss = losses.packed_cross_entropy(out, caps, self.lens, self.inv, unit_grad=True)            # :189-197
        f_logit = D(out, obj.detach(), mot.detach(), att_mask=att_mask, alpha_all=alpha.detach())          # :218
        loss_g = -f_logit.mean()
        total = cap_loss + loss_g * self.lam
        gsync = self.world > 1 and 'nogsync' not in self.sync.debug               # (measurement switch: DLSG_SYNC_DEBUG=nogsync)
        if gsync:
            self.sync.begin_step()
            DF.GRAD_SYNC = self.sync
        try:
            total.backward()
        finally:
            DF.GRAD_SYNC = None
        if gsync:
            self.sync.wait()
            self.sync.write_back([p for p in G.parameters() if p.grad is not None])
        self.opt_g.step()
        return cap_loss.detach(), loss_g.detach(), loss_d, wass

    def load(self, frames, regions, captions, cap_lens=None):
        self.frames.copy_(frames, non_blocking=True)
        self.regions.copy_(regions, non_blocking=True)
        self.captions.copy_(captions, non_blocking=True)
        if cap_lens is not None:
            self.lens.copy_(torch.as_tensor(list(cap_lens), dtype=torch.int32))
            self.inv.fill_(1.0 / max(1, sum(min(int(c), self.max_words) for c in cap_lens)))

    def set_lambda(self, lam):
        self.lam.fill_(float(lam))

    def __call__(self):
        """Returns device scalars (cap_loss, loss_G, loss_D of the last D step, wasserstein of the last D step)."""
        if self.graph is not None:
            self.graph.replay()
            DF.WC.gen += 1    # parameters changed through raw pointers: eval-scope bf16 copies are stale now
            la.new_param_epoch()
            return self.out
        return self._body()
