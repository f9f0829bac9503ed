"""CPU emulation of the dlsg.ops primitives - TEST INFRASTRUCTURE ONLY.

Lets `-m "not gpu"` tests drive the real host orchestration (dlsg/functional.py, models/) on CPU
tensors so that layout / stride / backward-pass logic is checked without a GPU.  Each method
restates the semantics of one C-ABI entry (include/dlsg.h) with plain torch CPU ops and writes
into the caller's output views exactly like the kernels do.  bf16 operands are rounded like the
device path (fp32 accumulate).  Never imported by the product.
"""
import math

import torch


def _f(t):
    return t.float()


class CpuEmulBackend:
    name = 'cpu-emul'

    def __init__(self):
        self.launches = 0

    def gemm(self, a, b, out, bias=None, bias_axis='n', tanh=False, alpha=1.0, accum=False, splitk=1, impl=None, atomic=False,
             b_static=False):
        self.launches += 1
        accum = accum or atomic
        r = torch.matmul(_f(a), _f(b).transpose(-1, -2)) * alpha
        if splitk > 1:
            K = a.shape[-1]
            nkb = (K + 63) // 64
            per = (nkb + splitk - 1) // splitk
            for s in range(splitk):
                k0, k1 = s * per * 64, min(K, (s + 1) * per * 64)
                part = torch.matmul(_f(a[:, k0:k1]), _f(b[:, k0:k1]).t()) * alpha
                if s == 0 and bias is not None:
                    part = part + (bias if bias_axis == 'n' else bias.unsqueeze(-1))
                out[s].copy_(part + (_f(out[s]) if accum else 0))
            return
        if bias is not None:
            r = r + (bias if bias_axis == 'n' else bias.unsqueeze(-1))
        if tanh:
            r = torch.tanh(r)
        if accum:
            r = r + _f(out)
        out.copy_(r)

    def convert(self, src, dst=None, dstT=None):
        self.launches += 1
        if dst is not None:
            dst.copy_(src)
        if dstT is not None:
            dstT.copy_(src.transpose(-1, -2))

    def make_convert_plan(self, pairs, chunk_elems=16384, host=False):
        return {'pairs': list(pairs), 'n': len(pairs)}

    def multi_convert(self, plan):
        self.launches += 1
        for src, src2, dst in plan['pairs']:
            dst.copy_(src if src2 is None else src + src2)

    def make_adam_plan(self, segs, chunk_elems=16384):
        return {'segs': list(segs), 'n': len(segs)}

    def adam_multi(self, plan, step, lr, beta1, beta2, eps, lr_dev=None):
        self.launches += 1
        lr = float(lr_dev) if lr_dev is not None else lr
        for sg in plan['segs']:
            t = float(sg['step']) if sg.get('step') is not None else float(step)      # per-parameter step counters
            bc1, bc2 = 1.0 - beta1 ** t, 1.0 - beta2 ** t
            p_, g_, m_, v_ = sg['p'], sg['g'], sg['m'], sg['v']
            g_ = g_.float()
            m_.copy_(m_ + (g_ - m_) * (1 - beta1))
            v_.copy_(beta2 * v_ + (1 - beta2) * g_ * g_)
            p_.sub_((lr / bc1) * m_ / (v_.sqrt() / (bc2 ** 0.5) + eps))
            if sg.get('dst') is not None:
                sg['dst'].copy_(p_)

    def colsum(self, x, out):
        self.launches += 1
        out.add_(_f(x).reshape(-1, x.shape[-1]).sum(0))

    @staticmethod
    def _ln(t, gamma, beta):
        mean = t.mean(-1, keepdim=True)
        var = ((t - mean) ** 2).mean(-1, keepdim=True)
        rstd = torch.rsqrt(var + 1e-5)
        return (t - mean) * rstd * gamma + beta, mean, rstd

    def norm_fwd(self, x, gamma, beta, y=None, y2=None, res=None, stats=None, pre_tanh=False, post_tanh=False, drop=None):
        self.launches += 1
        assert drop is None or drop[0] == 0, 'dropout is not emulated on CPU'
        t = _f(x)
        if res is not None:
            t = t + _f(res)
        if pre_tanh:
            t = torch.tanh(t)
        o, mean, rstd = self._ln(t, gamma, beta)
        if post_tanh:
            o = torch.tanh(o)
        if stats is not None:
            stats.view(-1, 2).copy_(torch.cat([mean.reshape(-1, 1), rstd.reshape(-1, 1)], 1))
        if y is not None:
            y.copy_(o)
        if y2 is not None:
            y2.copy_(o)

    def norm_bwd(self, dy, x, gamma, beta, stats, dx=None, res=None, dgamma=None, dbeta=None, pre_tanh=False,
                 post_tanh=False, in_is_tanh=False, drop=None, dx_accum=False, dxsum=None):
        self.launches += 1
        assert drop is None or drop[0] == 0
        t = _f(x)
        if res is not None:
            t = t + _f(res)
        if pre_tanh:
            t = torch.tanh(t)
        D = t.shape[-1]
        st = stats.view(*t.shape[:-1], 2)
        mean, rstd = st[..., 0:1], st[..., 1:2]
        xh = (t - mean) * rstd
        g = _f(dy)
        if post_tanh:
            yt = torch.tanh(xh * gamma + beta)
            g = g * (1 - yt * yt)
        if dgamma is not None:
            dgamma.add_((g * xh).reshape(-1, D).sum(0))
            dbeta.add_(g.reshape(-1, D).sum(0))
        d = g * gamma
        s1 = d.mean(-1, keepdim=True)
        s2 = (d * xh).mean(-1, keepdim=True)
        r = rstd * (d - s1 - xh * s2)
        if pre_tanh or in_is_tanh:
            r = r * (1 - t * t)
        if dx is not None:
            if dx_accum:
                r = r + _f(dx)
            dx.copy_(r)
        if dxsum is not None:
            dxsum.add_(r.reshape(-1, D).sum(0))

    def norm_bwd2(self, x, dy, u, gamma, stats, g_dy=None, g_x=None, g_gamma=None):
        """Reference by automatic differentiation of the restated LayerNorm backward (the kernel uses closed forms)."""
        self.launches += 1
        D = x.shape[-1]
        with torch.enable_grad():
            x_ = x.detach().double().reshape(-1, D).requires_grad_(True)
            dy_ = dy.detach().double().reshape(-1, D).requires_grad_(True)
            gm = gamma.detach().double().requires_grad_(True)
            mean = x_.mean(-1, keepdim=True)
            r = torch.rsqrt(((x_ - mean) ** 2).mean(-1, keepdim=True) + 1e-5)
            a = (x_ - mean) * r
            g = dy_ * gm
            dx = r * (g - g.mean(-1, keepdim=True) - a * (g * a).mean(-1, keepdim=True))
            G_ = torch.autograd.grad((dx * u.detach().double().reshape(-1, D)).sum(), [dy_, x_, gm])
        if g_dy is not None:
            g_dy.copy_(G_[0].float().view(g_dy.shape))
        if g_x is not None:
            g_x.copy_(G_[1].float().view(g_x.shape))
        if g_gamma is not None:
            g_gamma.add_(G_[2].float())

    def lstm_cell_fwd(self, gates, c_prev, c_out, h_out=None, row_bias=None, bias=None, h2=None, h3=None, drop=None):
        self.launches += 1
        assert drop is None or drop[0] == 0
        g = gates.sum(0) if gates.dim() == 3 else gates.clone()
        g0 = gates[0] if gates.dim() == 3 else gates
        if row_bias is not None:
            g = g + row_bias
        if bias is not None:
            g = g + bias
        H = g.shape[1] // 4
        i, f, gg, o = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
        c = f * (c_prev if c_prev is not None else 0) + i * gg
        h = o * torch.tanh(c)
        g0.copy_(torch.cat([i, f, gg, o], 1))
        c_out.copy_(c)
        for dst in (h_out, h2, h3):
            if dst is not None:
                dst.copy_(h)

    def lstm_cell_bwd(self, acts, c_prev, c_new, dh, dc_next, dc_prev, dgates=None, dgates2=None, dgatesT=None, drop=None,
                      dh2=None, dc_next2=None, dgates_add=None, dh_total=None):
        self.launches += 1
        assert drop is None or drop[0] == 0
        if dh2 is not None:
            dh = dh + (dh2.sum(0) if dh2.dim() == 3 else dh2)
        if dh_total is not None:
            dh_total.copy_(dh)
        if dc_next2 is not None:
            dc_next = dc_next2 if dc_next is None else dc_next + dc_next2
        H = acts.shape[1] // 4
        i, f, g, o = acts[:, :H], acts[:, H:2 * H], acts[:, 2 * H:3 * H], acts[:, 3 * H:]
        tc = torch.tanh(c_new)
        dc = dh * o * (1 - tc * tc)
        if dc_next is not None:
            dc = dc + dc_next
        cp = c_prev if c_prev is not None else torch.zeros_like(c_new)
        d = torch.cat([dc * g * i * (1 - i), dc * cp * f * (1 - f), dc * i * (1 - g * g), dh * tc * o * (1 - o)], 1)
        if dgates_add is not None:
            d = d + dgates_add
        if dc_prev is not None:
            dc_prev.copy_(dc * f)
        if dgates is not None:
            dgates.copy_(d)
        if dgates2 is not None:
            dgates2.copy_(d)
        if dgatesT is not None:
            dgatesT[:, :d.shape[0]].copy_(d.t())

    def lstm_cell_bwd2(self, acts, c_prev, c_new, dh, dc_next, u, w, g_dh, g_dc, g_pre, g_cprev, u2=None, g_dh2=None):
        """Reference by automatic differentiation of the restated cell backward (the kernel uses closed forms)."""
        self.launches += 1
        if u2 is not None:
            u2s = u2.sum(0) if u2.dim() == 3 else u2
            u = u2s if u is None else u + u2s
        with torch.enable_grad():
            H = acts.shape[1] // 4
            a = acts.double()
            # recover pre-activations from the saved activations (sigmoid / tanh are invertible on the open interval)
            eps = 1e-12
            sig_inv = lambda y: torch.log(y.clamp(eps, 1 - eps)) - torch.log1p(-y.clamp(eps, 1 - eps))
            pre = torch.cat([sig_inv(a[:, :H]), sig_inv(a[:, H:2 * H]), torch.atanh(a[:, 2 * H:3 * H].clamp(-1 + eps, 1 - eps)),
                             sig_inv(a[:, 3 * H:])], 1).requires_grad_(True)
            c0 = (c_prev.double() if c_prev is not None else torch.zeros_like(c_new, dtype=torch.float64)).requires_grad_(True)
            dh_ = dh.double().requires_grad_(True)
            dc_ = (dc_next.double() if dc_next is not None else torch.zeros_like(c_new, dtype=torch.float64)).requires_grad_(True)
            i, f, g, o = torch.sigmoid(pre[:, :H]), torch.sigmoid(pre[:, H:2 * H]), torch.tanh(pre[:, 2 * H:3 * H]), torch.sigmoid(pre[:, 3 * H:])
            c = f * c0 + i * g
            tc = torch.tanh(c)
            D = dh_ * o * (1 - tc * tc) + dc_
            dpre = torch.cat([D * g * i * (1 - i), D * c0 * f * (1 - f), D * i * (1 - g * g), dh_ * tc * o * (1 - o)], 1)
            dc0 = D * f
            L2 = 0
            if u is not None:
                L2 = L2 + (dpre * u.double()).sum()
            if w is not None:
                L2 = L2 + (dc0 * w.double()).sum()
            if not torch.is_tensor(L2):
                grads = [torch.zeros_like(dh_), torch.zeros_like(dc_), torch.zeros_like(pre), torch.zeros_like(c0)]
            else:
                grads = torch.autograd.grad(L2, [dh_, dc_, pre, c0], allow_unused=True)
                grads = [torch.zeros_like(x) if g_ is None else g_ for g_, x in zip(grads, [dh_, dc_, pre, c0])]
        for dst, g_ in zip((g_dh, g_dc, g_pre, g_cprev), grads):
            if dst is not None:
                dst.copy_(g_.float())
        if g_dh2 is not None:
            g_dh2.copy_(grads[0].float())

    @staticmethod
    def fused_step_supported(H):
        return H % 4 == 0 and H <= 2048

    def lstm_cell_norm_fwd(self, gates, c_prev, c_out, gamma, beta, y, h_out=None, row_bias=None, bias=None, h2=None, h3=None,
                           drop=None, y2=None, stats=None, post_tanh=False, ydrop=None):
        B, H = c_out.shape
        h = torch.empty(B, H)
        self.lstm_cell_fwd(gates, c_prev, c_out, h_out=h, row_bias=row_bias, bias=bias, h2=h2, h3=h3, drop=drop)
        if h_out is not None:
            h_out.copy_(h)
        self.norm_fwd(h, gamma, beta, y=y, y2=y2, stats=stats, post_tanh=post_tanh, drop=ydrop)
        self.launches -= 1

    def cell_norm_attn2_fwd(self, cell, attn):
        """The two stages back to back (the kernel fuses them into one launch); same shape limits as the kernel."""
        g = cell['gates']
        H = g.shape[-1] // 4
        KW, VW = attn['KW'], attn['VW']
        if not (H % 4 == 0 and H <= 1024 and H == KW.shape[3] and VW.shape[3] <= 1024 and KW.shape[0] <= 2 and KW.shape[2] <= 8
                and (g.shape[0] if g.dim() == 3 else 1) <= 4):
            return False
        c = dict(cell)
        self.lstm_cell_norm_fwd(c.pop('gates'), c.pop('c_prev'), c.pop('c_out'), c.pop('gamma'), c.pop('beta'), c.pop('y'), **c)
        self.attn2_fwd(**attn)
        self.launches -= 1
        return True

    def norm_lstm_cell_bwd(self, acts, c_prev, c_new, dc_next, dc_prev, dy, x, gamma, beta, stats, dgamma, dbeta, dh=None, dh2=None,
                           dgates=None, dgates2=None, dgatesT=None, dgates_sum=None, drop=None, post_tanh=False, ydrop=None):
        B, H = c_new.shape
        dx = torch.zeros(B, H)
        self.norm_bwd(dy, x, gamma, beta, stats, dx=dx, post_tanh=post_tanh, drop=ydrop)
        # per-row LayerNorm parameter-gradient contributions
        st = stats.view(B, 2)
        xn = (x - st[:, 0:1]) * st[:, 1:2]
        g = dy
        if post_tanh:
            yt = torch.tanh(xn * gamma + beta)
            g = g * (1 - yt * yt)
        dgamma.copy_(g * xn)
        dbeta.copy_(g)
        if dh is not None:
            dx = dx + dh
        d = torch.empty(B, 4 * H)
        self.lstm_cell_bwd(acts, c_prev, c_new, dx, dc_next, dc_prev, dgates=d, dgates2=dgates2, dgatesT=dgatesT, drop=drop, dh2=dh2)
        if dgates is not None:
            dgates.copy_(d)
        if dgates_sum is not None:
            dgates_sum.add_(d)
        self.launches -= 1

    @staticmethod
    def _sm(x, dim, scale, mask, mask_mode):
        v = x * scale
        if mask_mode == 1:
            v = torch.where(mask > 0, v, torch.full_like(v, -9e15))
        return torch.softmax(v, dim)

    def softmax_fwd(self, x, y, dim, scale=1.0, mask=None, mask_mode=0):
        self.launches += 1
        s = self._sm(x, dim, scale, mask, mask_mode)
        if mask_mode == 2:
            s = torch.where(mask > 0, s, torch.zeros_like(s))
        y.copy_(s)

    def softmax_bwd(self, x, dy, dx, dim, scale=1.0, mask=None, mask_mode=0):
        self.launches += 1
        s = self._sm(x, dim, scale, mask, mask_mode)
        g = dy
        if mask_mode == 2:
            g = torch.where(mask > 0, g, torch.zeros_like(g))
        r = scale * s * (g - (g * s).sum(dim, keepdim=True))
        if mask_mode == 1:
            r = torch.where(mask > 0, r, torch.zeros_like(r))
        dx.copy_(r)

    def softmax_bwd2(self, x, dy, u, dim, g_dy=None, g_x=None, scale=1.0, mask=None, mask_mode=0):
        """Reference by automatic differentiation of the restated softmax backward (the kernel uses closed forms)."""
        self.launches += 1
        with torch.enable_grad():
            x_ = x.detach().double().requires_grad_(True)
            dy_ = dy.detach().double().requires_grad_(True)
            s = self._sm(x_, dim, scale, mask, mask_mode)
            g = dy_ if mask_mode != 2 else torch.where(mask > 0, dy_, torch.zeros_like(dy_))
            r = scale * s * (g - (g * s).sum(dim, keepdim=True))
            if mask_mode == 1:
                r = torch.where(mask > 0, r, torch.zeros_like(r))
            a, b = torch.autograd.grad((r * u.double()).sum(), [dy_, x_])
        if g_dy is not None:
            g_dy.copy_(a.float())
        if g_x is not None:
            g_x.copy_(b.float())

    def ew(self, op, ins, outs, cols=0):
        from dlsg import _lib as L
        self.launches += 1
        n = ins[0].numel()
        v = [t.reshape(-1) for t in ins]
        if op == L.EW_TANH_BWD:
            res = [v[0] * (1 - v[1] * v[1])]
        elif op == L.EW_TANH_BWD2:
            res = [v[2] * (1 - v[1] * v[1]), -2 * v[1] * v[0] * v[2]]
        elif op == L.EW_MUL_BWD:
            res = [v[0] * v[2], v[0] * v[1]]
        elif op == L.EW_MUL_BWD2:
            res = [v[3] * v[2] + v[4] * v[1], v[4] * v[0], v[3] * v[0]]
        elif op == L.EW_LERP_ROWS:
            e = v[2].repeat_interleave(cols)
            res = [v[0] * e + v[1] * (1 - e)]
        elif op == L.EW_LERP_ROWS_BWD:
            e = v[1].repeat_interleave(cols)
            res = [v[0] * e, v[0] * (1 - e)]
        else:
            raise ValueError(op)
        for dst, r in zip(outs, res):
            if dst is not None:
                dst.copy_(r.view(dst.shape))

    def node_attn_fwd(self, Kp, Vp, qp, alpha, ctx, rows_per_node=1):
        self.launches += 1
        nh, nodes, P, H = Kp.shape
        rows = qp.shape[0]
        idx = torch.arange(rows) // rows_per_node
        q = qp.view(rows, nh, H)
        for h in range(nh):
            K, V = Kp[h][idx], Vp[h][idx]                        # (rows,P,H)
            lg = torch.einsum('rph,rh->rp', K, q[:, h]) / math.sqrt(H)
            a = torch.softmax(lg, 1)
            if alpha is not None:
                alpha[:, h * P:(h + 1) * P].copy_(a)
            ctx[:, h * H:(h + 1) * H].copy_(torch.einsum('rp,rph->rh', a, V))

    def node_attn_bwd(self, Kp, Vp, qp, alpha, dctx, dqp, dKp, dVp, dalpha_ext=None):
        self.launches += 1
        nh, nodes, P, H = Kp.shape
        rows = qp.shape[0]
        q = qp.view(rows, nh, H)
        sc = 1.0 / math.sqrt(H)
        for h in range(nh):
            a = alpha[:, h * P:(h + 1) * P]
            dc = dctx[:, h * H:(h + 1) * H]
            da = torch.einsum('rph,rh->rp', Vp[h], dc)
            if dalpha_ext is not None:
                da = da + dalpha_ext[:, h * P:(h + 1) * P]
            dl = a * (da - (a * da).sum(1, keepdim=True)) * sc
            dqp.view(rows, nh, H)[:, h].copy_(torch.einsum('rp,rph->rh', dl, Kp[h]))
            dKp[h].add_(torch.einsum('rp,rh->rph', dl, q[:, h]))
            dVp[h].add_(torch.einsum('rp,rh->rph', a, dc))

    @staticmethod
    def attn2_supported(nh, P, Hk, Hv):
        return 1 <= nh <= 2 and 1 <= P <= 8 and Hk <= 1024 and Hv <= 1024 and Hk % 4 == 0 and Hv % 4 == 0

    def attn2_fwd(self, KW, VW, q, alpha, co, scale, rows_per_node=1, ln=None):
        self.launches += 1
        nh, nodes, P, Hk = KW.shape
        Hv = VW.shape[3]
        idx = torch.arange(q.shape[0]) // rows_per_node
        for h in range(nh):
            lg = torch.einsum('rpk,rk->rp', KW[h][idx], q) * scale
            a = torch.softmax(lg, 1)
            if alpha is not None:
                alpha[:, h * P:(h + 1) * P].copy_(a)
            co[:, h * Hv:(h + 1) * Hv].copy_(torch.einsum('rp,rpv->rv', a, VW[h][idx]))
        if ln is not None:
            for h in range(nh):
                d = ln.get('drop')
                d = None if d is None else (d[0], d[1], d[2] + h * ln['drop_head_stride'])
                self.norm_fwd(co[:, h * Hv:(h + 1) * Hv], ln['gamma'][h], ln['beta'][h], y=ln['y'][:, h * Hv:(h + 1) * Hv],
                              stats=ln['stats'][h], pre_tanh=True, drop=d)
                self.launches -= 1

    def attn2_bwd(self, KW, VW, q, alpha, dco, dq, dKW, dVW, scale, dalpha_ext=None, ln=None, save=None):
        self.launches += 1
        nh, nodes, P, Hk = KW.shape
        Hv = VW.shape[3]
        if ln is not None:
            dco = torch.empty_like(ln['co'])
            for h in range(nh):
                sl = slice(h * Hv, (h + 1) * Hv)
                d = ln.get('drop')
                d = None if d is None else (d[0], d[1], d[2] + h * ln['drop_head_stride'])
                dg, db = torch.zeros(Hv), torch.zeros(Hv)
                # per-row parameter-gradient contributions: run the row-wise reference one row at a time
                x, dy, st = ln['co'][:, sl], ln['dy'][:, sl], ln['stats'][h]
                t = torch.tanh(x)
                xh = (t - st[:, :1]) * st[:, 1:2]
                assert d is None or d[0] == 0, 'dropout is not emulated on CPU'
                dyd = dy
                ln['dgamma_rows'][:, sl].copy_(dyd * xh)
                ln['dbeta_rows'][:, sl].copy_(dyd)
                self.norm_bwd(dy, x, ln['gamma'][h], None, st, dx=dco[:, sl], dgamma=dg, dbeta=db, pre_tanh=True, drop=d)
                self.launches -= 1
        for h in range(nh):
            a = alpha[:, h * P:(h + 1) * P]
            dc = dco[:, h * Hv:(h + 1) * Hv]
            da = torch.einsum('rpv,rv->rp', VW[h], dc)
            if dalpha_ext is not None:
                da = da + dalpha_ext[:, h * P:(h + 1) * P]
            dl = a * (da - (a * da).sum(1, keepdim=True)) * scale
            dq.add_(torch.einsum('rp,rpk->rk', dl, KW[h]))
            if save is not None:                      # deferred node gradients: record, attn2_bwd_nodes accumulates
                save[0][:, h * P:(h + 1) * P].copy_(dl)
                save[1][:, h * Hv:(h + 1) * Hv].copy_(dc)
                continue
            dKW[h].add_(torch.einsum('rp,rk->rpk', dl, q))
            dVW[h].add_(torch.einsum('rp,rv->rpv', a, dc))

    def attn2_bwd_nodes(self, q_all, dl_all, alpha_all, dco_all, dKW, dVW, accumulate=False):
        self.launches += 1
        nh, rows, P, Hk = dKW.shape
        Hv = dVW.shape[3]
        for h in range(nh):
            k = torch.einsum('trp,trk->rpk', dl_all[:, :, h * P:(h + 1) * P], q_all)
            v = torch.einsum('trp,trv->rpv', alpha_all[:, :, h * P:(h + 1) * P], dco_all[:, :, h * Hv:(h + 1) * Hv])
            if accumulate:
                dKW[h].add_(k)
                dVW[h].add_(v)
            else:
                dKW[h].copy_(k)
                dVW[h].copy_(v)

    # ---- one LSTM step in one launch (csrc/lstm_step.cu)
    @staticmethod
    def lstm_step_supported(B, H):
        return 1 <= B <= 64 and H >= 128 and H % 128 == 0

    def lstm_step_fwd(self, W, h_in, gin, c_in, c_out, acts, h_out=None, h_op=None):
        self.launches += 1
        for d in range(len(W)):
            H = W[d].shape[1]
            pre = _f(gin[d])
            if h_in is not None:
                pre = pre + _f(h_in[d]) @ _f(W[d]).t()
            i, f, g, o = pre[:, :H], pre[:, H:2 * H], pre[:, 2 * H:3 * H], pre[:, 3 * H:]
            i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
            c = f * (c_in[d] if (c_in is not None and c_in[d] is not None) else 0) + i * g
            h = o * torch.tanh(c)
            acts[d].copy_(torch.cat([i, f, g, o], 1))
            c_out[d].copy_(c)
            if h_out is not None:
                h_out[d].copy_(h)
            if h_op is not None:
                h_op[d].copy_(h)

    # ---- fused region -> frame aggregation (csrc/region_agg.cu): plain composition of the same algebra
    @staticmethod
    def region_aggregate_supported(T, TR, H, dtype):
        return dtype == torch.bfloat16 and H == 1024 and 1 <= T <= 26

    @staticmethod
    def _bfr(x):
        return x.to(torch.bfloat16).float()

    def region_aggregate_fwd(self, Y, F, gamma, beta, scale, T, agg=None, U=None, stats=None, St=None, tconst=None, scores_only=False):
        self.launches += 1
        for e in range(len(Y)):
            y = _f(Y[e])
            H = y.shape[1]
            B = F[e].shape[0] // T
            TR = y.shape[0] // B
            mu = y.mean(1, keepdim=True)
            var = ((y * y).mean(1, keepdim=True) - mu * mu).clamp_min(0)
            rstd = torch.rsqrt(var + 1e-5)
            xh = ((y - mu) * rstd).view(B, TR, H)
            f = _f(F[e])
            fg = self._bfr(f * gamma[e])                                   # the kernel's bf16 A operand
            S = fg.view(B, T, H) @ xh.transpose(1, 2) + (f @ beta[e]).view(B, T, 1)
            if stats is not None:
                stats[e].copy_(torch.cat([mu, rstd], 1))
            if St is not None:
                St[e].copy_(S)
            tc = torch.zeros(B * T, 4)
            tc[:, 0] = fg.sum(1)
            tc[:, 1] = f @ beta[e]
            if not scores_only:
                m = (S * scale).max(-1).values
                tc[:, 2] = m.reshape(-1)
                tc[:, 3] = (1.0 / torch.exp(S * scale - m.unsqueeze(-1)).sum(-1)).reshape(-1)
            if tconst is not None:
                tconst[e].copy_(tc)
            if scores_only:
                continue
            A = torch.softmax(S * scale, dim=2)
            u = (A @ xh).view(B * T, H)
            if U is not None:
                U[e].copy_(u)
            agg[e].copy_(u * gamma[e] + beta[e])

    @staticmethod
    def region_aggregate_bwd_workspace(B, T, TR):
        return 16

    def region_aggregate_bwd(self, Y, stats, St, dSm, F, dA, U, tcF, tcA, gamma, beta, scale, T, dpre, dF, dgamma, dbeta, dbias=None,
                             work=None):
        self.launches += 2
        for e in range(len(Y)):
            y = _f(Y[e])
            H = y.shape[1]
            B = F[e].shape[0] // T
            TR = y.shape[0] // B
            mu, rstd = stats[e][:, 0:1], stats[e][:, 1:2]
            xh = ((y - mu) * rstd).view(B, TR, H)
            A = torch.exp(St[e] * scale - tcF[e][:, 2].view(B, T, 1)) * tcF[e][:, 3].view(B, T, 1)
            c = (A * dSm[e]).sum(-1, keepdim=True)
            dS = scale * A * (dSm[e] - c)                                  # softmax backward
            dAg = self._bfr(_f(dA[e]) * gamma[e]).view(B, T, H)
            Fg = self._bfr(_f(F[e]) * gamma[e]).view(B, T, H)
            dxh = self._bfr(A).transpose(1, 2) @ dAg + self._bfr(dS).transpose(1, 2) @ Fg        # (B,TR,H)
            a = dxh.mean(-1, keepdim=True)
            b = (dxh * xh).mean(-1, keepdim=True)
            dp = rstd.view(B, TR, 1) * (dxh - a - xh * b) * (1 - y.view(B, TR, H) ** 2)
            dpre[e].copy_(dp.view(B * TR, H))
            V = dS @ xh                                                    # (B,T,H)
            dF[e].copy_(_f(dA[e]) + (V * gamma[e]).view(B * T, H))
            dgamma[e].add_((_f(dA[e]) * U[e] + _f(F[e]) * V.view(B * T, H)).sum(0))
            dbeta[e].add_(_f(dA[e]).sum(0))
            if dbias is not None:
                dbias[e].add_(dp.view(B * TR, H).sum(0))

    @staticmethod
    def latent_psl_supported(T, P, H):
        return P <= 8 and T <= 32 and H % 4 == 0

    def latent_psl_fwd(self, X, theta, Gs, N):
        self.launches += 1
        g = torch.softmax(X @ theta.t(), dim=1)
        Gs.copy_(g)
        N.copy_(g.transpose(1, 2) @ X)

    def latent_psl_fwd_multi(self, X, theta, Gs, N):
        for e in range(len(X)):
            self.latent_psl_fwd(X[e], theta[e], Gs[e], N[e])
        self.launches -= len(X) - 1

    def latent_psl_bwd_multi(self, X, theta, Gs, dN, dX, dtheta):
        for e in range(len(X)):
            self.latent_psl_bwd(X[e], theta[e], Gs[e], dN[e], dX[e], dtheta[e])
        self.launches -= len(X) - 1

    def latent_psl_bwd(self, X, theta, Gs, dN, dX, dtheta):
        self.launches += 1
        dGs = X @ dN.transpose(1, 2)
        dG = Gs * (dGs - (Gs * dGs).sum(1, keepdim=True))
        dX.copy_(Gs @ dN + dG @ theta)
        dtheta.add_((dG.transpose(1, 2) @ X).sum(0))

    def embedding_gather(self, table, ids, out=None, out2=None, drop=None):
        self.launches += 1
        assert drop is None or drop[0] == 0
        v = table[ids]
        for dst in (out, out2):
            if dst is not None:
                dst.copy_(v)

    def embedding_scatter_add(self, dtable, ids, dout, drop=None):
        self.launches += 1
        dtable.index_add_(0, ids, _f(dout))

    def mean_nodes_fwd(self, x, y):
        self.launches += 1
        y.copy_(x.mean(1))

    def mean_nodes_bwd(self, dy, dx):
        self.launches += 1
        dx.add_(dy.unsqueeze(1) / dx.shape[1])

    def axpby(self, x, a, y, b):
        self.launches += 1
        y.copy_(a * x.reshape(y.shape) + (b * y if b != 0 else 0))

    def dropout(self, x, y, drop):
        raise AssertionError('dropout is not emulated on CPU')

    def add_rowbcast(self, x, pe, y, drop=None):
        assert drop is None or drop[0] == 0
        self.launches += 1
        y.copy_(x + pe.reshape(-1).repeat(x.numel() // pe.numel()).view(x.shape))

    def relu_(self, x):
        self.launches += 1
        x.clamp_(min=0)

    def relu_bwd(self, r, dr, dx):
        self.launches += 1
        dx.copy_(torch.where(r > 0, dr, torch.zeros_like(dr)))

    def mul(self, a, b, y):
        self.launches += 1
        y.copy_(a * b)

    def row_argmax(self, logits, ids):
        self.launches += 1
        ids.copy_(logits.max(1)[1])

    def log_softmax(self, logits, out):
        self.launches += 1
        out.copy_(torch.log_softmax(logits, 1))

    def ce_masked(self, logits, targets, lens, loss_sum, dlogits, inv_count, inv_count_dev=None):
        self.launches += 1
        if inv_count_dev is not None:
            inv_count = float(inv_count_dev)
        B, Lw, V = logits.shape
        lp = torch.log_softmax(logits, -1)
        m = (torch.arange(Lw).unsqueeze(0) < lens.unsqueeze(1)).float()
        nll = -lp.gather(2, targets.unsqueeze(2)).squeeze(2)
        loss_sum.add_((nll * m).sum() * inv_count)
        if dlogits is not None:
            g = torch.exp(lp)
            g.scatter_add_(2, targets.unsqueeze(2), -torch.ones(B, Lw, 1))
            dlogits.copy_(g * m.unsqueeze(2) * inv_count)

    def beam_topk(self, logits, last, end_index, k, top_lp, top_id, normalize=True):
        self.launches += 1
        lp = torch.log_softmax(logits, 1) if normalize else logits
        if last is not None:
            forced = torch.full_like(lp, float('-inf'))
            forced[:, end_index] = 0
            lp = torch.where((last == end_index).unsqueeze(1), forced, lp)
        v, i = self._topk_stable(lp, k)
        top_lp.copy_(v)
        top_id.copy_(i)

    @staticmethod
    def _topk_stable(x, k):
        # value desc, lowest index first on ties (the device kernels' tie rule)
        idx = torch.argsort(x, dim=1, descending=True, stable=True)[:, :k]
        return x.gather(1, idx), idx

    def beam_merge(self, top_lp, top_id, last_lp, B, beam, k, new_lp, new_cls, backptr, all_end, end_index):
        self.launches += 1
        s = (top_lp.view(B, beam, k) + last_lp.view(B, beam, 1)).reshape(B, beam * k)
        v, i = self._topk_stable(s, beam)
        cls = top_id.view(B, beam * k).gather(1, i)
        new_lp.copy_(v)
        new_cls.copy_(cls)
        backptr.copy_(i // k)
        if all_end is not None and bool((cls != end_index).any()):
            all_end.zero_()

    def beam_gather(self, src, dst, backptr, B, beam):
        self.launches += 1
        idx = (torch.arange(B).unsqueeze(1) * beam + backptr.view(B, beam)).reshape(-1)
        dst.copy_(src[idx])

    def beam_gather_multi(self, pairs, backptr, B, beam):
        self.launches += 1
        idx = (torch.arange(B).unsqueeze(1) * beam + backptr.view(B, beam)).reshape(-1)
        for src, dst in pairs:
            dst.copy_(src[idx])

    def beam_backtrack(self, preds, backs, S, B, beam, out):
        self.launches += 1
        cur = torch.arange(beam).unsqueeze(0).expand(B, beam).clone()
        for t in range(S - 1, -1, -1):
            out[:, :, t] = preds[t].gather(1, cur)
            if t > 0:
                cur = backs[t - 1].gather(1, cur)
