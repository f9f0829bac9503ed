"""The 26-step two-LSTM decoder with node attention (models/layer.py:394-602) over libdlsg kernels.

Hoisted out of the time loop (the reference recomputes them every step):
  * K/V projections of the latent nodes (sublayer.py:29,31) - once per sequence;
  * the global-feature + bias contribution to the query-LSTM gates (layer.py:571) - once per sequence;
  * the vocabulary projection (layer.py:600) - one (B*T, Hd) x (V, Hd)^T GEMM after the loop when teacher
    forced, a per-step GEMM only on the steps whose arg-max is fed back.
Per step the two LSTM gate GEMMs read one concatenated operand row each:
  Xq[t] = [lang_h(t-1) | word(t) | query_h(t-1)]      Xl[t] = [ctx_obj | ctx_motion | q | lang_h(t-1)]
and every kernel writes its result straight into the slot of the operand row that consumes it.
"""
import math
import os

import torch

from . import linalg as la
from . import ops
from .functional import WC, next_seed, site, _c
from .linalg import empty, zeros, small_zeros, op, op_empty, op_zeros, ceil8

START = 1


# the decode step's cell + LayerNorm + attention as one launch when the shape allows (measurement switch)
FUSED_CELL_ATTN = os.environ.get('DLSG_FUSED_CELL_ATTN', '1') != '0'
# ... up to this many rows: one CTA per row runs the two stages one after the other, which wins at the training batch (64 rows:
# 6.230 -> 6.195 ms per step) and loses once the rows alone fill the GPU (beam-5 at 640 rows: 19.9 k -> 19.0 k captions/s)
FUSED_CELL_ATTN_MAX_ROWS = int(os.environ.get('DLSG_FUSED_CELL_ATTN_MAX_ROWS', '128'))
splitk_for = la.splitk_rows          # (= la.splitk_for up to 64 rows; explicit partials for the 65..~300-row decode batches too)


flat2 = la.flat2


class DecoderCore:
    """Shapes, packed weights and the single-step launcher shared by training, greedy and beam decoding."""

    def __init__(self, t, pf, multi_modal, p_drop):
        self.t, self.pf, self.p = t, pf, p_drop
        g = lambda n: t[pf + n]
        self.nh = 2 if multi_modal else 1
        self.Hq = g('query_lstm.weight_hh').shape[1]
        self.Hd = g('lang_lstm.weight_hh').shape[1]
        self.V, self.W = g('word_embed.weight').shape
        self.H = g('context_att.K.weight').shape[0]
        self.GH = g('query_lstm.weight_ih').shape[1] - self.Hd - self.W
        nh, H, Hq, Hd, W = self.nh, self.H, self.Hq, self.Hd, self.W
        # operand-row layouts (segment offsets are multiples of 8 elements = 16 B in bf16)
        self.oW = ceil8(Hd)
        self.oQ = self.oW + ceil8(W)
        self.Kq = self.oQ + Hq
        self.oq = nh * H
        self.ol = self.oq + ceil8(Hq)
        self.Kl = self.ol + Hd
        be = ops.backend()
        self.fused = be.fused_step_supported(Hq) and be.fused_step_supported(Hd)   # cell+LayerNorm fused kernels
        self._pack()

    def _pack(self):
        be = ops.backend()
        t, pf = self.t, self.pf
        g = lambda n: t[pf + n].detach()
        nh, H, Hq, Hd, W, GH = self.nh, self.H, self.Hq, self.Hd, self.W, self.GH
        heads = ['context_att', 'context_att_2'][:nh]
        names = (['query_lstm.weight_ih', 'query_lstm.weight_hh', 'query_lstm.bias_ih', 'query_lstm.bias_hh',
                  'lang_lstm.weight_ih', 'lang_lstm.weight_hh', 'lang_lstm.bias_ih', 'lang_lstm.bias_hh'] +
                 ['%s.%s' % (h, n) for h in heads for n in ('Q.weight', 'output_layer.0.weight')])
        ps = [t[pf + n] for n in names]

        def build():
            ref = g('query_lstm.weight_ih')
            wih, whh = g('query_lstm.weight_ih'), g('query_lstm.weight_hh')
            Wq = op_zeros((4 * Hq,), self.Kq, ref)
            WC.cv(wih[:, :Hd], Wq[:, :Hd])
            WC.cv(wih[:, Hd + GH:], Wq[:, self.oW:self.oW + W])
            WC.cv(whh, Wq[:, self.oQ:])
            Wg = op_empty((4 * Hq,), GH, ref)
            WC.cv(wih[:, Hd:Hd + GH], Wg)
            bq = empty((4 * Hq,), ref)
            WC.sum2(g('query_lstm.bias_ih'), g('query_lstm.bias_hh'), bq)
            lih, lhh = g('lang_lstm.weight_ih'), g('lang_lstm.weight_hh')
            Wl = op_zeros((4 * Hd,), self.Kl, ref)
            WC.cv(lih[:, :nh * H], Wl[:, :nh * H])
            WC.cv(lih[:, nh * H:], Wl[:, self.oq:self.oq + Hq])
            WC.cv(lhh, Wl[:, self.ol:])
            bl = empty((4 * Hd,), ref)
            WC.sum2(g('lang_lstm.bias_ih'), g('lang_lstm.bias_hh'), bl)
            Wqp = op_empty((nh * H,), Hq, ref)
            Wo = op_empty((nh, H), H, ref)
            for i, h in enumerate(heads):
                WC.cv(g(h + '.Q.weight'), Wqp[i * H:(i + 1) * H])
                WC.cv(g(h + '.output_layer.0.weight'), Wo[i])
            d = dict(Wq=Wq, Wg=Wg, bq=bq, Wl=Wl, bl=bl, Wqp=Wqp, Wo=Wo)
            return d
        self.pk = WC.packed(('dec', id(t[pf + 'query_lstm.weight_ih'])), None, la.pver(*ps), build)
        self.heads = heads

    def packT(self):
        """Transposed packs for the data-gradient GEMMs (training backward only): views - the GEMM reads them in place."""
        pk = self.pk
        if 'WqT' not in pk:
            pk.update(WqT=pk['Wq'].t(), WlT=pk['Wl'].t(), WgT=pk['Wg'].t(), WqpT=pk['Wqp'].t(), WoT=pk['Wo'].transpose(1, 2))
        return pk

    # ------------------------------------------------------------------ sequence-invariant precompute
    def precompute(self, nodes):
        """nodes: (nh, Bn, P, H) fp32.  Returns Kp, Vp (nh,Bn,P,H), glob (Bn,GH), Gq (Bn,4Hq), nodes_op."""
        be = ops.backend()
        t, pf = self.t, self.pf
        nh, Bn, P, H = nodes.shape
        n2 = op(nodes.view(nh * Bn * P, H))
        n_op = [n2[i * Bn * P:(i + 1) * Bn * P] for i in range(nh)]
        Kp = empty((nh, Bn, P, H), nodes)
        Vp = empty((nh, Bn, P, H), nodes)
        for i, h in enumerate(self.heads):
            be.gemm(n_op[i], WC.get(t[pf + h + '.K.weight']), Kp[i].view(Bn * P, H))
            be.gemm(n_op[i], WC.get(t[pf + h + '.V.weight']), Vp[i].view(Bn * P, H))
        glob = empty((Bn, self.GH), nodes)
        for i in range(nh):
            be.mean_nodes_fwd(nodes[i], glob[:, i * H:(i + 1) * H])
        Gq = empty((Bn, 4 * self.Hq), nodes)
        be.gemm(op(glob), self.pk['Wg'], Gq, bias=self.pk['bq'])
        self.hoist = be.attn2_supported(nh, P, self.Hq, H)
        self.KpVp = (Kp, Vp)
        if self.hoist:
            # fold the query projection into K and the output projection into V once per sequence:
            #   K_p.(Wq q) = (K_p Wq).q        Wo (sum_p a_p V_p) = sum_p a_p (Wo V_p)
            KW = empty((nh, Bn, P, self.Hq), nodes)
            VW = empty((nh, Bn, P, H), nodes)
            for i, h in enumerate(self.heads):
                be.gemm(op(Kp[i].view(Bn * P, H)), WC.get(t[pf + h + '.Q.weight'], transpose=True), KW[i].view(Bn * P, self.Hq))
                be.gemm(op(Vp[i].view(Bn * P, H)), WC.get(t[pf + h + '.output_layer.0.weight']), VW[i].view(Bn * P, H))
            return KW, VW, glob, Gq, n_op
        return Kp, Vp, glob, Gq, n_op

    # ------------------------------------------------------------------ one decode step
    def step(self, b, i, j, Kp, Vp, Gq, rows_per_node=1, drops=(None, None, None, None), gq_rows=None, lang_y=None,
             lang_stats=None):
        """Run step reading operand rows b.Xq[i], state slot i and writing slot j (= next step's inputs).

        b: namespace of buffers (see alloc()).  Gq rows are indexed by node set (rows_per_node).
        lang_y: (R, Hd) view receiving tanh(LN(lang_h)) (the vocabulary projection's operand row, layer.py:599)."""
        be = ops.backend()
        pk, t, pf = self.pk, self.t, self.pf
        nh, H, Hq, Hd = self.nh, self.H, self.Hq, self.Hd
        oq, ol, oQ = self.oq, self.ol, self.oQ
        dq, dc, dl, _ = drops
        R = b.Xq.shape[1]
        # split-K partial sums go straight into the cell kernel (no reduce launch)
        # weights are static inside the loop (b_static): their tiles are requested while the previous step's cell kernel runs
        # (the backend drops the request when the launch directly before is a weight conversion, e.g. at step 0)
        be.gemm(b.Xq[i], pk['Wq'], b.gq[:, i] if b.Sq > 1 else b.gq[0, i], splitk=b.Sq, b_static=True)
        # fused: split-K partials + hoisted global-feature bias -> cell -> q = dropout(LN(query_h))
        qy, qy2 = (b.q32[i], b.Xl[i][:, oq:oq + Hq]) if self.hoist else (b.Xl[i][:, oq:oq + Hq], None)
        lnq_w, lnq_b = t[pf + 'query_lstm_layernorm.weight'], t[pf + 'query_lstm_layernorm.bias']
        rb = Gq if gq_rows is None else gq_rows
        attn = None
        if self.hoist:
            # attention over the hoisted KW / VW (Kp, Vp hold them) + the context output layer tanh -> LayerNorm -> dropout of
            # both heads, written straight into the lang-LSTM operand row
            attn = dict(KW=Kp, VW=Vp, q=b.q32[i], alpha=b.alpha[i], co=b.co[i], scale=1.0 / math.sqrt(H), rows_per_node=rows_per_node,
                        ln=dict(gamma=[t[pf + h + '.output_layer.2.weight'] for h in self.heads],
                                beta=[t[pf + h + '.output_layer.2.bias'] for h in self.heads],
                                y=b.Xl[i][:, :nh * H], stats=b.statc[i], drop=dc, drop_head_stride=1 << 28))
        one_launch = False
        if self.fused and self.hoist and FUSED_CELL_ATTN and R <= FUSED_CELL_ATTN_MAX_ROWS:
            # cell + LayerNorm + attention + output layer of the step in ONE launch (csrc/fused_step.cu)
            one_launch = be.cell_norm_attn2_fwd(dict(gates=b.gq[:, i], c_prev=b.cq[i], c_out=b.cq[j], gamma=lnq_w, beta=lnq_b, y=qy,
                                                     h_out=b.qh[i], row_bias=rb, h2=b.Xq[j][:, oQ:oQ + Hq], y2=qy2, stats=b.statq[i],
                                                     ydrop=dq), attn)
        if one_launch:
            pass
        elif self.fused:
            be.lstm_cell_norm_fwd(b.gq[:, i], b.cq[i], b.cq[j], lnq_w, lnq_b, qy, h_out=b.qh[i], row_bias=rb,
                                  h2=b.Xq[j][:, oQ:oQ + Hq], y2=qy2, stats=b.statq[i], ydrop=dq)
        else:
            be.lstm_cell_fwd(b.gq[:, i], b.cq[i], b.cq[j], h_out=b.qh[i], row_bias=rb, h2=b.Xq[j][:, oQ:oQ + Hq])
            be.norm_fwd(b.qh[i], lnq_w, lnq_b, y=qy, y2=qy2, stats=b.statq[i], drop=dq)
        if one_launch:
            pass
        elif self.hoist:
            be.attn2_fwd(**attn)
        else:
            be.gemm(b.Xl[i][:, oq:oq + Hq], pk['Wqp'], b.qp[i])
            be.node_attn_fwd(Kp, Vp, b.qp[i], b.alpha[i], b.ctxr[i], rows_per_node)
            be.gemm(b.ctxr[i].view(R, nh, H).transpose(0, 1), pk['Wo'], b.co[i].view(R, nh, H).transpose(0, 1))
            for k, h in enumerate(self.heads):
                be.norm_fwd(b.co[i][:, k * H:(k + 1) * H], t[pf + h + '.output_layer.2.weight'], t[pf + h + '.output_layer.2.bias'],
                            y=b.Xl[i][:, k * H:(k + 1) * H], stats=b.statc[i, k], pre_tanh=True,
                            drop=(None if dc is None else (dc[0], dc[1], dc[2] + (k << 28))))
        be.gemm(b.Xl[i], pk['Wl'], b.gl[:, i] if b.Sl > 1 else b.gl[0, i], splitk=b.Sl, b_static=True)
        # fused: cell -> lang_h = dropout(h) (recurrent state, layer.py:594) -> tanh(LN(lang_h))
        lnl_w, lnl_b = t[pf + 'lang_lstm_layernorm.weight'], t[pf + 'lang_lstm_layernorm.bias']
        stl = lang_stats if lang_stats is not None else b.statl[i]
        if self.fused:
            be.lstm_cell_norm_fwd(b.gl[:, i], b.cl[i], b.cl[j], lnl_w, lnl_b, lang_y, h_out=b.lh[j], bias=pk['bl'],
                                  h2=b.Xq[j][:, :Hd], h3=b.Xl[j][:, ol:ol + Hd], drop=dl, stats=stl, post_tanh=True)
        else:
            be.lstm_cell_fwd(b.gl[:, i], b.cl[i], b.cl[j], h_out=b.lh[j], bias=pk['bl'], h2=b.Xq[j][:, :Hd],
                             h3=b.Xl[j][:, ol:ol + Hd], drop=dl)
            be.norm_fwd(b.lh[j], lnl_w, lnl_b, y=lang_y, stats=stl, post_tanh=True)

    def alloc(self, S, R, P, like):
        """Buffers for S step slots (+1) of R rows."""
        nh, H, Hq, Hd = self.nh, self.H, self.Hq, self.Hd
        b = type('Buf', (), {})()
        b.Xq = op_zeros((S + 1, R), self.Kq, like)
        b.Xl = op_zeros((S + 1, R), self.Kl, like)
        b.Sq, b.Sl = splitk_for(R, 4 * Hq, self.Kq), splitk_for(R, 4 * Hd, self.Kl)
        b.gq = empty((b.Sq, S, R, 4 * Hq), like)       # [0] ends up holding the activated gates (saved for BPTT)
        b.gl = empty((b.Sl, S, R, 4 * Hd), like)
        b.cq = zeros((S + 1, R, Hq), like)
        b.cl = zeros((S + 1, R, Hd), like)
        b.qh = empty((S, R, Hq), like)
        b.lh = zeros((S + 1, R, Hd), like)
        b.statq = empty((S, R, 2), like)
        b.statl = empty((S, R, 2), like)
        b.statc = empty((S, nh, R, 2), like)
        b.alpha = empty((S, R, nh * P), like)
        b.co = empty((S, R, nh * H), like)
        if getattr(self, 'hoist', False):
            b.q32 = empty((S, R, Hq), like)
        else:
            b.qp = empty((S, R, nh * H), like)
            b.ctxr = op_empty((S, R), nh * H, like)
        return b


def _stack_nodes(n1, n2, multi_modal):
    """layer.py:407-413: baseline concatenates both node sets along the node axis, multi-modal keeps two heads."""
    if n2 is not None and not multi_modal:
        raise NotImplementedError('Decoder(multi_modal=False) with cnn_feats_2 is not used by any reference model')
    if n2 is None:
        return _c(n1).unsqueeze(0)
    return torch.stack([_c(n1), _c(n2)], 0)


class DecoderTrainBlock:
    """Training branch (captions given): returns logits (B,T,V) and alpha (B,T,nh*P)."""

    def __init__(self, prefix, multi_modal, p_drop, training, T, tf_flags):
        self.pf, self.mm, self.p, self.training, self.T, self.tf = prefix, multi_modal, p_drop, training, T, tf_flags

    def forward(self, t):
        be = ops.backend()
        pf, T = self.pf, self.T
        core = DecoderCore(t, pf, self.mm, self.p)
        nodes = _stack_nodes(t['n1'], t.get('n2'), self.mm)
        nh, B, P, H = nodes.shape
        Hq, Hd, W, V = core.Hq, core.Hd, core.W, core.V
        caps = t['captions']
        Kp, Vp, glob, Gq, n_op = core.precompute(nodes)
        b = core.alloc(T, B, P, nodes)
        seed = next_seed()
        p = self.p if self.training else 0.0
        table = t[pf + 'word_embed.weight'].detach()
        wid = torch.empty((T + 1, B), dtype=torch.int64, device=nodes.device)
        wid[0].fill_(START)
        all_tf = all(self.tf)
        dw = site(p, seed, 900)
        oW = core.oW
        if all_tf:
            wid[1:T + 1].copy_(caps[:, :T].t())
            be.embedding_gather(table, wid[:T].view(-1), out=flat2(b.Xq[:T])[:, oW:oW + W], drop=dw)
        else:
            be.embedding_gather(table, wid[0], out=b.Xq[0][:, oW:oW + W], drop=dw)
        Dall = op_zeros((B, T), Hd, nodes)
        Wout = WC.get(t[pf + 'word_restore.weight'])
        bout = t[pf + 'word_restore.bias'].detach()
        logits = empty((B, T, V), nodes)
        step_logits = [False] * T
        drops_t = []
        for i in range(T):
            drops = (site(p, seed, 4 * i), site(0.1 if self.training else 0.0, seed, 4 * i + 1), site(p, seed, 4 * i + 2), None)
            drops_t.append(drops)
            core.step(b, i, i + 1, Kp, Vp, Gq, 1, drops, lang_y=Dall[:, i])
            if not all_tf:
                if self.tf[i]:
                    wid[i + 1].copy_(caps[:, i])
                else:
                    be.gemm(Dall[:, i], Wout, logits[:, i], bias=bout)
                    step_logits[i] = True
                    be.row_argmax(logits[:, i], wid[i + 1])
                if i + 1 < T:
                    dwi = None if dw is None else (dw[0], dw[1], dw[2] + (i + 1) * B * W)
                    be.embedding_gather(table, wid[i + 1], out=b.Xq[i + 1][:, oW:oW + W], drop=dwi)
        D2 = flat2(Dall)
        if all_tf:
            be.gemm(D2, Wout, logits.view(B * T, V), bias=bout)
        else:
            for i in range(T):
                if not step_logits[i]:
                    be.gemm(Dall[:, i], Wout, logits[:, i], bias=bout)
        alpha = b.alpha.permute(1, 0, 2)
        sv = dict(t=t, core=core, b=b, Kp=Kp, Vp=Vp, glob=glob, n_op=n_op, nodes=nodes, wid=wid, Dall=Dall, dw=dw,
                  drops=drops_t, dims=(nh, B, P, H, T))
        return [logits, alpha], sv

    def backward(self, sv, gouts):
        be = ops.backend()
        t, pf, core, b = sv['t'], self.pf, sv['core'], sv['b']
        nh, B, P, H, T = sv['dims']
        Hq, Hd, W, V, GH = core.Hq, core.Hd, core.W, core.V, core.GH
        Kq, Kl, oW, oQ, oq, ol = core.Kq, core.Kl, core.oW, core.oQ, core.oq, core.ol
        pk = core.packT()
        Kp, Vp = sv['Kp'], sv['Vp']
        ref = Kp
        grads = {}
        dlogits, dalpha = gouts
        heads = core.heads

        def lnp(name):
            w, bb = t[pf + name + '.weight'], t[pf + name + '.bias']
            dw_, db_ = small_zeros(w.shape, w), small_zeros(bb.shape, bb)
            grads[pf + name + '.weight'], grads[pf + name + '.bias'] = dw_, db_
            return w, bb, dw_, db_
        lnq = lnp('query_lstm_layernorm')
        lnl = lnp('lang_lstm_layernorm')
        lnc = [lnp(h + '.output_layer.2') for h in heads]
        TB = T * B
        # ---- vocabulary projection backward (batched over time)
        D2 = flat2(sv['Dall'])
        dDall = zeros((B, T, Hd), ref)
        if dlogits is not None:
            dl2 = _c(dlogits).view(B * T, V)
            dlo = op(dl2)
            dloT = dlo.t()
            wout = t[pf + 'word_restore.weight']
            be.gemm(dlo, WC.get(wout, transpose=True), dDall.view(B * T, Hd))
            grads[pf + 'word_restore.weight'] = la.mm(dloT, D2.t())
            dbo = small_zeros((V,), ref)
            be.colsum(dl2, dbo)
            grads[pf + 'word_restore.bias'] = dbo
        from . import functional as _DF
        # (data parallel) the vocabulary-projection gradients (16 M parameters, final here) travel during the BPTT
        _DF.early_sync(grads, (pf + 'word_restore.weight', pf + 'word_restore.bias'))
        da_ext = None
        if dalpha is not None:
            da_ext = _c(dalpha.permute(1, 0, 2))                 # (T,B,nh*P)
        # ---- BPTT
        dXq = zeros((T + 1, B, Kq), ref)
        dXl = zeros((T + 1, B, Kl), ref)
        dgq_all = op_empty((TB,), 4 * Hq, ref)
        dgl_all = op_empty((TB,), 4 * Hd, ref)
        dgq32 = empty((B, 4 * Hq), ref)
        dgq_sum = zeros((B, 4 * Hq), ref)
        hoist = core.hoist
        if hoist:
            lc_g, lc_b = empty((T, B, nh * H), ref), empty((T, B, nh * H), ref)     # per-row ctx-LayerNorm gradient terms
            # d(logits) and d(co) of every step: the node gradients dKW / dVW are summed over time by ONE launch after the loop
            dl_all, dco_sv = empty((T, B, nh * P), ref), empty((T, B, nh * H), ref)
        else:
            dqp_all = op_empty((TB,), nh * H, ref)
            dco_all = op_empty((T, B), nh * H, ref)
            dctxr = empty((B, nh * H), ref)
        if hoist:
            dKp, dVp = empty(Kp.shape, ref), empty(Vp.shape, ref)    # dKW, dVW: written once by attn2_bwd_nodes
        else:
            dKp, dVp = zeros(Kp.shape, ref), zeros(Vp.shape, ref)
        att_scale = 1.0 / math.sqrt(H)
        dcq, dcq2 = small_zeros((B, Hq), ref), empty((B, Hq), ref)
        dcl, dcl2 = zeros((B, Hd), ref), empty((B, Hd), ref)
        fused = core.fused
        if fused:
            # per-row LayerNorm parameter-gradient contributions of every step; reduced by one colsum each after BPTT
            lq_g, lq_b = empty((T, B, Hq), ref), empty((T, B, Hq), ref)
            ll_g, ll_b = empty((T, B, Hd), ref), empty((T, B, Hd), ref)
        for i in range(T - 1, -1, -1):
            j = i + 1
            dq, dc, dl, _ = sv['drops'][i]
            rows = slice(i * B, (i + 1) * B)
            # lang LN+tanh -> grad wrt dropped lang_h(i): accumulate onto the recurrent grad from step i+1 (dXq[j][:, :Hd])
            # fused: grad wrt dropped lang_h(i) = LN/tanh path (dDall) + recurrent paths from step i+1 (Xq and Xl rows)
            if fused:
                be.norm_lstm_cell_bwd(b.gl[0, i], b.cl[i], b.cl[j], dcl, dcl2, dDall[:, i], b.lh[j], lnl[0], lnl[1], b.statl[i],
                                      ll_g[i], ll_b[i], dh=dXq[j][:, :Hd], dh2=dXl[j][:, ol:ol + Hd], dgates2=dgl_all[rows],
                                      drop=dl, post_tanh=True)
            else:
                be.norm_bwd(dDall[:, i], b.lh[j], lnl[0], lnl[1], b.statl[i], dx=dXq[j][:, :Hd], dgamma=lnl[2], dbeta=lnl[3],
                            post_tanh=True, dx_accum=True)
                be.lstm_cell_bwd(b.gl[0, i], b.cl[i], b.cl[j], dXq[j][:, :Hd], dcl, dcl2, dgates2=dgl_all[rows],
                                 drop=dl, dh2=dXl[j][:, ol:ol + Hd])
            dcl, dcl2 = dcl2, dcl
            be.gemm(dgl_all[rows], pk['WlT'], dXl[i], atomic=True, b_static=True)      # dXl / dXq start as zeros: split-K lands by atomic adds
            if hoist:
                # one kernel: output-layer backward (dropout, LayerNorm, tanh) of both heads -> d(alpha), softmax backward,
                # dq += sum_h sum_p dl KW, dKW / dVW accumulated over time; LN parameter gradients as per-row contributions
                be.attn2_bwd(Kp, Vp, b.q32[i], b.alpha[i], None, dXl[i][:, oq:oq + Hq], dKp, dVp, att_scale,
                             dalpha_ext=(da_ext[i] if da_ext is not None else None),
                             ln=dict(dy=dXl[i][:, :nh * H], co=b.co[i], gamma=[lnc[k][0] for k in range(nh)], stats=b.statc[i],
                                     dgamma_rows=lc_g[i], dbeta_rows=lc_b[i], drop=dc, drop_head_stride=1 << 28),
                             save=(dl_all[i], dco_sv[i]))       # node gradients deferred to ONE launch after the loop
            else:
                for k, h in enumerate(heads):
                    be.norm_bwd(dXl[i][:, k * H:(k + 1) * H], b.co[i][:, k * H:(k + 1) * H], lnc[k][0], lnc[k][1], b.statc[i, k],
                                dx=dco_all[i][:, k * H:(k + 1) * H], dgamma=lnc[k][2], dbeta=lnc[k][3],
                                pre_tanh=True, drop=(None if dc is None else (dc[0], dc[1], dc[2] + (k << 28))))
                be.gemm(dco_all[i].view(B, nh, H).transpose(0, 1), pk['WoT'], dctxr.view(B, nh, H).transpose(0, 1))
                be.node_attn_bwd(Kp, Vp, b.qp[i], b.alpha[i], dctxr, dqp_all[rows], dKp, dVp,
                                 dalpha_ext=(da_ext[i] if da_ext is not None else None))
                be.gemm(dqp_all[rows], pk['WqpT'], dXl[i][:, oq:oq + Hq], accum=True)
            # query LN -> grad wrt query_h(i): accumulate onto recurrent grad from step i+1 (dXq[j][:, oQ:])
            # fused: grad wrt query_h(i) = LN path (dq) + recurrent path from step i+1; gate grads also summed over time
            if fused:
                be.norm_lstm_cell_bwd(b.gq[0, i], b.cq[i], b.cq[j], dcq, dcq2, dXl[i][:, oq:oq + Hq], b.qh[i], lnq[0], lnq[1],
                                      b.statq[i], lq_g[i], lq_b[i], dh=dXq[j][:, oQ:oQ + Hq], dgates2=dgq_all[rows],
                                      dgates_sum=dgq_sum, ydrop=dq)
            else:
                be.norm_bwd(dXl[i][:, oq:oq + Hq], b.qh[i], lnq[0], lnq[1], b.statq[i], dx=dXq[j][:, oQ:oQ + Hq], dgamma=lnq[2],
                            dbeta=lnq[3], drop=dq, dx_accum=True)
                be.lstm_cell_bwd(b.gq[0, i], b.cq[i], b.cq[j], dXq[j][:, oQ:oQ + Hq], dcq, dcq2, dgates=dgq32,
                                 dgates2=dgq_all[rows])
                be.axpby(dgq32, 1.0, dgq_sum, 1.0)
            dcq, dcq2 = dcq2, dcq
            be.gemm(dgq_all[rows], pk['WqT'], dXq[i], atomic=True, b_static=True)
        # ---- parameter gradients batched over time
        if fused:
            be.colsum(lq_g.view(TB, Hq), lnq[2])
            be.colsum(lq_b.view(TB, Hq), lnq[3])
            be.colsum(ll_g.view(TB, Hd), lnl[2])
            be.colsum(ll_b.view(TB, Hd), lnl[3])
        if hoist:
            be.attn2_bwd_nodes(b.q32[:T], dl_all, b.alpha[:T], dco_sv, dKp, dVp)
            for k in range(nh):
                be.colsum(lc_g.view(TB, nh * H)[:, k * H:(k + 1) * H], lnc[k][2])
                be.colsum(lc_b.view(TB, nh * H)[:, k * H:(k + 1) * H], lnc[k][3])
        Xq2, Xl2 = flat2(b.Xq[:T]), flat2(b.Xl[:T])
        # time-batched weight-gradient GEMMs read the gate gradients transposed IN PLACE (MN-major operands)
        dgqT, dglT = dgq_all.t(), dgl_all.t()
        dWq = la.mm(dgqT, Xq2.t())                                # (4Hq, Kq)
        dWl = la.mm(dglT, Xl2.t())                                # (4Hd, Kl)
        dgs_op = op(dgq_sum)
        dWg = la.mm(dgq_sum.t(), sv['glob'].t())                  # (4Hq, GH)
        dglob = empty((B, GH), ref)
        be.gemm(dgs_op, pk['WgT'], dglob)
        dwih = empty((4 * Hq, Hd + GH + W), ref)
        be.convert(dWq[:, :Hd], dst=dwih[:, :Hd])
        be.convert(dWg, dst=dwih[:, Hd:Hd + GH])
        be.convert(dWq[:, oW:oW + W], dst=dwih[:, Hd + GH:])
        grads[pf + 'query_lstm.weight_ih'] = dwih
        grads[pf + 'query_lstm.weight_hh'] = dWq[:, oQ:oQ + Hq]
        dbq = small_zeros((4 * Hq,), ref)
        be.colsum(dgq_all, dbq)
        grads[pf + 'query_lstm.bias_ih'] = dbq
        grads[pf + 'query_lstm.bias_hh'] = dbq
        dlih = empty((4 * Hd, nh * H + Hq), ref)
        be.convert(dWl[:, :nh * H], dst=dlih[:, :nh * H])
        be.convert(dWl[:, oq:oq + Hq], dst=dlih[:, nh * H:])
        grads[pf + 'lang_lstm.weight_ih'] = dlih
        grads[pf + 'lang_lstm.weight_hh'] = dWl[:, ol:ol + Hd]
        dbl = small_zeros((4 * Hd,), ref)
        be.colsum(dgl_all, dbl)
        grads[pf + 'lang_lstm.bias_ih'] = dbl
        grads[pf + 'lang_lstm.bias_hh'] = dbl
        dnodes = empty((nh, B, P, H), ref)
        n_op = sv['n_op']
        if hoist:
            # unfold the hoisted projections: KW = K Wq, VW = V Wo^T  (K, V = node projections saved by precompute)
            K0, V0 = core.KpVp
            dKW, dVW = dKp, dVp
            dKp, dVp = empty(K0.shape, ref), empty(V0.shape, ref)
            for k, h in enumerate(heads):
                dKWo, dVWo = op(dKW[k].view(B * P, Hq)), op(dVW[k].view(B * P, H))
                Ko, Vo = op(K0[k].view(B * P, H)), op(V0[k].view(B * P, H))
                grads[pf + h + '.Q.weight'] = la.mm(Ko.t(), dKWo.t())                                  # (H, Hq)
                grads[pf + h + '.output_layer.0.weight'] = la.mm(dVWo.t(), Vo.t())                    # (H, H)
                be.gemm(dKWo, WC.get(t[pf + h + '.Q.weight']), dKp[k].view(B * P, H))
                be.gemm(dVWo, WC.get(t[pf + h + '.output_layer.0.weight'], transpose=True), dVp[k].view(B * P, H))
        else:
            dWqp = la.mm(dqp_all.t(), Xl2[:, oq:oq + Hq].t())         # (nh*H, Hq)
            dco2, ctx2 = flat2(dco_all), flat2(b.ctxr)
        for k, h in enumerate(heads):
            if not hoist:
                grads[pf + h + '.Q.weight'] = dWqp[k * H:(k + 1) * H]
                grads[pf + h + '.output_layer.0.weight'] = la.mm(dco2[:, k * H:(k + 1) * H].t(), ctx2[:, k * H:(k + 1) * H].t())
            dK2, dV2 = dKp[k].view(B * P, H), dVp[k].view(B * P, H)
            dKo, dVo = op(dK2), op(dV2)
            grads[pf + h + '.K.weight'] = la.mm(dKo.t(), n_op[k].t())
            grads[pf + h + '.V.weight'] = la.mm(dVo.t(), n_op[k].t())
            be.gemm(dKo, WC.get(t[pf + h + '.K.weight'], transpose=True), dnodes[k].view(B * P, H))
            be.gemm(dVo, WC.get(t[pf + h + '.V.weight'], transpose=True), dnodes[k].view(B * P, H), accum=True)
            be.mean_nodes_bwd(dglob[:, k * H:(k + 1) * H], dnodes[k])
        # ---- word embedding: every step's word row came from table[wid[i]] (teacher forced or arg-max fed back)
        dE = zeros((V, W), ref)
        be.embedding_scatter_add(dE, sv['wid'][:T].reshape(-1), flat2(dXq[:T])[:, oW:oW + W], drop=sv['dw'])
        grads[pf + 'word_embed.weight'] = dE
        grads['n1'] = dnodes[0]
        if nh > 1:
            grads['n2'] = dnodes[1]
        return grads


# =============================================================================================== inference
def _decode_setup(t, pf, multi_modal, n1, n2, R_per_clip):
    WC.eval_scope()
    core = DecoderCore(t, pf, multi_modal, 0.0)
    nodes = _stack_nodes(n1, n2, multi_modal)
    nh, B, P, H = nodes.shape
    Kp, Vp, glob, Gq, _ = core.precompute(nodes)
    b = core.alloc(3, B * R_per_clip, P, nodes)
    if R_per_clip > 1:
        Gq = Gq.repeat_interleave(R_per_clip, 0)
    return core, nodes, Kp, Vp, Gq, b


def _logits_buf(rows, V, like):
    """(rows, V) fp32 logits of one decode step with a row pitch padded to 8 elements: V = 10547 is odd, and only 16-byte
    aligned rows get the GEMM epilogue's vector stores (the arg-max / top-k kernels take the pitch)."""
    return empty((rows, (V + 7) // 8 * 8), like)[:, :V]


def _step_logits(core, b, i, j, Kp, Vp, Gq, rpn, dbuf, logits):
    be = ops.backend()
    t, pf = core.t, core.pf
    core.step(b, i, j, Kp, Vp, Gq, rpn, lang_y=dbuf)
    be.gemm(dbuf, WC.get(t[pf + 'word_restore.weight']), logits, bias=t[pf + 'word_restore.bias'].detach(), b_static=True)


def decode_greedy(t, pf, multi_modal, n1, n2, T):
    """layer.py:426-447 with captions=None, beam_size==1: always T steps, ids (B,T) int64."""
    be = ops.backend()
    core, nodes, Kp, Vp, Gq, b = _decode_setup(t, pf, multi_modal, n1, n2, 1)
    B = nodes.shape[1]
    table = t[pf + 'word_embed.weight'].detach()
    ids = torch.empty((B, T + 1), dtype=torch.int64, device=nodes.device)
    ids[:, 0].fill_(START)
    dbuf = op_zeros((B,), core.Hd, nodes)
    logits = _logits_buf(B, core.V, nodes)
    oW, W = core.oW, core.W
    for s in range(T):
        i, j = s % 2, (s + 1) % 2
        be.embedding_gather(table, ids[:, s], out=b.Xq[i][:, oW:oW + W])
        _step_logits(core, b, i, j, Kp, Vp, Gq, 1, dbuf, logits)
        be.row_argmax(logits, ids[:, s + 1])
    return ids[:, 1:]


def decode_beam(t, pf, multi_modal, n1, n2, T, beam, end_index, per_node=None):
    """layer.py:449-460 + allennlp_beamsearch.py:51-294, all beams batched (B*beam rows per step).

    Returns (best (B,S), all_predictions (B,beam,S), log_probs (B,beam)); S = steps taken (early stop when
    every beam of every clip has emitted <end>, exactly like allennlp_beamsearch.py:168)."""
    preds, backs, final_lp, first_lp = decode_beam_core(t, pf, multi_modal, n1, n2, T, beam, end_index, per_node)
    return decode_beam_finish(preds, backs, final_lp, first_lp, end_index)


def decode_beam_core(t, pf, multi_modal, n1, n2, T, beam, end_index, per_node=None):
    """The T beam steps without any host synchronisation (CUDA-graph capturable).  Returns the raw search trace:
    preds (T,B,beam), back-pointers (T-1,B,beam), log-probs after the last step and after step 0."""
    be = ops.backend()
    k = per_node or beam
    core, nodes, Kp, Vp, Gq, b = _decode_setup(t, pf, multi_modal, n1, n2, beam)
    B = nodes.shape[1]
    R = B * beam
    dev = nodes.device
    table = t[pf + 'word_embed.weight'].detach()
    oW, W = core.oW, core.W
    if k > core.V:
        from .errors import ConfigurationError
        raise ConfigurationError('Target vocab size (%d) too small relative to per_node_beam_size (%d).\n'
                                 'Please decrease beam_size or per_node_beam_size.' % (core.V, k))
    preds = torch.empty((T, B, beam), dtype=torch.int64, device=dev)
    backs = torch.empty((max(T - 1, 1), B, beam), dtype=torch.int64, device=dev)
    lps = empty((2, B, beam), nodes)
    top_lp = empty((R, k), nodes)
    top_id = torch.empty((R, k), dtype=torch.int64, device=dev)
    start = torch.full((R,), START, dtype=torch.int64, device=dev)
    dbuf = op_zeros((R,), core.Hd, nodes)
    logits = _logits_buf(R, core.V, nodes)
    # step 0: every beam row of a clip is identical; candidates come from the first row of each clip
    be.embedding_gather(table, start, out=b.Xq[0][:, oW:oW + W])
    _step_logits(core, b, 0, 1, Kp, Vp, Gq, beam, dbuf, logits)
    be.beam_topk(logits[::beam], None, end_index, beam, lps[0], preds[0])
    cur, i, j, g = 0, 1, 2, 0          # state lives in slot i; slot j receives the step; g is the gather target
    for s in range(1, T):
        last = preds[s - 1].view(R)
        be.embedding_gather(table, last, out=b.Xq[i][:, oW:oW + W])
        _step_logits(core, b, i, j, Kp, Vp, Gq, beam, dbuf, logits)
        be.beam_topk(logits, last, end_index, k, top_lp, top_id)
        be.beam_merge(top_lp, top_id, lps[cur], B, beam, k, lps[1 - cur], preds[s], backs[s - 1], None, end_index)
        cur = 1 - cur
        # keep only the state rows of the surviving ancestors (h/c of both LSTMs; node tensors are indexed, not copied)
        be.beam_gather_multi([(_fullrows(buf[j]), _fullrows(buf[g])) for buf in (b.Xq, b.Xl, b.cq, b.cl)], backs[s - 1], B, beam)
        i, j, g = g, i, j
    return preds, backs, lps[cur], lps[0]


def decode_beam_finish(preds, backs, last_lp, first_lp, end_index):
    """Early-stop length (one D2H read per search), back-track, best beam (layer.py:455-460)."""
    be = ops.backend()
    T, B, beam = preds.shape
    dev = preds.device
    # early-stop semantics: the reference breaks before step s when all of preds[s-1] are <end>
    ended = (preds == end_index).view(T, -1).all(1).tolist()        # one D2H read per search
    S = T
    for s in range(1, T):
        if ended[s - 1]:
            S = s
            break
    out = torch.empty((B, beam, S), dtype=torch.int64, device=dev)
    if S == 1:
        out.copy_(preds[0].unsqueeze(-1))
        final_lp = first_lp
    else:
        be.beam_backtrack(preds, backs, S, B, beam, out)
        # last_lp holds step T-1; after an early stop every later step only appended <end> with log-prob 0
        final_lp = last_lp
    best = torch.empty((B,), dtype=torch.int64, device=dev)
    be.row_argmax(final_lp, best)
    return out[torch.arange(B, device=dev), best], out, final_lp


def _fullrows(x):
    """The full (padded) contiguous rows behind an operand-buffer slot view."""
    if x.is_contiguous():
        return x
    return x.as_strided((x.shape[0], x.stride(0)), (x.stride(0), 1))


def decode_api(t, multi_modal, word, qh, qc, lh, lc, global_feat, n1, n2=None, rows_per_node=1):
    """Decoder.decode (layer.py:569-602) with explicit state tensors: one step for R rows.

    Returns (word_logits, query_h, query_c, lang_h, lang_c, alpha (R, nh*P, 1))."""
    be = ops.backend()
    WC.eval_scope()
    core = DecoderCore(t, '', multi_modal, 0.0)
    nodes = _stack_nodes(n1, n2, multi_modal)
    nh, Bn, P, H = nodes.shape
    R = word.shape[0]
    Kp, Vp, _, Gq, _ = core.precompute(nodes)
    # the caller supplies global_feat explicitly: rebuild the hoisted gate term from it
    Gq = empty((global_feat.shape[0], 4 * core.Hq), nodes)
    be.gemm(op(_c(global_feat)), core.pk['Wg'], Gq, bias=core.pk['bq'])
    b = core.alloc(1, R, P, nodes)
    oW, oQ, ol, W, Hq, Hd = core.oW, core.oQ, core.ol, core.W, core.Hq, core.Hd
    be.convert(_c(lh), dst=b.Xq[0][:, :Hd])
    be.convert(_c(word), dst=b.Xq[0][:, oW:oW + W])
    be.convert(_c(qh), dst=b.Xq[0][:, oQ:oQ + Hq])
    be.convert(_c(lh), dst=b.Xl[0][:, ol:ol + Hd])
    b.cq[0].copy_(qc)
    b.cl[0].copy_(lc)
    rpn = rows_per_node if Gq.shape[0] != R else 1
    if Kp.shape[1] == R:
        rpn = 1
    dbuf = op_zeros((R,), Hd, nodes)
    logits = empty((R, core.V), nodes)
    gq_rows = Gq if Gq.shape[0] == R else Gq.repeat_interleave(R // Gq.shape[0], 0)
    core.step(b, 0, 1, Kp, Vp, gq_rows, R // Kp.shape[1], lang_y=dbuf)
    be.gemm(dbuf, WC.get(t['word_restore.weight']), logits, bias=t['word_restore.bias'].detach())
    return logits, b.qh[0], b.cq[1], b.lh[1], b.cl[1], b.alpha[0].unsqueeze(-1)


def beam_step_api(t, multi_modal, batch_size, last_predictions, state):
    """Decoder.beam_step (layer.py:489-567): AllenNLP step-function contract over a state dict with keys
    query_lstm_h/c, lang_lstm_h/c, cnn_feats, global_feat[, cnn_feats_2], each (group, ...)."""
    be = ops.backend()
    G_ = last_predictions.shape[0]
    table = t['word_embed.weight']
    word = empty((G_, table.shape[1]), table)
    be.embedding_gather(table, _c(last_predictions), out=word)
    logits, qh, qc, lh, lc, _ = decode_api(t, multi_modal, word, state['query_lstm_h'], state['query_lstm_c'],
                                           state['lang_lstm_h'], state['lang_lstm_c'], state['global_feat'],
                                           state['cnn_feats'], state.get('cnn_feats_2'))
    logp = empty(logits.shape, logits)
    be.log_softmax(logits, logp)
    new_state = dict(state)
    new_state.update(query_lstm_h=qh, query_lstm_c=qc, lang_lstm_h=lh, lang_lstm_c=lc)
    return logp, new_state
