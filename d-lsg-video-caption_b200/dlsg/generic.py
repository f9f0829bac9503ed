"""Generic differentiable primitives over libdlsg kernels (used by the discriminator path and by the
stand-alone sub-layer forwards).  Each primitive is a torch.autograd.Function whose backward is itself
written with these primitives, so gradients of gradients (WGAN-GP, run_gun.py:362-371) work.
"""
import math
import os

import torch

from . import _lib as L
from . import linalg as la
from . import ops
from .linalg import empty, zeros


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def param_scope(fn):
    """Decorator for generic-path module forwards: opens a parameter-copy epoch (dlsg.linalg.param_epoch_scope) so bf16
    weight copies are refreshed once per top-level forward (fused optimizers do not bump tensor versions)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        with la.param_epoch_scope():
            return fn(*a, **k)
    return wrapped


# ----------------------------------------------------------------------------------------------- gradient scope
_INPUT_GRADS_ONLY = False


class input_grads_only:
    """Scope for a backward pass that wants gradients of NON-LEAF tensors only (the WGAN-GP penalty's
    torch.autograd.grad(m_logit, tok_mixed, create_graph=True), run_gun.py:362-371): inside it the backward functions skip
    the gradients of leaves and of views of leaves (weights, biases: a weight-gradient GEMM + operand casts + a column sum
    per Linear, all of which autograd would compute - needs_input_grad is fixed at forward time - and then discard)."""

    def __enter__(self):
        global _INPUT_GRADS_ONLY
        self.old, _INPUT_GRADS_ONLY = _INPUT_GRADS_ONLY, True

    def __exit__(self, *exc):
        global _INPUT_GRADS_ONLY
        _INPUT_GRADS_ONLY = self.old


def _param_like(t):
    """A leaf that requires grad (a parameter) or a view of one."""
    base = t._base if (t._is_view() and t._base is not None) else t
    return base.grad_fn is None and base.requires_grad


def _want(ctx, i, t):
    if not ctx.needs_input_grad[i]:
        return False
    return not (_INPUT_GRADS_ONLY and _param_like(t))


# ----------------------------------------------------------------------------------------------- matmul
class _MmNT(torch.autograd.Function):
    """y[..., M, N] = a[..., M, K] @ b[..., N, K]^T (batched when 3-D)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        if (SMALL_BMM_SIMT and a.dim() == 3 and a.dtype == torch.float32 and b.dtype == torch.float32
                and (a.shape[-1] <= 32 or (SMALL_BMM_SIMT > 1 and min(a.shape[-2], b.shape[-2]) <= 32))):
            # measurement switch, OFF by default: the per-sample products of the critic's attention / node scoring on the fp32
            # FFMA kernel (no operand casts) are SLOWER than the mostly empty tensor-core tiles + two casts - 45.6 ms per GAN
            # iteration with all of them there, 40.5 ms with only the short reductions (K = 26 / 5), 37.7 ms on the tensor-core
            # kernel (profiles/r04i_*, r04j_*): the FFMA kernel's 64 x 64 tiles are ~28 us per batched launch at these shapes.
            return la.mm32(a, b)
        return la.mm(a, b, memo=True)

    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        da = db = None
        # An operand that is a transposed VIEW (a weight read as W^T by a data-gradient product) gets its gradient computed in
        # the layout of the tensor underneath and handed back as a transposed view of that: once autograd has undone the
        # transpose the gradient is contiguous - no strided accumulate / copy kernels behind every such product.
        if _want(ctx, 0, a):
            if _tview(a):
                da = bmm_nt(b.transpose(-1, -2), dy).transpose(-1, -2)            # (dy @ b)^T = b^T @ dy^T
            else:
                da = bmm_nt(dy, b.transpose(-1, -2))                              # dy @ b
        if _want(ctx, 1, b):
            if _tview(b):
                db = bmm_nt(a.transpose(-1, -2), dy.transpose(-1, -2)).transpose(-1, -2)   # (dy^T @ a)^T = a^T @ dy
            else:
                db = bmm_nt(dy.transpose(-1, -2), a.transpose(-1, -2))            # dy^T @ a
        return da, db


def _tview(x):
    return x.dim() >= 2 and x.shape[-1] > 1 and x.shape[-2] > 1 and x.stride(-1) != 1 and x.stride(-2) == 1


def bmm_nt(a, b):
    """a (..,M,K) @ b (..,N,K)^T with identical leading batch dims (2-D or 3-D)."""
    return _MmNT.apply(a, b)


def matmul_nn(a, b):
    return bmm_nt(a, b.transpose(-1, -2))


class _ColSum(torch.autograd.Function):
    """out[c] = sum_r x[r,c] (bias gradients) on the colsum kernel; its backward is a broadcast."""

    @staticmethod
    def forward(ctx, x):
        ctx.rows = x.shape[0]
        out = la.small_zeros((x.shape[1],), x)
        ops.backend().colsum(_c(x), out)
        return out

    @staticmethod
    def backward(ctx, g):
        return g.unsqueeze(0).expand(ctx.rows, g.shape[0])


class _LinearB(torch.autograd.Function):
    """y = a @ w^T + bias with the bias added in the GEMM epilogue; bias gradient on the colsum kernel."""

    @staticmethod
    def forward(ctx, a, w, bias):
        ctx.save_for_backward(a, w)
        ctx.bias_is_param = _param_like(bias)
        return la.mm(a, w, memo=True, bias=bias.detach())

    @staticmethod
    def backward(ctx, dy):
        a, w = ctx.saved_tensors
        da = dw = db = None
        if _want(ctx, 0, a):
            da = bmm_nt(dy, w.transpose(-1, -2))
        if _want(ctx, 1, w):
            dw = bmm_nt(dy.transpose(-1, -2), a.transpose(-1, -2))
        if ctx.needs_input_grad[2] and not (_INPUT_GRADS_ONLY and ctx.bias_is_param):
            db = _ColSum.apply(dy)
        return da, dw, db


def linear(x, w, b=None, tanh=False):
    """nn.Linear over the last dim of x (any leading dims)."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    y = bmm_nt(x2, w) if b is None else _LinearB.apply(x2, w, b)
    if tanh:
        y = tanh_(y)
    return y.view(*lead, w.shape[0])


def _ew(op, ins, n_out, cols=0, want=None):
    """One dlsg_ew launch: contiguous fp32 inputs -> n_out fresh outputs (want[k] False: that output is not computed)."""
    ins = [_c(t) for t in ins]
    outs = [torch.empty_like(ins[0]) if (want is None or want[k]) else None for k in range(n_out)]
    ops.backend().ew(op, ins, outs, cols)
    return outs


class _Tanh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = torch.tanh(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return _TanhBwd.apply(dy, y)


class _TanhBwd(torch.autograd.Function):
    """dx = dy (1 - y^2) (one launch); its own backward is the closed form (u (1 - y^2), -2 y dy u) (one launch)."""

    @staticmethod
    def forward(ctx, dy, y):
        dy, y = _c(dy), _c(y)
        ctx.save_for_backward(dy, y)
        return _ew(L.EW_TANH_BWD, [dy, y], 1)[0].view(dy.shape)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u):
        dy, y = ctx.saved_tensors
        g_dy, g_y = _ew(L.EW_TANH_BWD2, [dy, y, u], 2, want=ctx.needs_input_grad[:2])
        return g_dy, g_y


def tanh_(x):
    return _Tanh.apply(x)


class _Mul(torch.autograd.Function):
    """Same-shape product whose backward pair (dy b, dy a) and second-order triple are one launch each."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return a * b

    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        return _MulBwd.apply(dy, a, b)


class _MulBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dy, a, b):
        dy, a, b = _c(dy), _c(a), _c(b)
        ctx.save_for_backward(dy, a, b)
        da, db = _ew(L.EW_MUL_BWD, [dy, a, b], 2)
        return da, db

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u0, u1):
        dy, a, b = ctx.saved_tensors
        return tuple(_ew(L.EW_MUL_BWD2, [dy, a, b, u0, u1], 3))


def mul(a, b):
    if a.shape == b.shape and a.dtype == torch.float32 and b.dtype == torch.float32 and (a.requires_grad or b.requires_grad):
        return _Mul.apply(a, b)
    return a * b


class _LerpRows(torch.autograd.Function):
    """out[r] = a[r] e[r] + b[r] (1 - e[r]) with one coefficient per leading row (the WGAN-GP interpolate, run_gun.py:355-358);
    e carries no gradient.  The backward pair (dy e, dy (1 - e)) is linear in dy, so its own backward is this function."""

    @staticmethod
    def forward(ctx, a, b, e):
        a, b = _c(a), _c(b)
        ctx.save_for_backward(e)
        return _ew(L.EW_LERP_ROWS, [a, b, e.reshape(-1)], 1, cols=a.numel() // e.numel())[0]

    @staticmethod
    def backward(ctx, dy):
        (e,) = ctx.saved_tensors
        da, db = _LerpRowsBwd.apply(dy, e)
        return da, db, None


class _LerpRowsBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dy, e):
        dy = _c(dy)
        ctx.save_for_backward(e)
        da, db = _ew(L.EW_LERP_ROWS_BWD, [dy, e.reshape(-1)], 2, cols=dy.numel() // e.numel())
        return da, db

    @staticmethod
    def backward(ctx, u0, u1):
        (e,) = ctx.saved_tensors
        return _LerpRows.apply(u0, u1, e), None


def lerp_rows(a, b, e):
    """a * e + b * (1 - e), e of shape (rows, 1, ..., 1) broadcast over everything but the leading dim."""
    return _LerpRows.apply(a, b, e)


# ----------------------------------------------------------------------------------------------- softmax
class _Softmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dim, scale, mask, mask_mode):
        x3, shp = _as3(x, dim)
        m3 = _as3(mask, dim)[0] if mask is not None else None
        y = torch.empty_like(x3)
        ops.backend().softmax_fwd(x3, y, 1, scale, m3, mask_mode)
        ctx.save_for_backward(x, mask)
        ctx.dim, ctx.scale, ctx.mask_mode = dim, scale, mask_mode
        return _from3(y, shp, dim)

    @staticmethod
    def backward(ctx, dy):
        x, mask = ctx.saved_tensors
        return _softmax_bwd(x, dy, ctx.dim, ctx.scale, mask, ctx.mask_mode), None, None, None, None


def _as3(x, dim):
    """Move `dim` to the middle of a contiguous (outer, n, inner) view."""
    dim = dim % x.dim()
    x = _c(x)
    outer = 1
    for s in x.shape[:dim]:
        outer *= s
    inner = 1
    for s in x.shape[dim + 1:]:
        inner *= s
    return x.view(outer, x.shape[dim], inner), x.shape


def _from3(y, shp, dim):
    return y.view(shp)


class _SoftmaxBwd(torch.autograd.Function):
    """dx of the softmax (one launch); its own backward with respect to (x, dy) is the closed-form dlsg_softmax_bwd2."""

    @staticmethod
    def forward(ctx, x, dy, dim, scale, mask, mask_mode):
        x3, shp = _as3(x, dim)
        d3 = _as3(dy, dim)[0]
        m3 = _as3(mask, dim)[0] if mask is not None else None
        dx = torch.empty_like(x3)
        ops.backend().softmax_bwd(x3, d3, dx, 1, scale, m3, mask_mode)
        ctx.save_for_backward(x3, d3, m3)
        ctx.shp, ctx.scale, ctx.mask_mode = shp, scale, mask_mode
        return dx.view(shp)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u):
        x3, d3, m3 = ctx.saved_tensors
        u3 = _c(u).view(x3.shape)
        g_x = torch.empty_like(x3) if ctx.needs_input_grad[0] else None
        g_dy = torch.empty_like(x3) if ctx.needs_input_grad[1] else None
        ops.backend().softmax_bwd2(x3, d3, u3, 1, g_dy=g_dy, g_x=g_x, scale=ctx.scale, mask=m3, mask_mode=ctx.mask_mode)
        return (g_x.view(ctx.shp) if g_x is not None else None), (g_dy.view(ctx.shp) if g_dy is not None else None), None, None, None, None


def _softmax_bwd(x, dy, dim, scale, mask, mask_mode):
    """dx of softmax; differentiable again through the closed-form `_SoftmaxBwd`."""
    if torch.is_grad_enabled() and (x.requires_grad or dy.requires_grad):
        return _SoftmaxBwd.apply(x, dy, dim, scale, mask, mask_mode)
    x3, shp = _as3(x, dim)
    d3 = _as3(dy, dim)[0]
    m3 = _as3(mask, dim)[0] if mask is not None else None
    dx = torch.empty_like(x3)
    ops.backend().softmax_bwd(x3, d3, dx, 1, scale, m3, mask_mode)
    return dx.view(shp)


def softmax(x, dim, scale=1.0, mask=None, mask_mode=0):
    return _Softmax.apply(x, dim, scale, mask, mask_mode if mask is not None else 0)


# ----------------------------------------------------------------------------------------------- norm family
class _Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, pre_tanh, drop):
        x2 = _c(x).view(-1, x.shape[-1])
        y = torch.empty_like(x2)
        stats = empty((x2.shape[0], 2), x2)
        ops.backend().norm_fwd(x2, gamma.detach(), beta.detach(), y=y, stats=stats, pre_tanh=pre_tanh, drop=drop)
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.pre_tanh, ctx.drop = pre_tanh, drop
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, stats = ctx.saved_tensors
        if torch.is_grad_enabled() and (x.requires_grad or dy.requires_grad or gamma.requires_grad):
            # double-backward path.  When only dx is wanted (the WGAN-GP penalty differentiates the critic's INPUT
            # gradient: no parameter gradients in the first backward) it is one fused kernel whose own backward is the
            # closed-form dlsg_norm_bwd2; otherwise a differentiable restatement in torch ops.
            if x.shape[-1] <= 1024 and not SECOND_ORDER_THROUGH_NORM_PARAMS:
                return _norm_bwd_fused_diff(x, gamma, dy, stats, ctx.pre_tanh, ctx.drop) + (None, None)
            return _norm_bwd_diff(x, gamma, dy, ctx.pre_tanh, ctx.drop) + (None, None)
        x2 = _c(x).view(-1, x.shape[-1])
        d2 = _c(dy).view(-1, x.shape[-1])
        dx = torch.empty_like(x2)
        dg, db = la.small_zeros(gamma.shape, gamma), la.small_zeros(beta.shape, beta)
        ops.backend().norm_bwd(d2, x2, gamma.detach(), beta.detach(), stats, dx=dx, dgamma=dg, dbeta=db,
                               pre_tanh=ctx.pre_tanh, drop=ctx.drop)
        return dx.view(x.shape), dg, db, None, None


# The fused differentiable LayerNorm backward supports second derivatives through dx only (what the WGAN-GP penalty needs:
# it differentiates the critic's INPUT gradient).  Set True to route through the all-torch restatement instead, which is
# differentiable through the parameter gradients as well.
SECOND_ORDER_THROUGH_NORM_PARAMS = False


class _NormBwdCore(torch.autograd.Function):
    """(dt, dgamma, dbeta) = LayerNorm backward of dy at the normalised input t: forward = the fused row kernel;
    backward (for a cotangent of dt) = dlsg_norm_bwd2 (closed form)."""

    @staticmethod
    def forward(ctx, dy, t, gamma, stats):
        D = t.shape[-1]
        t2, d2 = _c(t).view(-1, D), _c(dy).view(-1, D)
        dt = torch.empty_like(t2)
        g0 = gamma.detach()
        dg, db = la.small_zeros(gamma.shape, gamma), la.small_zeros(gamma.shape, gamma)
        ops.backend().norm_bwd(d2, t2, g0, g0, stats, dx=dt, dgamma=dg, dbeta=db)
        ctx.save_for_backward(d2, t2, gamma, stats)
        ctx.set_materialize_grads(False)
        return dt.view(t.shape), dg, db

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u, u_gamma, u_beta):
        if u_gamma is not None or u_beta is not None:
            raise NotImplementedError('second derivative through LayerNorm parameter gradients: set '
                                      'dlsg.generic.SECOND_ORDER_THROUGH_NORM_PARAMS = True')
        if u is None:
            return None, None, None, None
        d2, t2, gamma, stats = ctx.saved_tensors
        u2 = _c(u).view(t2.shape)
        g_dy, g_t = torch.empty_like(t2), torch.empty_like(t2)
        g_gamma = la.small_zeros(gamma.shape, gamma) if ctx.needs_input_grad[2] else None
        ops.backend().norm_bwd2(t2, d2, u2, gamma.detach(), stats, g_dy=g_dy, g_x=g_t, g_gamma=g_gamma)
        return g_dy.view(u.shape), g_t.view(u.shape), g_gamma, None


def _norm_bwd_fused_diff(x, gamma, dy, stats, pre_tanh, drop):
    if drop is not None:
        dy = dropout_mask_apply(dy, drop)
    t = tanh_(x) if pre_tanh else x
    dt, dgamma, dbeta = _NormBwdCore.apply(dy, t, gamma, stats)
    if pre_tanh:
        dt = _TanhBwd.apply(dt, t)
    return dt, dgamma, dbeta


def _norm_bwd_diff(x, gamma, dy, pre_tanh, drop):
    if drop is not None:
        dy = dropout_mask_apply(dy, drop)
    t = tanh_(x) if pre_tanh else x
    D = x.shape[-1]
    mean = t.mean(-1, keepdim=True)
    var = ((t - mean) ** 2).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    xh = (t - mean) * rstd
    dgamma = (dy * xh).reshape(-1, D).sum(0)
    dbeta = dy.reshape(-1, D).sum(0)
    d = dy * gamma
    dt = rstd * (d - d.mean(-1, keepdim=True) - xh * (d * xh).mean(-1, keepdim=True))
    if pre_tanh:
        dt = dt * (1 - t * t)
    return dt, dgamma, dbeta


def norm(x, gamma, beta, pre_tanh=False, p_drop=0.0):
    drop = None
    if p_drop > 0:
        from .functional import next_seed
        drop = (float(p_drop), next_seed(), 0)
    return _Norm.apply(x, gamma, beta, pre_tanh, drop)


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, drop):
        ctx.drop = drop
        y = torch.empty_like(_c(x))
        ops.backend().dropout(_c(x), y, drop)
        return y

    @staticmethod
    def backward(ctx, dy):
        return _Dropout.apply(dy, ctx.drop), None


def dropout_mask_apply(x, drop):
    return _Dropout.apply(x, drop)


def dropout(x, p):
    if p <= 0:
        return x
    from .functional import next_seed
    return _Dropout.apply(x, (float(p), next_seed(), 0))


def add_pe(x, pe, p):
    y = x + pe[:, :x.size(1)]
    return dropout(y, p)


class _ColumnGather(torch.autograd.Function):
    """out[n, :] = w[:, ids[n]]  for a (D, V) weight: the token embedding of a ONE-HOT input under a k=1 Conv1d / Linear
    (model.py:147 on run_gun.py:449-453's to_onehot: a (N,V) x (V,D) product that is a pure column gather).
    Forward: one transposing copy of the weight (V,D) + the embedding-gather kernel; backward: the scatter-add kernel."""

    @staticmethod
    def forward(ctx, w, ids):
        D, V = w.shape
        wt = empty((V, D), w)
        ops.backend().convert(w.detach(), dstT=wt)
        out = empty((ids.shape[0], D), w)
        ops.backend().embedding_gather(wt, ids, out)
        ctx.save_for_backward(ids)
        ctx.shape = (D, V)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        D, V = ctx.shape
        dwt = zeros((V, D), dout)
        ops.backend().embedding_scatter_add(dwt, ids, _c(dout))
        return dwt.t(), None


def column_gather(w, ids):
    """w (D,V) (any row pitch), ids (N,) int64 -> (N,D) = rows of w^T."""
    return _ColumnGather.apply(w, ids.reshape(-1))


# ----------------------------------------------------------------------------------------------- misc (D path)
def resblock_blc(x, w3, b3):
    """(B,L,C) layout: r = relu(x); r + 0.3 * conv1d_k3_pad1(r)   (sublayer.py:117-119 with the in-place ReLU)."""
    r = torch.relu(x)
    rp = torch.nn.functional.pad(r, (0, 0, 1, 1))
    conv = (linear(rp[:, :-2], w3[:, :, 0]) + linear(rp[:, 1:-1], w3[:, :, 1]) + linear(rp[:, 2:], w3[:, :, 2]) + b3)
    return r + 0.3 * conv


def resblock(x_bcl, w3, b3):
    return resblock_blc(x_bcl.transpose(1, 2), w3, b3).transpose(1, 2)


class _MmNTW(torch.autograd.Function):
    """y = a @ w^T with a ready-made GEMM operand copy `w_op` of the weight (bf16 in bf16 mode; a transposed view of it
    serves the data-gradient product in place), so a recurrent weight is converted once per sequence, not per step."""

    @staticmethod
    def forward(ctx, a, w, w_op):
        ctx.save_for_backward(a, w)
        ctx.w_op = w_op
        out = torch.empty(a.shape[:-1] + (w_op.shape[-2],), dtype=torch.float32, device=a.device)
        ops.backend().gemm(la.op_cached(a), w_op, out)
        return out

    @staticmethod
    def backward(ctx, dy):
        a, w = ctx.saved_tensors
        da = dw = None
        if ctx.needs_input_grad[0]:
            da = _MmNTW.apply(dy, w.transpose(-1, -2), ctx.w_op.transpose(-1, -2))
        if ctx.needs_input_grad[1]:
            dw = bmm_nt(dy.transpose(-1, -2), a.transpose(-1, -2))
        return da, dw, None


class _Cell(torch.autograd.Function):
    """LSTM cell pointwise (gate order i,f,g,o): (pre (B,4H), c_prev (B,H)) -> (h, c, acts), one fused kernel.  The
    backward is `_CellBwd` (one fused kernel) whose own backward is the closed-form dlsg_lstm_cell_bwd2 kernel, so the
    WGAN-GP double backward through the discriminator LSTM costs one launch per step and level."""

    @staticmethod
    def forward(ctx, pre, c_prev):
        acts = _c(pre).clone()
        c = torch.empty_like(c_prev)
        h = torch.empty_like(c_prev)
        ops.backend().lstm_cell_fwd(acts, _c(c_prev), c, h_out=h)
        ctx.save_for_backward(pre, c_prev)
        ctx.acts, ctx.c = acts, c
        ctx.mark_non_differentiable(acts)
        return h, c, acts

    @staticmethod
    def backward(ctx, dh, dc, _dacts):
        pre, c_prev = ctx.saved_tensors
        return _CellBwd.apply(dh, dc, pre, c_prev, ctx.acts, ctx.c.detach())


class _CellBwd(torch.autograd.Function):
    """(dh, dc_next; pre, c_prev) -> (dpre, dc_prev).  `acts`/`c` are the forward's saved buffers (functions of pre and
    c_prev: their dependence is folded into the closed-form second-order terms)."""

    @staticmethod
    def forward(ctx, dh, dc, pre, c_prev, acts, c):
        dh = _c(dh) if dh is not None else torch.zeros_like(c)
        dc = _c(dc) if dc is not None else None
        dpre = torch.empty_like(acts)
        dc_prev = torch.empty_like(c)
        ops.backend().lstm_cell_bwd(acts, _c(c_prev), c, dh, dc, dc_prev, dgates=dpre)
        ctx.save_for_backward(dh, dc)
        ctx.acts, ctx.c, ctx.c_prev = acts, c, c_prev.detach()
        return dpre, dc_prev

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u, w):
        dh, dc = ctx.saved_tensors
        acts, c = ctx.acts, ctx.c
        g_dh, g_dc = torch.empty_like(c), torch.empty_like(c)
        g_pre, g_c0 = torch.empty_like(acts), torch.empty_like(c)
        ops.backend().lstm_cell_bwd2(acts, _c(ctx.c_prev), c, dh, dc, _c(u) if u is not None else None,
                                     _c(w) if w is not None else None, g_dh, g_dc, g_pre, g_c0)
        return g_dh, (g_dc if dc is not None else None), g_pre, g_c0, None, None


def _lstm_bptt_diff(gin, w_hh, dhs, need_dgin, need_dw):
    """Differentiable BPTT (used only when the gradient itself must be differentiated: the WGAN-GP penalty).  The
    forward is recomputed step by step with `_Cell` so that the result is a differentiable function of gin, w_hh, dhs."""
    B, T, H4 = gin.shape
    H = H4 // 4
    w_op = la.op(w_hh.detach())
    c = gin.new_zeros(B, H)
    h = None
    pres, cprevs, actss, cbufs, hs = [], [], [], [], []
    gin_t = gin.unbind(1)                     # one unbind / stack pair instead of T select nodes (each of whose backward
    dhs_t = dhs.unbind(1)                     # would materialise a full zero-padded (B,T,.) tensor)
    for t in range(T):
        pre = gin_t[t] if h is None else gin_t[t] + _MmNTW.apply(h, w_hh, w_op)
        pres.append(pre)
        cprevs.append(c)
        h, c, acts = _Cell.apply(pre, c)
        actss.append(acts)
        cbufs.append(c.detach())
        hs.append(h)
    dh_rec = dc = None
    dpres = [None] * T
    w_t, w_op_t = w_hh.transpose(-1, -2), w_op.transpose(-1, -2)
    for t in range(T - 1, -1, -1):
        dh = dhs_t[t] if dh_rec is None else dhs_t[t] + dh_rec
        dpres[t], dc = _CellBwd.apply(dh, dc, pres[t], cprevs[t], actss[t], cbufs[t])
        if t > 0:
            dh_rec = _MmNTW.apply(dpres[t], w_t, w_op_t)
    dgin = torch.stack(dpres, 1) if need_dgin else None
    dw = None
    if need_dw and T > 1:
        dp = torch.stack(dpres[1:], 1).reshape(-1, H4)              # rows (b, t>=1)
        hp = torch.stack(hs[:-1], 1).reshape(-1, H)                 # h fed into step t
        dw = bmm_nt(dp.transpose(0, 1), hp.transpose(0, 1))         # time-batched recurrent weight gradient (4H,H)
    elif need_dw:
        dw = torch.zeros_like(w_hh)
    return dgin, dw


# Under create_graph the LSTM's backward is the fused twice-differentiable node `_LstmBptt2` (second derivatives through dgin
# only: what the WGAN-GP penalty needs); False routes through the step-by-step restatement `_lstm_bptt_diff`, which is
# differentiable through the recurrent weight gradient as well.
FUSED_LSTM_BPTT2 = True
# Forward loop of the generic LSTM on the one-launch step kernel (recurrent product + cell) when the shape allows it.
FUSED_LSTM_STEP = True
# Small per-sample (batched) products on the fp32 FFMA kernel (measurement switch DLSG_SMALL_BMM_SIMT; measured slower)
SMALL_BMM_SIMT = int(os.environ.get('DLSG_SMALL_BMM_SIMT', '0'))      # 0 off (kept), 1 short reductions only, 2 every small product
# The second-order reverse loop over the forward (loop B of `_LstmBptt2`) rides on the first-order backward loop of `_LstmSeq`.
MERGE_LSTM_REVERSE_LOOPS = True


class _LstmBptt2(torch.autograd.Function):
    """The BPTT of the zero-state LSTM as ONE twice-differentiable node: (gin, w_hh, dhs) -> (dgin, dw_hh), reusing the
    buffers `_LstmSeq.forward` saved (activated gates, cell states, h operands: nothing is recomputed).
    forward = the fused first-order loop (cell backward + recurrent data-gradient GEMM per step) that also keeps every
    step's total dh and incoming dc;
    backward (for a cotangent U of dgin: the WGAN-GP penalty, run_gun.py:362-375) = two more fused loops,
      A (t ascending, the reverse of the BPTT recurrence): u_t = U_t + g_dh(t-1) W^T, then the closed-form
        dlsg_lstm_cell_bwd2 -> cotangents of dh_t (= of dhs_t), of dc, and the injections g_pre_t / g_c0_t into the
        forward's pre-activations / cell states;
      B (t descending, an ordinary BPTT over the same forward with those injections): -> cotangent of gin - by default NOT
        run here: it rides on the first-order backward loop of the forward node (MERGE_LSTM_REVERSE_LOOPS);
    and two time-batched weight-gradient GEMMs.  2 launches per step and loop, no torch arithmetic.  A cotangent of
    dw_hh (second derivative through the recurrent weight gradient) is not supported: `_lstm_bptt_diff` covers it."""

    @staticmethod
    def forward(ctx, gin, w_hh, dhs, hs, bufs, need_dw, link):
        be = ops.backend()
        acts, cs, hin, w_op = bufs
        ctx.link = link
        T, B, H4 = acts.shape
        H = H4 // 4
        dhs = _c(dhs)
        bf = la.precision() == 'bf16'
        dg32 = empty((B, T, H4), gin)                     # batch-major like gin: returned as it is
        dg_op = la.op_empty((T, B), H4, gin) if bf else empty((T, B, H4), gin)   # time-major GEMM operand copy
        Sd = la.splitk_rows(B, H, H4)
        dhrec = zeros((Sd, B, H), gin)
        dcs = empty((T + 1, B, H), gin)                   # dcs[t+1] enters step t (none at T-1), dcs[t] leaves it
        dht = empty((T, B, H), gin)                       # dh_t = dhs_t + recurrent part
        w_t = w_op.transpose(-1, -2)
        for t in range(T - 1, -1, -1):
            be.lstm_cell_bwd(acts[t], cs[t], cs[t + 1], dhs[:, t], dcs[t + 1] if t + 1 < T else None, dcs[t], dgates=dg32[:, t],
                             dgates2=dg_op[t], dh2=dhrec, dh_total=dht[t])
            if t > 0:
                be.gemm(dg_op[t], w_t, dhrec if Sd > 1 else dhrec[0], splitk=Sd)
        dw = la.mm(la.flat2(dg_op).t(), la.flat2(hin).t()) if need_dw else None
        ctx.save_for_backward(gin, w_hh)
        ctx.bufs, ctx.mine = bufs, (dg_op, dcs, dht)
        ctx.set_materialize_grads(False)
        return dg32, dw

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, U, u_w):
        if u_w is not None:
            raise NotImplementedError('second derivative through the recurrent weight gradient: use _lstm_bptt_diff')
        if U is None:
            return None, None, None, None, None, None, None
        be = ops.backend()
        gin, w_hh = ctx.saved_tensors
        acts, cs, hin, w_op = ctx.bufs
        dg_op, dcs, dht = ctx.mine
        T, B, H4 = acts.shape
        H = H4 // 4
        bf = la.precision() == 'bf16'
        if U.stride(2) != 1:
            U = U.contiguous()
        S = la.splitk_rows(B, H4, H)
        up = empty((S, B, H4), gin)                       # split-K partials of g_dh(t-1) W^T
        gdh = empty((B, T, H), gin)                       # batch-major like dhs: returned as it is
        gdh_in = la.op_zeros((T, B), H, gin)              # slot t: g_dh(t-1) as a GEMM operand (slot 0 = 0)
        gdc = empty((T, B, H), gin)
        gpre = empty((T, B, H4), gin)
        gc0 = empty((T, B, H), gin)
        for t in range(T):
            if t > 0:
                be.gemm(gdh_in[t], w_op, up if S > 1 else up[0], splitk=S)
            be.lstm_cell_bwd2(acts[t], cs[t], cs[t + 1], dht[t], dcs[t + 1] if t + 1 < T else None, U[:, t],
                              gdc[t - 1] if t > 0 else None, gdh[:, t], gdc[t], gpre[t], gc0[t], u2=(up if t > 0 else None),
                              g_dh2=(gdh_in[t + 1] if t + 1 < T else None))
        if ctx.link is not None and MERGE_LSTM_REVERSE_LOOPS:
            # Loop B is an ordinary BPTT over the same forward as the first-order backward of `_LstmSeq`, which is linear in its
            # inputs and which autograd runs AFTER this node (`hs`, that node's output, is a formal input here: the edge puts
            # it into every backward pass that reaches this node, behind it): hand it loop A's injections and let ONE loop
            # produce the sum of both gradients.
            link = ctx.link
            link['inject'] = (gpre, gc0, dg_op, gdh_in)

            def check():
                if link.pop('inject', None) is not None:
                    raise RuntimeError('dlsg.generic._LstmBptt2: the LSTM forward node did not run its backward in this pass, so the '
                                       'second-order gradient of its inputs was lost; set dlsg.generic.MERGE_LSTM_REVERSE_LOOPS = False')
            torch.autograd.Variable._execution_engine.queue_callback(check)
            return None, None, gdh, None, None, None, None
        a32 = empty((B, T, H4), gin)
        a_op = la.op_empty((T, B), H4, gin) if bf else empty((T, B, H4), gin)
        Sd = la.splitk_rows(B, H, H4)
        hrec = zeros((Sd, B, H), gin)
        zero_dh = zeros((B, H), gin)
        dc, dc2 = empty((B, H), gin), empty((B, H), gin)
        w_t = w_op.transpose(-1, -2)
        for t in range(T - 1, -1, -1):
            last = t == T - 1
            be.lstm_cell_bwd(acts[t], cs[t], cs[t + 1], zero_dh, None if last else dc, dc2, dgates=a32[:, t],
                             dgates2=a_op[t], dh2=hrec, dc_next2=(None if last else gc0[t + 1]), dgates_add=gpre[t])
            dc, dc2 = dc2, dc
            if t > 0:
                be.gemm(a_op[t], w_t, hrec if Sd > 1 else hrec[0], splitk=Sd)
        g_w = None
        if ctx.needs_input_grad[1]:
            g_w = la.mm(la.flat2(a_op).t(), la.flat2(hin).t())                 # sum_t a_pre_t^T h(t-1)
            la.mm(la.flat2(dg_op).t(), la.flat2(gdh_in).t(), out=g_w, accum=True)   # + sum_t dpre_t^T g_dh(t-1)
        return a32, g_w, gdh, None, None, None, None


class _LstmSeq(torch.autograd.Function):
    """Zero-state uni-directional LSTM over all steps: gin (B,T,4H) = x W_ih^T + b_ih + b_hh -> h (B,T,H).
    Forward: the one-launch step kernel (recurrent product + cell; up to four 64-row groups sharing the weight) or, outside
    its range, recurrent GEMM + cell kernel per step.  First-order backward: cell backward + data-gradient GEMM per step
    (split-K partials summed by the cell kernel), one time-batched weight-gradient GEMM; it also carries the injections of
    a pending second-order pass (`_LstmBptt2.backward`).  Under create_graph the backward is the twice-differentiable node
    `_LstmBptt2` (or the step-by-step restatement `_lstm_bptt_diff` when FUSED_LSTM_BPTT2 is off)."""

    @staticmethod
    def forward(ctx, gin, w_hh):
        be = ops.backend()
        gin = _c(gin)
        B, T, H4 = gin.shape
        H = H4 // 4
        w_op = la.op(w_hh.detach())
        hs = empty((B, T, H), gin)
        ng = (B + 63) // 64                                # 64-row groups for the one-launch step kernel
        if (FUSED_LSTM_STEP and la.precision() == 'bf16' and ng <= 4 and B % ng == 0 and w_op.is_contiguous()
                and be.lstm_step_supported(B // ng, H)):
            # recurrent product + cell in ONE launch per step (csrc/lstm_step.cu; the row groups share the weight)
            sl = [slice(g * (B // ng), (g + 1) * (B // ng)) for g in range(ng)]
            acts = empty((T, B, H4), gin)
            cs = empty((T + 1, B, H), gin)
            cs[0].zero_()
            hin = la.op_empty((T, B), H, gin)
            hin[0].zero_()
            for t in range(T):
                be.lstm_step_fwd([w_op] * ng, [hin[t][s_] for s_ in sl] if t > 0 else None, [gin[s_, t] for s_ in sl],
                                 [cs[t][s_] for s_ in sl] if t > 0 else None, [cs[t + 1][s_] for s_ in sl], [acts[t][s_] for s_ in sl],
                                 h_out=[hs[s_, t] for s_ in sl], h_op=([hin[t + 1][s_] for s_ in sl] if t + 1 < T else None))
            ctx.save_for_backward(gin, w_hh, hs)
            ctx.bufs = (acts, cs, hin, w_op)
            return hs
        S = la.splitk_for(B, H4, H)
        gates = empty((S, T, B, H4), gin)                 # split-K partials; [0] ends up holding the activated gates
        gates[:, 0].zero_()                               # (step 0 has no recurrent product; later steps are overwritten by theirs)
        cs = empty((T + 1, B, H), gin)
        cs[0].zero_()
        hin = la.op_zeros((T, B), H, gin)                 # h fed INTO step t, as a GEMM operand
        for t in range(T):
            if t > 0:
                be.gemm(hin[t], w_op, gates[:, t] if S > 1 else gates[0, t], splitk=S)
            be.lstm_cell_fwd(gates[:, t], cs[t], cs[t + 1], row_bias=gin[:, t], h2=hs[:, t],
                             h3=(hin[t + 1] if t + 1 < T else None))
        ctx.save_for_backward(gin, w_hh, hs)
        ctx.bufs = (gates[0], cs, hin, w_op)
        return hs

    @staticmethod
    def backward(ctx, dhs):
        gin, w_hh, hs = ctx.saved_tensors
        need_dgin, need_dw = ctx.needs_input_grad
        if torch.is_grad_enabled():
            if FUSED_LSTM_BPTT2:
                if not hasattr(ctx, 'link'):
                    ctx.link = {}
                dgin, dw = _LstmBptt2.apply(gin, w_hh, dhs, hs, ctx.bufs, need_dw and not _INPUT_GRADS_ONLY, ctx.link)
                return (dgin if need_dgin else None), (dw if need_dw else None)
            return _lstm_bptt_diff(gin, w_hh, dhs, need_dgin, need_dw)
        be = ops.backend()
        acts, cs, hin, w_op = ctx.bufs
        T, B, H4 = acts.shape
        H = H4 // 4
        dhs = _c(dhs)
        bf = la.precision() == 'bf16'
        dg32 = empty((B, T, H4), gin)                     # batch-major like gin: returned as it is
        dg_op = la.op_empty((T, B), H4, gin) if bf else empty((T, B, H4), gin)   # time-major GEMM operand copy
        Sd = la.splitk_rows(B, H, H4)
        dhrec = zeros((Sd, B, H), gin)
        dc, dc2 = empty((B, H), gin), empty((B, H), gin)
        w_t = w_op.transpose(-1, -2)
        inj = ctx.link.pop('inject', None) if hasattr(ctx, 'link') else None     # second-order injections (see _LstmBptt2.backward)
        gpre, gc0 = (inj[0], inj[1]) if inj is not None else (None, None)
        for t in range(T - 1, -1, -1):
            be.lstm_cell_bwd(acts[t], cs[t], cs[t + 1], dhs[:, t], dc if t + 1 < T else None, dc2, dgates=(dg32[:, t] if need_dgin else None),
                             dgates2=dg_op[t], dh2=dhrec, dc_next2=(gc0[t + 1] if (inj is not None and t + 1 < T) else None),
                             dgates_add=(gpre[t] if inj is not None else None))
            dc, dc2 = dc2, dc
            if t > 0:
                be.gemm(dg_op[t], w_t, dhrec if Sd > 1 else dhrec[0], splitk=Sd)
        dw = None
        if need_dw:
            dw = la.mm(la.flat2(dg_op).t(), la.flat2(hin).t())      # sum_t dgates_t^T h_{t-1}  (hin[0] = 0)
            if inj is not None:
                la.mm(la.flat2(inj[2]).t(), la.flat2(inj[3]).t(), out=dw, accum=True)      # + sum_t dpre_t^T g_dh(t-1)
        return (dg32 if need_dgin else None), dw


def lstm(x, w_ih, w_hh, b_ih, b_hh):
    """Uni-directional nn.LSTM(batch_first) with zero initial state (model.py:152), differentiable twice."""
    return _LstmSeq.apply(linear(x, w_ih, b_ih + b_hh), w_hh)


def sum_dim1(x):
    return x.sum(dim=1)


def topk_indices(x, k):
    return torch.topk(x, k, -1)[1]


def gather_rows(p, idx):
    return torch.gather(p, 1, idx.unsqueeze(-1).expand(idx.shape[0], idx.shape[1], p.shape[-1]))


def weighted_mean_score(s, adj_alpha, groups=1):
    """layer.py:713-714: per-sample alpha-weighted node score, then the batch mean -> 0-dim scalar.  groups > 1: the
    batch is `groups` independent calls stacked along dim 0 (dlsg.gan batches D(real), D(fake), D(mixed)): one mean
    per group -> (groups,)."""
    per = (s * adj_alpha).sum(-1) / adj_alpha.sum(-1)
    if groups == 1:
        return per.mean(-1)
    return per.view(groups, -1).mean(-1)


def fuse_scores(s_obj, s_mot, f, groups=1):
    if groups > 1:
        n = f.shape[0] // groups
        s_obj, s_mot = s_obj.repeat_interleave(n), s_mot.repeat_interleave(n)
    return s_obj * f[:, 0] + s_mot * f[:, 1]


def embedding(ids, table):
    out = empty(tuple(ids.shape) + (table.shape[1],), table)
    ops.backend().embedding_gather(table.detach(), ids.reshape(-1), out=out.view(-1, table.shape[1]))
    return out
