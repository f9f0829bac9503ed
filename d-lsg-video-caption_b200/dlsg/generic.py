"""Generic differentiable primitives over libdlsg kernels (used by the discriminator path and by the
stand-alone sub-layer forwards).  Each primitive is a torch.autograd.Function whose backward is itself
written with these primitives, so gradients of gradients (WGAN-GP, run_gun.py:362-371) work.
"""
import math

import torch

from . import linalg as la
from . import ops
from .linalg import empty, zeros


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------- matmul
class _MmNT(torch.autograd.Function):
    """y[..., M, N] = a[..., M, K] @ b[..., N, K]^T (batched when 3-D)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return la.mm(a.detach(), b.detach())

    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        da = db = None
        if ctx.needs_input_grad[0]:
            da = bmm_nt(dy, b.transpose(-1, -2))          # dy @ b
        if ctx.needs_input_grad[1]:
            db = bmm_nt(dy.transpose(-1, -2), a.transpose(-1, -2))   # dy^T @ a
        return da, db


def bmm_nt(a, b):
    """a (..,M,K) @ b (..,N,K)^T with identical leading batch dims (2-D or 3-D)."""
    return _MmNT.apply(a, b)


def matmul_nn(a, b):
    return bmm_nt(a, b.transpose(-1, -2))


def linear(x, w, b=None, tanh=False):
    """nn.Linear over the last dim of x (any leading dims)."""
    lead = x.shape[:-1]
    y = bmm_nt(x.reshape(-1, x.shape[-1]), w)
    if b is not None:
        y = y + b
    if tanh:
        y = tanh_(y)
    return y.view(*lead, w.shape[0])


class _Tanh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = torch.tanh(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return dy * (1 - y * y)


def tanh_(x):
    return _Tanh.apply(x)


def mul(a, b):
    return a * b


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


def _softmax_bwd(x, dy, dim, scale, mask, mask_mode):
    """dx of softmax; differentiable again (expressed with torch elementwise ops on the recomputed softmax)."""
    if torch.is_grad_enabled() and (x.requires_grad or dy.requires_grad):
        s = _Softmax.apply(x, dim, scale, mask, 1 if mask_mode == 1 else 0)
        g = dy if mask_mode != 2 else torch.where(mask > 0, dy, torch.zeros_like(dy))
        r = scale * s * (g - (g * s).sum(dim, keepdim=True))
        if mask_mode == 1:
            r = torch.where(mask > 0, r, torch.zeros_like(r))
        return r
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
            # double-backward path: differentiable restatement on the saved statistics
            return _norm_bwd_diff(x, gamma, dy, ctx.pre_tanh, ctx.drop) + (None, None)
        x2 = _c(x).view(-1, x.shape[-1])
        d2 = _c(dy).view(-1, x.shape[-1])
        dx = torch.empty_like(x2)
        dg, db = zeros(gamma.shape, gamma), zeros(beta.shape, beta)
        ops.backend().norm_bwd(d2, x2, gamma.detach(), beta.detach(), stats, dx=dx, dgamma=dg, dbeta=db,
                               pre_tanh=ctx.pre_tanh, drop=ctx.drop)
        return dx.view(x.shape), dg, db, None, None


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


# ----------------------------------------------------------------------------------------------- misc (D path)
def resblock_blc(x, w3, b3):
    """(B,L,C) layout: r = relu(x); r + 0.3 * conv1d_k3_pad1(r)   (sublayer.py:117-119 with the in-place ReLU)."""
    r = torch.relu(x)
    rp = torch.nn.functional.pad(r, (0, 0, 1, 1))
    conv = (linear(rp[:, :-2], w3[:, :, 0]) + linear(rp[:, 1:-1], w3[:, :, 1]) + linear(rp[:, 2:], w3[:, :, 2]) + b3)
    return r + 0.3 * conv


def resblock(x_bcl, w3, b3):
    return resblock_blc(x_bcl.transpose(1, 2), w3, b3).transpose(1, 2)


def lstm(x, w_ih, w_hh, b_ih, b_hh):
    """Uni-directional nn.LSTM(batch_first) with zero initial state (model.py:152), differentiable twice."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    gin = linear(x, w_ih, b_ih + b_hh)
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    outs = []
    for t in range(T):
        g = gin[:, t] + (bmm_nt(h, w_hh) if t > 0 else 0)
        i, f, gg, o = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), tanh_(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
        c = f * c + i * gg
        h = o * tanh_(c)
        outs.append(h)
    return torch.stack(outs, 1)


def sum_dim1(x):
    return x.sum(dim=1)


def topk_indices(x, k):
    return torch.topk(x, k, -1)[1]


def gather_rows(p, idx):
    return torch.gather(p, 1, idx.unsqueeze(-1).expand(idx.shape[0], idx.shape[1], p.shape[-1]))


def weighted_mean_score(s, adj_alpha):
    return ((s * adj_alpha).sum(-1) / adj_alpha.sum(-1)).mean(-1)


def fuse_scores(s_obj, s_mot, f):
    return s_obj * f[:, 0] + s_mot * f[:, 1]


def embedding(ids, table):
    out = empty(tuple(ids.shape) + (table.shape[1],), table)
    ops.backend().embedding_gather(table.detach(), ids.reshape(-1), out=out.view(-1, table.shape[1]))
    return out
