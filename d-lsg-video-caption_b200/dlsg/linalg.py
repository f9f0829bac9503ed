"""Precision policy + GEMM-operand plumbing shared by every block.

precision 'bf16' (default): dense projections run on the tcgen05 kernel with bf16 operands and fp32
accumulation; operands are K-major (unit stride on the reduction axis) or MN-major (transposed views), 16-byte
aligned, pitch a multiple of 8 elements.  `op()` materialises a cast copy (dlsg_convert2d) only when a tensor is
not already usable; transposed copies are never needed.
precision 'fp32': every product runs on the strided fp32 FFMA kernel on the original tensors
(strict-parity mode: logits within 1e-4 of the fp32 reference, bit-exact token ids).
Tiny fp32 x fp32 products (26x26 attention, P<=8 pooling) use the FFMA kernel in both modes.
"""
import os

import torch

from . import ops

_PRECISION = os.environ.get('DLSG_PRECISION', 'bf16')


def set_precision(p):
    global _PRECISION
    assert p in ('bf16', 'fp32')
    _PRECISION = p


def precision():
    return _PRECISION


def opdtype():
    return torch.bfloat16 if _PRECISION == 'bf16' else torch.float32


def ceil8(n):
    return (n + 7) // 8 * 8


def empty(shape, like, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device=like.device)


def zeros(shape, like, dtype=torch.float32):
    return torch.zeros(shape, dtype=dtype, device=like.device)


_POOL = None


class ZeroPool:
    """One zero-filled fp32 slab per block backward, handed out in 16-byte aligned slices: the ~50 small parameter-
    gradient accumulators of a step (LayerNorm gamma/beta, biases, theta) cost ONE fill launch per block instead of
    one each."""

    def __init__(self, device, numel=1 << 16):
        self.buf = torch.zeros(numel, dtype=torch.float32, device=device)
        self.off = 0

    def take(self, shape):
        n = 1
        for s_ in shape:
            n *= s_
        if self.off + n > self.buf.numel():
            return None
        v = self.buf[self.off:self.off + n].view(shape)
        self.off += (n + 3) // 4 * 4
        return v


def begin_pool(device):
    global _POOL
    _POOL = ZeroPool(device)


def end_pool():
    global _POOL
    _POOL = None


def small_zeros(shape, like):
    """fp32 zeros for a small gradient accumulator: a slice of the active ZeroPool when a block backward is running."""
    shape = tuple(shape)
    if _POOL is not None and _POOL.buf.device == like.device:
        v = _POOL.take(shape)
        if v is not None:
            return v
    return torch.zeros(shape, dtype=torch.float32, device=like.device)


def op_empty(rows_shape, K, like):
    """Operand buffer (..., K) in the operand dtype with an aligned row pitch; returns the [..., :K] view."""
    buf = torch.empty(tuple(rows_shape) + (ceil8(K),), dtype=opdtype(), device=like.device)
    return buf[..., :K]


def op_zeros(rows_shape, K, like):
    buf = torch.zeros(tuple(rows_shape) + (ceil8(K),), dtype=opdtype(), device=like.device)
    return buf[..., :K]


def flat2(x):
    """Collapse the leading dims of a (possibly column-padded) operand buffer view into rows."""
    rows = 1
    for s_ in x.shape[:-1]:
        rows *= s_
    return x.as_strided((rows, x.shape[-1]), (x.stride(-2), 1), x.storage_offset())


def _tc_ok(t):
    """bf16, 16-byte aligned, and either K-major (unit stride on the last axis) or MN-major (a transposed view: unit
    stride on the second-last axis) with every other pitch a multiple of 8 elements - the tcgen05 GEMM reads both
    in place (MN-major through 64x64 TMA boxes and the UMMA MN-major descriptor)."""
    if t.dtype != torch.bfloat16 or t.data_ptr() % 16 != 0:
        return False
    if t.stride(-1) == 1 or t.shape[-1] == 1:
        unit = t.dim() - 1
    elif t.dim() >= 2 and (t.stride(-2) == 1 or t.shape[-2] == 1):
        unit = t.dim() - 2
    else:
        return False
    for d in range(t.dim()):
        if d != unit and t.shape[d] > 1 and t.stride(d) % 8 != 0:
            return False
    return True


def op(t):
    """Return `t` (2-D or batched 3-D, reduction axis last) as a GEMM operand for the current precision.  A transposed
    view stays a transposed view (of a cast copy when it is not bf16 yet): no transposed copy is ever made."""
    if _PRECISION == 'fp32':
        if t.dtype == torch.float32:
            return t
        out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
        _convert_into(t, out)
        return out
    if _tc_ok(t):
        return t
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        if t.dim() >= 2 and t.stride(-2) == 1:
            tt = t.transpose(-1, -2)                       # the row-major tensor underneath: cast it, hand back the view
            out = op_empty(tt.shape[:-1], tt.shape[-1], t)
            ops.backend().convert(tt, dst=out)
            return out.transpose(-1, -2)
        t = t.contiguous()          # e.g. one tap of a conv1d weight (C_out, C_in, k)[:, :, j]
    out = op_empty(t.shape[:-1], t.shape[-1], t)
    _convert_into(t, out)
    return out


# ---- memoised operands for the generic (autograd-composed) path -----------------------------------------------------
# The discriminator path is thousands of small GEMMs whose operands are immutable autograd values and parameters; the
# same tensor typically feeds several products (dy -> data gradient AND weight gradient; a weight -> forward, backward
# and double backward).  op_cached() converts each once.  Entries hold a reference to the source tensor (so ids are
# never reused while an entry lives) and are validated by tensor._version; parameter entries additionally carry an
# epoch, bumped at the start of every top-level generic-path module forward, because fused optimizers update
# parameters without bumping _version.
import collections as _collections

_PMEMO = {}
_AMEMO = _collections.OrderedDict()
_AMEMO_MAX = 256
_EPOCH = [0]
_SCOPE_DEPTH = [0]
_MANUAL_EPOCH = [False]


def new_param_epoch():
    _EPOCH[0] += 1
    _AMEMO.clear()


def set_manual_param_epochs(flag):
    """True: the caller (dlsg.gan.GanIteration) bumps the epoch itself right after each optimizer step, so a weight is
    converted once per optimizer step instead of once per module forward."""
    _MANUAL_EPOCH[0] = bool(flag)


class param_epoch_scope:
    def __enter__(self):
        if _SCOPE_DEPTH[0] == 0 and not _MANUAL_EPOCH[0]:
            new_param_epoch()
        _SCOPE_DEPTH[0] += 1

    def __exit__(self, *exc):
        _SCOPE_DEPTH[0] -= 1
        return False


def op_cached(t):
    """op() with memoisation - ONLY for tensors nobody writes in place through a raw pointer (autograd values of the
    generic path, parameters)."""
    if _PRECISION != 'bf16' or t.dtype == torch.bfloat16 or t.dim() < 2:
        return op(t)
    if t.stride(-1) != 1 and t.shape[-1] != 1 and t.stride(-2) == 1:
        return op_cached(t.transpose(-1, -2)).transpose(-1, -2)      # one copy serves both orientations
    root = t._base if t._base is not None else t
    key = (id(root), t.data_ptr(), tuple(t.shape), tuple(t.stride()))
    if isinstance(root, torch.nn.Parameter):
        ver = (root._version, _EPOCH[0])
        hit = _PMEMO.get(key)
        if hit is not None and hit[0] is root and hit[1] == ver:
            return hit[2]
        out = op(t)
        _PMEMO[key] = (root, ver, out)
        return out
    hit = _AMEMO.get(key)
    if hit is not None and hit[0] is root and hit[1] == root._version:
        _AMEMO.move_to_end(key)
        return hit[2]
    out = op(t)
    _AMEMO[key] = (root, root._version, out)
    while len(_AMEMO) > _AMEMO_MAX:
        _AMEMO.popitem(last=False)
    return out


def _convert_into(t, out):
    be = ops.backend()
    if t.stride(-1) == 1 or t.shape[-1] == 1:
        be.convert(t, dst=out)
    else:
        assert t.stride(-2) == 1, 'operand needs a unit stride on one of its last two axes'
        be.convert(t.transpose(-1, -2), dstT=out)


def mm(a, b, out=None, out_dtype=torch.float32, memo=False, **epi):
    """out = epi(a @ b^T); a (..,M,K), b (..,N,K).  Operands are made precision-appropriate via op() (memo=True: via
    op_cached(), generic path only)."""
    a2, b2 = (op_cached(a), op_cached(b)) if memo else (op(a), op(b))
    if out is None:
        out = torch.empty(a.shape[:-1] + (b.shape[-2],), dtype=out_dtype, device=a.device)
    ops.backend().gemm(a2, b2, out, **epi)
    return out


def mm32(a, b, out=None, **epi):
    """Small fp32 product on the strided FFMA kernel regardless of precision mode."""
    if out is None:
        out = torch.empty(a.shape[:-1] + (b.shape[-2],), dtype=torch.float32, device=a.device)
    ops.backend().gemm(a, b, out, **epi)
    return out


def splitk_rows(rows, n_out, K, sms=148, cap=4):
    """Split-K factor for a recurrent GEMM with MORE than 64 rows (the critic's stacked LSTM: 192 rows, dlsg.generic): few
    128 x 128 tiles and a long K loop.  Same rule as the kernel's automatic split (gemm_tc.cu), but explicit, so that the
    consuming cell kernel sums the partials instead of a reduce launch per step; capped at what its vector path sums."""
    if rows <= 64:
        return splitk_for(rows, n_out, K, sms)
    if precision() != 'bf16':
        return 1
    tiles = ((rows + 127) // 128) * ((n_out + 127) // 128)
    kb = (K + 63) // 64
    if tiles * 2 > sms or kb < 16:
        return 1
    s = min(sms // tiles, kb // 8, cap)
    per = (kb + s - 1) // s
    return max(1, (kb + per - 1) // per)


class WeightCache:
    """bf16 copies of parameters (one per weight; the transposed operand is a view of it) and packed forms.

    Validity is NOT keyed on tensor._version alone: fused optimizers (torch.optim.Adam(fused=True)) update parameters
    without bumping it.  Entries carry the *scope* they were built in:
      ('train', g) - built during training-block forward number g; reused only by that block's own backward;
      ('eval', g)  - built under no_grad after g training forwards; reused by later no_grad calls (all decode steps)
                     until the next training forward.
    So every training forward re-converts the weights it uses (they changed since the last optimizer step), inference
    converts once."""

    def __init__(self):
        self._c = {}
        self.force = False      # True while a CUDA graph is being captured: refresh kernels must be recorded every time
        self.gen = 0
        self.scope = ('eval', 0)
        self._rec = None        # list of (src, src2, dst) while a forward is being recorded (record())
        self._pinned = None     # PinnedWeights in effect: lookups are served from its snapshot, no conversion launches

    # ---- conversions that can be recorded and later replayed as ONE multi-segment launch ---------------------------
    def cv(self, src, dst):
        """dst = cast(src) for a 2-D parameter view (the only way the cache builders convert weights)."""
        ops.backend().convert(src, dst=dst)
        if self._rec is not None:
            self._rec.append((src, None, dst))

    def sum2(self, a, b, dst):
        """dst = a + b (fp32 vectors: the two LSTM bias vectors, summed once per weight update)."""
        be = ops.backend()
        be.axpby(a, 1.0, dst, 0.0)
        be.axpby(b, 1.0, dst, 1.0)
        if self._rec is not None:
            self._rec.append((a.view(1, -1), b.view(1, -1), dst.view(1, -1)))

    def record(self, fn):
        """Run fn() (a training forward) with a cleared cache, logging every weight conversion and the cache entries it
        creates; returns a PinnedWeights that can refresh all of them in one launch."""
        assert self._pinned is None and self._rec is None
        self._c.clear()
        self._rec = []
        try:
            fn()
            rec, snap = self._rec, dict(self._c)
        finally:
            self._rec = None
        return PinnedWeights(rec, snap)

    def pin(self, pinned):
        self._pinned = pinned

    def unpin(self):
        self._pinned = None

    def begin_train_block(self):
        self.gen += 1
        self.scope = ('train', self.gen)
        return self.scope

    def eval_scope(self):
        self.scope = ('eval', self.gen)
        return self.scope

    def set_scope(self, scope):
        self.scope = scope

    def get(self, w, transpose=False, key=None):
        if _PRECISION == 'fp32':
            return w.detach().t() if transpose else w.detach()
        if transpose:
            return self.get(w, key=key).t()         # MN-major view of the same bf16 copy (read in place by the GEMM)
        k = (id(w) if key is None else key, False)
        if self._pinned is not None:
            hit = self._pinned.snap.get(k)
            if hit is not None and hit[0][1] == w.data_ptr():
                return hit[1]
        ver = (w._version, w.data_ptr(), tuple(w.shape), self.scope)
        hit = self._c.get(k)
        if hit is not None and hit[0] == ver and not self.force:
            return hit[1]
        src = w.detach()
        out = op_empty((src.shape[0],), src.shape[1], src)
        self.cv(src, out)
        self._c[k] = (ver, out)
        return out

    def packed(self, key, parts, versions, builder):
        """Cache an arbitrary packed operand (e.g. concatenated LSTM weights) keyed on part versions."""
        if self._pinned is not None:
            hit = self._pinned.snap.get(key)
            if hit is not None and hit[0][0] == _PRECISION and tuple(v[1] for v in hit[0][2:]) == tuple(v[1] for v in versions):
                return hit[1]
        ver = (_PRECISION, self.scope) + tuple(versions)
        hit = self._c.get(key)
        if hit is not None and hit[0] == ver and not self.force:
            return hit[1]
        val = builder()
        self._c[key] = (ver, val)
        return val

    def clear(self):
        self._c.clear()


class PinnedWeights:
    """Snapshot of the operand copies one training forward uses, plus the conversion list that produced them: refresh()
    re-runs ALL of them as one multi-segment launch (after an optimizer step / before the next forward).  While pinned
    (WeightCache.pin) the blocks are served these buffers and launch no conversion of their own."""

    def __init__(self, rec, snap):
        self.snap = snap
        self.pairs = [(s_.detach(), (s2.detach() if s2 is not None else None), d) for s_, s2, d in rec]
        self.plan = ops.backend().make_convert_plan(self.pairs) if self.pairs else None

    def refresh(self):
        if self.plan is not None:
            ops.backend().multi_convert(self.plan)


def pver(*params):
    return tuple((p._version, p.data_ptr()) for p in params)


def splitk_for(rows, n_out, K, sms=148):
    """Split-K factor for a skinny recurrent GEMM (rows <= 64 ride the UMMA N side, weights fill 128-row tiles): enough
    CTAs to cover `sms` SMs (148 = the whole GPU; 74 when two such chains run concurrently on two streams), >= 4
    k-blocks (of 64) per split, no empty split.  1 when the batch is not skinny."""
    if rows > 64 or precision() != 'bf16':
        return 1
    tiles = (n_out + 127) // 128
    kb = (K + 63) // 64
    if tiles * 2 > sms or kb < 8:
        return 1
    s = min(sms // tiles, kb // 4, 16)          # floor: tiles*s CTAs fit in one wave of the persistent grid
    per = (kb + s - 1) // s
    return max(1, (kb + per - 1) // per)
