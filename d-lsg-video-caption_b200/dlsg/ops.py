"""Tensor-level wrappers over the libdlsg C-ABI (one Python call = one kernel launch).

All arguments are torch CUDA tensors / strided views; outputs are written in place into
caller-provided tensors (nothing is allocated here, so every call is CUDA-graph capturable).
`backend` is the CUDA library; tests may install a CPU emulation of these primitives
(tests/cpu_emul.py) to exercise the host orchestration without a GPU - the product never does.
"""
import ctypes as C
import os

import torch

from . import _lib as L

F32, BF16 = L.F32, L.BF16
# DLSG_STATIC_PREFETCH=1 requests the early weight fetch (DLSG_GEMM_B_STATIC) of the recurrent GEMMs.  OFF by default: measured
# on B200 it does not change the captured step (6.55 vs 6.55 ms with a 6-stage ring, 6.64 vs 6.63 with 8 stages,
# gpurun r02w): behind the ~1.3 us CTA prologue the fetch wins nothing the L2-resident weights had not already given.
STATIC_PREFETCH = os.environ.get('DLSG_STATIC_PREFETCH', '0') == '1'


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError('unsupported dtype %s' % t.dtype)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rows2d(t):
    """(rows, cols, ld) of a 2-D (or flattenable leading dims) view with unit inner stride."""
    assert t.stride(-1) == 1 or t.shape[-1] == 1, 'inner dim must be contiguous: %s %s' % (t.shape, t.stride())
    if t.dim() == 1:
        return 1, t.shape[0], t.shape[0]
    if t.dim() == 2:
        return t.shape[0], t.shape[1], t.stride(0)
    # leading dims must collapse to a uniform row stride
    ld = t.stride(-2)
    rows = 1
    exp = ld
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] != 1:
            assert t.stride(d) == exp, 'leading dims not collapsible: %s %s' % (t.shape, t.stride())
        exp *= t.shape[d]
        rows *= t.shape[d]
    return rows, t.shape[-1], ld


class CudaBackend:
    name = 'cuda'

    WORKSPACE_BYTES = 64 << 20

    def __init__(self):
        self.lib = L.load()
        self._launches = 0
        self._writer_last = {}          # stream handle -> the last libdlsg launch on it wrote GEMM-operand copies of weights
        self._ws = {}

    # Every wrapper counts its launch with `self.launches += 1`; the setter also notes, per stream, that the newest launch is
    # an ordinary kernel.  The wrappers that WRITE operand copies of weights (convert / multi_convert / adam_multi) then mark
    # themselves with _mark_writer().  gemm(b_static=True) asks for the early weight fetch only when the launch directly
    # before it on its stream is not such a writer: with programmatic dependent launch a kernel may start while its direct
    # predecessor still runs (never earlier: every kernel triggers its dependents after its own dependency wait), so that is
    # the one case in which the early fetch could read a weight copy that is still being written.
    @property
    def launches(self):
        return self._launches

    @launches.setter
    def launches(self, v):
        self._launches = v
        if STATIC_PREFETCH and torch.cuda.is_available():    # (the tracking only matters when the early fetch is requested)
            self._writer_last[_stream()] = False

    def _mark_writer(self):
        if STATIC_PREFETCH:
            self._writer_last[_stream()] = True

    def _workspace(self, dev):
        """Per-device scratch for automatic split-K (allocated once, before any graph capture uses it)."""
        key = (dev, torch.cuda.current_stream(dev).cuda_stream)      # concurrent streams must not share scratch
        ws = self._ws.get(key)
        if ws is None:
            ws = torch.empty(self.WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
            self._ws[key] = ws
        return ws

    def _ck(self, t):
        if not t.is_cuda:
            raise L.DlsgError('dlsg ops need CUDA tensors (no CPU fallback)')

    # ------------------------------------------------------------------ GEMM
    def gemm(self, a, b, out, bias=None, bias_axis='n', tanh=False, alpha=1.0, accum=False, splitk=1,
             impl=None, atomic=False, b_static=False):
        """out[(batch,)M,N] = epi(alpha * a[(batch,)M,K] @ b[(batch,)N,K]^T + bias).
        atomic=True: out += result through fp32 atomic adds (DLSG_EPI_ATOMIC): skinny problems split K with no reduce launch.
        b_static=True: promise that `b` (a weight) is not written by the kernel launched just before this one on the stream
        (DLSG_GEMM_B_STATIC): its first tiles are fetched while that kernel still runs.  Same results.

        `out` may be a transposed view (unit stride on M) -> STORE_T.  For splitk>1 `out` has a
        leading split dim: (splitk, M, N) and receives partial sums."""
        self._ck(a)
        g = L.GemmT()
        if splitk > 1:
            assert a.dim() == 2 and out.dim() == 3 and out.shape[0] == splitk
            g.stride_split = out.stride(0)
            o2 = out[0]
        else:
            o2 = out
        if a.dim() == 3:
            batch = a.shape[0]
            assert b.dim() == 3 and o2.dim() == 3
            g.stride_a, g.stride_b, g.stride_d = a.stride(0), b.stride(0), o2.stride(0)
            a2, b2, o2 = a[0], b[0], o2[0]
        else:
            batch = 1
            a2, b2 = a, b
        M, K = a2.shape
        N, K2 = b2.shape
        assert K == K2 and tuple(o2.shape) == (M, N), (a.shape, b.shape, out.shape)
        flags = 0
        if o2.stride(1) == 1 or N == 1:
            g.ldd = o2.stride(0)
        else:
            assert o2.stride(0) == 1, 'out must have a unit stride'
            g.ldd = o2.stride(1)
            flags |= L.EPI_STORE_T
        if bias is not None:
            flags |= L.EPI_BIAS_N if bias_axis == 'n' else L.EPI_BIAS_M
        if tanh:
            flags |= L.EPI_TANH
        if accum:
            flags |= L.EPI_ACCUM
        if atomic:
            assert not accum and not tanh and splitk <= 1 and o2.dtype == torch.float32
            flags |= L.EPI_ATOMIC
        if b_static and STATIC_PREFETCH and self._writer_last.get(_stream(), True) is False:
            flags |= L.GEMM_B_STATIC
        g.A, g.B, g.D, g.bias = a2.data_ptr(), b2.data_ptr(), o2.data_ptr(), _ptr(bias)
        g.M, g.N, g.K, g.batch = M, N, K, batch
        g.sam, g.sak, g.sbn, g.sbk = a2.stride(0), a2.stride(1), b2.stride(0), b2.stride(1)
        g.a_dtype, g.b_dtype, g.d_dtype = _dt(a2), _dt(b2), _dt(o2)
        if impl is None:
            # tcgen05 path: bf16 operands, each K-major (unit K stride) or MN-major (a transposed view: unit row stride)
            tc_a = g.sak == 1 or K == 1 or g.sam == 1 or M == 1
            tc_b = g.sbk == 1 or K == 1 or g.sbn == 1 or N == 1
            impl = L.GEMM_TC if (g.a_dtype == BF16 and g.b_dtype == BF16 and tc_a and tc_b) else L.GEMM_SIMT
        g.impl, g.flags, g.splitk, g.alpha = impl, flags, splitk, alpha
        if impl == L.GEMM_TC and splitk <= 1:
            ws = self._workspace(a.device)
            g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel()
        self.launches += 1
        L.check(self.lib.dlsg_gemm(C.byref(g), _stream()), 'dlsg_gemm')

    # ------------------------------------------------------------------ conversion
    def convert(self, src, dst=None, dstT=None):
        """dst[..., r, c] = src[..., r, c]; dstT[..., c, r] = src[..., r, c] (2-D or batched 3-D views)."""
        self._ck(src)
        ref = dst if dst is not None else dstT
        if src.dim() == 3:
            batch, rows, cols = src.shape
            bs = (src.stride(0), dst.stride(0) if dst is not None else 0, dstT.stride(0) if dstT is not None else 0)
            s2, d2, t2 = src[0], (dst[0] if dst is not None else None), (dstT[0] if dstT is not None else None)
        else:
            batch, bs = 1, (0, 0, 0)
            rows, cols = src.shape
            s2, d2, t2 = src, dst, dstT
        assert s2.stride(1) == 1 or cols == 1
        self.launches += 1
        self._mark_writer()
        L.check(self.lib.dlsg_convert2d_batched(
            s2.data_ptr(), _dt(s2), s2.stride(0), _ptr(d2), _dt(ref), d2.stride(0) if d2 is not None else 0,
            _ptr(t2), t2.stride(0) if t2 is not None else 0, rows, cols, batch, bs[0], bs[1], bs[2], _stream()),
            'dlsg_convert2d')

    def make_convert_plan(self, pairs, chunk_elems=16384, host=False):
        """pairs: list of (src, src2 | None, dst) 2-D tensor views (fp32 sources with one common pitch per pair, or a bf16
        source without src2; unit inner strides).  Returns a plan (tables + references that keep the buffers alive) for
        multi_convert().  host=False: device-resident tables (uploaded here: not inside a graph capture); host=True: the
        table stays on the host and rides in the kernel parameters of every launch (capturable with fresh pointers)."""
        import numpy as np
        segs = (L.SegT * max(1, len(pairs)))()
        chunks = []
        keep = []
        for i, (src, src2, dst) in enumerate(pairs):
            assert src.dim() == 2 and dst.shape == src.shape and (src.dtype == torch.float32 or src2 is None), (src.shape, dst.shape, src.dtype)
            assert (src.stride(1) == 1 or src.shape[1] == 1) and (dst.stride(1) == 1 or dst.shape[1] == 1)
            rows, cols = src.shape
            sg = segs[i]
            sg.src, sg.dst, sg.rows, sg.cols = src.data_ptr(), dst.data_ptr(), rows, cols
            sg.ld_src, sg.ld_dst = (src.stride(0) if rows > 1 else cols), (dst.stride(0) if rows > 1 else cols)
            sg.src_dtype, sg.dst_dtype = _dt(src), _dt(dst)
            if src2 is not None:
                assert src2.shape == src.shape and src2.dtype == torch.float32 and (rows == 1 or src2.stride(0) == src.stride(0))
                sg.src2 = src2.data_ptr()
            keep.append((src, src2, dst))
            if host:
                continue
            per = max(1, chunk_elems // max(1, cols))
            for r0 in range(0, rows, per):
                chunks.append((i, r0, min(per, rows - r0)))
        if host:
            return {'table': segs, 'n': len(pairs), 'chunk_elems': chunk_elems, 'keep': keep, 'host': True}
        dev = pairs[0][0].device
        seg_t = torch.from_numpy(np.frombuffer(bytes(segs), dtype=np.uint8).copy()).to(dev)
        chunk_t = torch.tensor(chunks, dtype=torch.int32).reshape(-1).to(dev)
        return {'segs': seg_t, 'chunks': chunk_t, 'n': len(chunks), 'keep': keep}

    def multi_convert(self, plan):
        if plan['n'] == 0:
            return
        if plan.get('host'):
            self.launches += (plan['n'] + 383) // 384
            self._mark_writer()
            L.check(self.lib.dlsg_multi_convert_host(C.cast(plan['table'], C.c_void_p), plan['n'], plan['chunk_elems'], _stream()),
                    'dlsg_multi_convert_host')
            return
        self.launches += 1
        self._mark_writer()
        L.check(self.lib.dlsg_multi_convert(plan['segs'].data_ptr(), plan['chunks'].data_ptr(), plan['n'], _stream()),
                'dlsg_multi_convert')

    def make_adam_plan(self, segs, chunk_elems=16384):
        """segs: list of dicts {p, m, v: 2-D fp32 views with one common pitch, g: fp32 | bf16 view of the same shape (own
        pitch), dst: bf16 2-D view | None, step: optional device scalar with this parameter's own step count}.  The plan is a
        host table: it rides to the device inside the kernel parameters (no upload; graph-capturable as is)."""
        table = (L.AdamSegT * max(1, len(segs)))()
        for i, sg in enumerate(segs):
            p_, g_, m_, v_, d_ = sg['p'], sg['g'], sg['m'], sg['v'], sg.get('dst')
            rows, cols = p_.shape
            ld = p_.stride(0) if rows > 1 else cols
            for t_ in (p_, m_, v_):
                assert t_.dtype == torch.float32 and t_.shape == p_.shape and (cols == 1 or t_.stride(1) == 1)
                assert rows == 1 or t_.stride(0) == ld, 'p, m, v of a segment must share one pitch'
            assert g_.shape == p_.shape and (cols == 1 or g_.stride(1) == 1), 'gradient view must match the parameter segment'
            e = table[i]
            e.p, e.g, e.m, e.v = p_.data_ptr(), g_.data_ptr(), m_.data_ptr(), v_.data_ptr()
            e.rows, e.cols, e.ld = rows, cols, ld
            e.ld_g, e.g_dtype = (g_.stride(0) if rows > 1 else cols), _dt(g_)
            if sg.get('step') is not None:
                assert sg['step'].dtype == torch.float32 and sg['step'].numel() == 1
                e.step = sg['step'].data_ptr()
            if d_ is not None:
                assert d_.dtype == torch.bfloat16 and d_.shape == p_.shape and (cols == 1 or d_.stride(1) == 1)
                e.dst16, e.ld_dst = d_.data_ptr(), (d_.stride(0) if rows > 1 else cols)
        return {'table': table, 'n': len(segs), 'chunk_elems': chunk_elems, 'keep': segs}

    def adam_multi(self, plan, step, lr, beta1, beta2, eps, lr_dev=None):
        if plan['n'] == 0:
            return
        assert step.dtype == torch.float32
        self.launches += (plan['n'] + 255) // 256
        self._mark_writer()
        L.check(self.lib.dlsg_adam_multi(C.cast(plan['table'], C.c_void_p), plan['n'], plan['chunk_elems'], step.data_ptr(),
                                         _ptr(lr_dev), float(lr), float(beta1), float(beta2), float(eps), _stream()), 'dlsg_adam_multi')

    def colsum(self, x, out):
        rows, cols, ld = _rows2d(x)
        self.launches += 1
        L.check(self.lib.dlsg_colsum(x.data_ptr(), _dt(x), ld, rows, cols, out.data_ptr(), _stream()), 'dlsg_colsum')

    # ------------------------------------------------------------------ norm family
    def norm_fwd(self, x, gamma, beta, y=None, y2=None, res=None, stats=None, pre_tanh=False, post_tanh=False,
                 drop=None):
        self._ck(x)
        p = L.NormFwdT()
        rows, D, p.ldx = _rows2d(x)
        p.x, p.x_dtype, p.rows, p.D = x.data_ptr(), _dt(x), rows, D
        if res is not None:
            r_, d_, p.ldres = _rows2d(res)
            assert (r_, d_) == (rows, D)
            p.res, p.res_dtype = res.data_ptr(), _dt(res)
        if y is not None:
            r_, d_, p.ldy = _rows2d(y)
            assert (r_, d_) == (rows, D), (y.shape, x.shape)
            p.y, p.y_dtype = y.data_ptr(), _dt(y)
        if y2 is not None:
            r_, d_, p.ldy2 = _rows2d(y2)
            assert (r_, d_) == (rows, D)
            p.y2, p.y2_dtype = y2.data_ptr(), _dt(y2)
        p.gamma, p.beta, p.stats = gamma.data_ptr(), beta.data_ptr(), _ptr(stats)
        p.flags = (L.NORM_PRE_TANH if pre_tanh else 0) | (L.NORM_POST_TANH if post_tanh else 0)
        if drop is not None and drop[0] > 0:
            p.drop_p, p.seed, p.offset = drop
        self.launches += 1
        L.check(self.lib.dlsg_norm_fwd(C.byref(p), _stream()), 'dlsg_norm_fwd')

    def norm_bwd(self, dy, x, gamma, beta, stats, dx=None, res=None, dgamma=None, dbeta=None, pre_tanh=False,
                 post_tanh=False, in_is_tanh=False, drop=None, dx_accum=False, dxsum=None):
        """dxsum (D) += column sums of dx (the bias gradient of the Linear that produced x): fused into the streaming bf16
        kernel when the call is eligible, otherwise one extra colsum launch."""
        self._ck(x)
        p = L.NormBwdT()
        rows, D, p.ldx = _rows2d(x)
        p.x, p.x_dtype, p.rows, p.D = x.data_ptr(), _dt(x), rows, D
        r_, d_, p.lddy = _rows2d(dy)
        assert (r_, d_) == (rows, D), (dy.shape, x.shape)
        p.dy, p.dy_dtype = dy.data_ptr(), _dt(dy)
        if res is not None:
            _, _, p.ldres = _rows2d(res)
            p.res, p.res_dtype = res.data_ptr(), _dt(res)
        if dx is not None:
            r_, d_, p.lddx = _rows2d(dx)
            assert (r_, d_) == (rows, D)
            p.dx, p.dx_dtype = dx.data_ptr(), _dt(dx)
        p.gamma, p.beta, p.stats = gamma.data_ptr(), beta.data_ptr(), stats.data_ptr()
        p.dgamma, p.dbeta = _ptr(dgamma), _ptr(dbeta)
        p.flags = ((L.NORM_PRE_TANH if pre_tanh else 0) | (L.NORM_POST_TANH if post_tanh else 0) |
                   (L.NORM_IN_IS_TANH if in_is_tanh else 0))
        p.dx_accum = 1 if dx_accum else 0
        if drop is not None and drop[0] > 0:
            p.drop_p, p.seed, p.offset = drop
        self.launches += 1
        fused_sum = dxsum is not None and self.lib.dlsg_norm_bwd_streaming(C.byref(p)) == 1
        if fused_sum:
            p.dxsum = dxsum.data_ptr()
        L.check(self.lib.dlsg_norm_bwd(C.byref(p), _stream()), 'dlsg_norm_bwd')
        if dxsum is not None and not fused_sum:
            self.colsum(dx, dxsum)

    def norm_bwd2(self, x, dy, u, gamma, stats, g_dy=None, g_x=None, g_gamma=None):
        """Backward of the plain LayerNorm backward w.r.t. (dy, x, gamma) for a cotangent u of dx; contiguous fp32 (rows, D)."""
        self._ck(x)
        p = L.NormBwd2T()
        rows, D, _ = _rows2d(x)
        for t_ in (x, dy, u, g_dy, g_x):
            assert t_ is None or (t_.is_contiguous() and t_.dtype == torch.float32 and t_.numel() == rows * D)
        p.x, p.dy, p.u, p.gamma, p.stats = x.data_ptr(), dy.data_ptr(), u.data_ptr(), gamma.data_ptr(), stats.data_ptr()
        p.g_dy, p.g_x, p.g_gamma, p.rows, p.D = _ptr(g_dy), _ptr(g_x), _ptr(g_gamma), rows, D
        self.launches += 1
        L.check(self.lib.dlsg_norm_bwd2(C.byref(p), _stream()), 'dlsg_norm_bwd2')

    # ------------------------------------------------------------------ LSTM cell
    def lstm_cell_fwd(self, gates, c_prev, c_out, h_out=None, row_bias=None, bias=None, h2=None, h3=None, drop=None):
        """gates: (B,4H) or (nsplit,B,4H) fp32 (overwritten with activated i,f,g,o in gates[0])."""
        self._ck(gates)
        p = L.CellFwdT()
        self._fill_cell_fwd(p, gates, c_prev, c_out, h_out, row_bias, bias, h2, h3, drop)
        self.launches += 1
        L.check(self.lib.dlsg_lstm_cell_fwd(C.byref(p), _stream()), 'dlsg_lstm_cell_fwd')

    @staticmethod
    def fused_step_supported(H):
        return H % 4 == 0 and H <= 2048

    def lstm_cell_norm_fwd(self, gates, c_prev, c_out, gamma, beta, y, h_out=None, row_bias=None, bias=None, h2=None, h3=None,
                           drop=None, y2=None, stats=None, post_tanh=False, ydrop=None):
        """Fused cell + LayerNorm: y = [tanh](LN(h)) (+dropout); h itself goes to h_out/h2/h3 as in lstm_cell_fwd."""
        self._ck(gates)
        q = L.CellNormFwdT()
        self._fill_cell_norm_fwd(q, gates, c_prev, c_out, gamma, beta, y, h_out, row_bias, bias, h2, h3, drop, y2, stats, post_tanh, ydrop)
        self.launches += 1
        L.check(self.lib.dlsg_lstm_cell_norm_fwd(C.byref(q), _stream()), 'dlsg_lstm_cell_norm_fwd')

    def _fill_cell_norm_fwd(self, q, gates, c_prev, c_out, gamma, beta, y, h_out, row_bias, bias, h2, h3, drop, y2, stats, post_tanh, ydrop):
        self._fill_cell_fwd(q.cell, gates, c_prev, c_out, h_out, row_bias, bias, h2, h3, drop)
        q.gamma, q.beta, q.stats = gamma.data_ptr(), beta.data_ptr(), _ptr(stats)
        q.y, q.ldy, q.y_dtype = y.data_ptr(), y.stride(0), _dt(y)
        if y2 is not None:
            q.y2, q.ldy2, q.y2_dtype = y2.data_ptr(), y2.stride(0), _dt(y2)
        q.post_tanh = 1 if post_tanh else 0
        if ydrop is not None and ydrop[0] > 0:
            q.ydrop_p, q.yseed, q.yoffset = ydrop

    def cell_norm_attn2_fwd(self, cell, attn):
        """The query LSTM's cell + LayerNorm and the hoisted attention step of one decode step in ONE launch
        (dlsg_cell_norm_attn2_fwd).  cell: dict of lstm_cell_norm_fwd's arguments; attn: dict of attn2_fwd's arguments (its
        `q` must be the cell part's `y`: the kernel takes it from there).  Returns False (nothing launched) when the shape
        is outside the fused kernel's range - the caller then issues the two launches."""
        self._ck(cell['gates'])
        f = L.CellNormAttn2FwdT()
        c = dict(h_out=None, row_bias=None, bias=None, h2=None, h3=None, drop=None, y2=None, stats=None, post_tanh=False, ydrop=None)
        c.update(cell)
        assert attn['q'].data_ptr() == c['y'].data_ptr() and c['y'].dtype == torch.float32 and not c['post_tanh']
        self._fill_cell_norm_fwd(f.cn, c['gates'], c['c_prev'], c['c_out'], c['gamma'], c['beta'], c['y'], c['h_out'], c['row_bias'], c['bias'],
                                 c['h2'], c['h3'], c['drop'], c['y2'], c['stats'], c['post_tanh'], c['ydrop'])
        self._fill_attn2_fwd(f.at, **attn)
        if not self.lib.dlsg_cell_norm_attn2_supported(C.byref(f)):
            return False
        self.launches += 1
        L.check(self.lib.dlsg_cell_norm_attn2_fwd(C.byref(f), _stream()), 'dlsg_cell_norm_attn2_fwd')
        return True

    @staticmethod
    def _fill_cell_fwd(p, gates, c_prev, c_out, h_out, row_bias, bias, h2, h3, drop):
        if gates.dim() == 3:
            p.nsplit, p.stride_split = gates.shape[0], gates.stride(0)
            g0 = gates[0]
        else:
            p.nsplit, g0 = 1, gates
        assert g0.is_contiguous()
        B, H4 = g0.shape
        p.gates, p.B, p.H = g0.data_ptr(), B, H4 // 4
        if row_bias is not None:
            p.row_bias, p.ld_row_bias = row_bias.data_ptr(), row_bias.stride(0)
        p.bias, p.c_prev, p.c_out, p.h_out = _ptr(bias), _ptr(c_prev), c_out.data_ptr(), _ptr(h_out)
        if h2 is not None:
            p.h2, p.ldh2, p.h2_dtype = h2.data_ptr(), h2.stride(0), _dt(h2)
        if h3 is not None:
            p.h3, p.ldh3, p.h3_dtype = h3.data_ptr(), h3.stride(0), _dt(h3)
        if drop is not None and drop[0] > 0:
            p.drop_p, p.seed, p.offset = drop

    def lstm_cell_bwd(self, acts, c_prev, c_new, dh, dc_next, dc_prev, dgates=None, dgates2=None, dgatesT=None,
                      drop=None, dh2=None, dc_next2=None, dgates_add=None, dh_total=None):
        """dh / dh2: (B,H) fp32 views (unit inner stride); their sum is the gradient wrt the (dropped) h.
        dc_next2 (B,H) is added to dc_next, dgates_add (B,4H) to the gate gradients (injections of the second-order
        reverse pass), dh_total (B,H) receives dh + dh2; all three contiguous fp32."""
        self._ck(acts)
        p = L.CellBwdT()
        self._fill_cell_bwd(p, acts, c_prev, c_new, dh, dc_next, dc_prev, dgates, dgates2, dgatesT, drop, dh2)
        for t_ in (dc_next2, dgates_add, dh_total):
            assert t_ is None or (t_.is_contiguous() and t_.dtype == torch.float32)
        p.dc_next2, p.dgates_add, p.dh_total = _ptr(dc_next2), _ptr(dgates_add), _ptr(dh_total)
        self.launches += 1
        L.check(self.lib.dlsg_lstm_cell_bwd(C.byref(p), _stream()), 'dlsg_lstm_cell_bwd')

    def lstm_cell_bwd2(self, acts, c_prev, c_new, dh, dc_next, u, w, g_dh, g_dc, g_pre, g_cprev, u2=None, g_dh2=None):
        """Backward of lstm_cell_bwd: cotangents u (of dgates) / w (of dc_prev) -> cotangents of dh, dc_next, the gate
        pre-activations and c_prev.  Contiguous fp32 (B,H) / (B,4H); c_prev, dc_next, u, w and outputs may be None.
        u2: (B,4H) or split-K partials (S,B,4H) added to u; g_dh2: second copy of g_dh (any dtype, own row pitch)."""
        self._ck(acts)
        p = L.CellBwd2T()
        for t_ in (acts, c_prev, c_new, dh, dc_next, w, g_dc, g_pre, g_cprev):
            assert t_ is None or (t_.is_contiguous() and t_.dtype == torch.float32)
        for t_ in (u, g_dh):                              # row-pitched (B,.) slices of batch-major (B,T,.) tensors are fine
            assert t_ is None or (t_.stride(1) == 1 and t_.dtype == torch.float32)
        if u is not None:
            p.ld_u = u.stride(0)
        if g_dh is not None:
            p.ld_g_dh = g_dh.stride(0)
        p.B, p.H = acts.shape[0], acts.shape[1] // 4
        p.acts, p.c_prev, p.c_new, p.dh, p.dc_next = acts.data_ptr(), _ptr(c_prev), c_new.data_ptr(), dh.data_ptr(), _ptr(dc_next)
        p.u, p.w = _ptr(u), _ptr(w)
        p.g_dh, p.g_dc, p.g_pre, p.g_cprev = _ptr(g_dh), _ptr(g_dc), _ptr(g_pre), _ptr(g_cprev)
        if u2 is not None:
            assert u2.dtype == torch.float32 and (u2[0] if u2.dim() == 3 else u2).is_contiguous()
            if u2.dim() == 3:
                p.u2_nsplit, p.u2_stride_split = u2.shape[0], u2.stride(0)
            p.u2 = u2.data_ptr()
        if g_dh2 is not None:
            assert g_dh2.stride(1) == 1
            p.g_dh2, p.ld_g_dh2, p.g_dh2_dtype = g_dh2.data_ptr(), g_dh2.stride(0), _dt(g_dh2)
        self.launches += 1
        L.check(self.lib.dlsg_lstm_cell_bwd2(C.byref(p), _stream()), 'dlsg_lstm_cell_bwd2')

    def norm_lstm_cell_bwd(self, acts, c_prev, c_new, dc_next, dc_prev, dy, x, gamma, beta, stats, dgamma, dbeta, dh=None, dh2=None,
                           dgates=None, dgates2=None, dgatesT=None, dgates_sum=None, drop=None, post_tanh=False, ydrop=None):
        """Fused LayerNorm backward (dy wrt y=[tanh](LN(x)), x = the cell's dropped h) + LSTM cell backward."""
        self._ck(acts)
        q = L.NormCellBwdT()
        assert dgates is None or dgates.is_contiguous(), 'the fused LayerNorm + cell backward writes contiguous dgates'
        self._fill_cell_bwd(q.cell, acts, c_prev, c_new, dh, dc_next, dc_prev, dgates, dgates2, dgatesT, drop, dh2)
        assert dy.stride(1) == 1 and x.stride(1) == 1
        q.dy, q.lddy, q.x, q.ldx = dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0)
        q.gamma, q.beta, q.stats, q.dgamma, q.dbeta = gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr()
        assert dgamma.shape == x.shape and dgamma.stride(0) == dbeta.stride(0), 'dgamma/dbeta are per-row (B,H) outputs'
        q.ld_dparam = dgamma.stride(0)
        q.dgates_sum = _ptr(dgates_sum)
        q.post_tanh = 1 if post_tanh else 0
        if ydrop is not None and ydrop[0] > 0:
            q.ydrop_p, q.yseed, q.yoffset = ydrop
        self.launches += 1
        L.check(self.lib.dlsg_norm_lstm_cell_bwd(C.byref(q), _stream()), 'dlsg_norm_lstm_cell_bwd')

    @staticmethod
    def _fill_cell_bwd(p, acts, c_prev, c_new, dh, dc_next, dc_prev, dgates, dgates2, dgatesT, drop, dh2):
        B, H4 = acts.shape
        p.acts, p.c_prev, p.c_new, p.dc_next = acts.data_ptr(), _ptr(c_prev), c_new.data_ptr(), _ptr(dc_next)
        if dh is not None:
            p.dh, p.lddh = dh.data_ptr(), dh.stride(0)
            assert dh.stride(1) == 1
        if dh2 is not None:
            if dh2.dim() == 3:                              # (S, B, H) split-K partial sums of the recurrent gradient
                p.dh2_nsplit, p.dh2_stride_split = dh2.shape[0], dh2.stride(0)
                dh2 = dh2[0]
            assert dh2.stride(1) == 1
            p.dh2, p.lddh2 = dh2.data_ptr(), dh2.stride(0)
        assert c_new.is_contiguous() and acts.is_contiguous()
        p.dgates, p.dc_prev, p.B, p.H = _ptr(dgates), _ptr(dc_prev), B, H4 // 4
        if dgates is not None:
            assert dgates.stride(1) == 1 and dgates.dtype == torch.float32
            p.ld_dgates = dgates.stride(0)
        if dgates2 is not None:
            p.dgates2, p.ld_dgates2, p.dgates2_dtype = dgates2.data_ptr(), dgates2.stride(0), _dt(dgates2)
        if dgatesT is not None:
            p.dgatesT, p.ld_dgatesT, p.dgatesT_dtype = dgatesT.data_ptr(), dgatesT.stride(0), _dt(dgatesT)
        if drop is not None and drop[0] > 0:
            p.drop_p, p.seed, p.offset = drop

    # ------------------------------------------------------------------ softmax
    @staticmethod
    def _softmax_desc(x, dim):
        dim = dim % x.dim()
        assert x.dim() == 3, 'softmax views are 3-D (outer, n, inner) after the caller reshapes'
        order = [d for d in range(3) if d != dim]
        o, i = order
        return x.shape[o], x.shape[dim], x.shape[i], x.stride(o), x.stride(dim), x.stride(i)

    @staticmethod
    def _same_layout(a, b):
        return a.shape == b.shape and all(sa == sb or n == 1 for sa, sb, n in zip(a.stride(), b.stride(), a.shape))

    def softmax_fwd(self, x, y, dim, scale=1.0, mask=None, mask_mode=0):
        self._ck(x)
        p = L.SoftmaxT()
        assert self._same_layout(x, y) and (mask is None or self._same_layout(x, mask))
        p.outer, p.n, p.inner, p.so, p.sn, p.si = self._softmax_desc(x, dim)
        p.x, p.y, p.mask, p.scale, p.mask_mode = x.data_ptr(), y.data_ptr(), _ptr(mask), scale, mask_mode
        self.launches += 1
        L.check(self.lib.dlsg_softmax_fwd(C.byref(p), _stream()), 'dlsg_softmax_fwd')

    def softmax_bwd(self, x, dy, dx, dim, scale=1.0, mask=None, mask_mode=0):
        self._ck(x)
        p = L.SoftmaxT()
        assert self._same_layout(x, dy) and self._same_layout(x, dx) and (mask is None or self._same_layout(x, mask))
        p.outer, p.n, p.inner, p.so, p.sn, p.si = self._softmax_desc(x, dim)
        p.x, p.mask, p.scale, p.mask_mode = x.data_ptr(), _ptr(mask), scale, mask_mode
        self.launches += 1
        L.check(self.lib.dlsg_softmax_bwd(C.byref(p), dy.data_ptr(), dx.data_ptr(), _stream()), 'dlsg_softmax_bwd')

    def softmax_bwd2(self, x, dy, u, dim, g_dy=None, g_x=None, scale=1.0, mask=None, mask_mode=0):
        """Backward of softmax_bwd wrt (dy, x) for a cotangent u of its dx (closed form, one launch)."""
        self._ck(x)
        p = L.SoftmaxT()
        for t_ in (dy, u, g_dy, g_x, mask):
            assert t_ is None or self._same_layout(x, t_)
        p.outer, p.n, p.inner, p.so, p.sn, p.si = self._softmax_desc(x, dim)
        p.x, p.mask, p.scale, p.mask_mode = x.data_ptr(), _ptr(mask), scale, mask_mode
        self.launches += 1
        L.check(self.lib.dlsg_softmax_bwd2(C.byref(p), dy.data_ptr(), u.data_ptr(), _ptr(g_dy), _ptr(g_x), _stream()), 'dlsg_softmax_bwd2')

    def ew(self, op, ins, outs, cols=0):
        """Small fused element-wise forms (dlsg_ew, op = dlsg._lib.EW_*): contiguous fp32 tensors of one size (the per-row
        operand of the LERP ops has n / cols elements); outputs may be None."""
        p = L.EwT()
        n = ins[0].numel()
        self._ck(ins[0])
        rows_e = {L.EW_LERP_ROWS: 2, L.EW_LERP_ROWS_BWD: 1}.get(op)
        for k, t_ in enumerate(ins):
            assert t_.is_contiguous() and t_.dtype == torch.float32
            assert t_.numel() == (n // cols if k == rows_e else n), (op, k, t_.shape)
            p.inp[k] = t_.data_ptr()
        for k, t_ in enumerate(outs):
            if t_ is not None:
                assert t_.is_contiguous() and t_.dtype == torch.float32 and t_.numel() == n
                p.out[k] = t_.data_ptr()
        p.n, p.cols, p.op = n, cols, op
        self.launches += 1
        L.check(self.lib.dlsg_ew(C.byref(p), _stream()), 'dlsg_ew')

    # ------------------------------------------------------------------ node attention
    def node_attn_fwd(self, Kp, Vp, qp, alpha, ctx, rows_per_node=1):
        """Kp,Vp (nh,nodes,P,H) fp32; qp (rows,nh*H) fp32; alpha (rows,nh*P) view; ctx (rows,nh*H) view."""
        self._ck(Kp)
        p = L.AttnFwdT()
        nh, nodes, P, H = Kp.shape
        assert Kp.is_contiguous() and Vp.is_contiguous() and qp.is_contiguous()
        p.Kp, p.Vp, p.qp, p.alpha, p.ctx = Kp.data_ptr(), Vp.data_ptr(), qp.data_ptr(), _ptr(alpha), ctx.data_ptr()
        p.rows, p.nh, p.P, p.H, p.rows_per_node, p.ctx_dtype = qp.shape[0], nh, P, H, rows_per_node, _dt(ctx)
        p.ldctx, p.ldalpha, p.nodes = ctx.stride(0), (alpha.stride(0) if alpha is not None else 0), nodes
        self.launches += 1
        L.check(self.lib.dlsg_node_attn_fwd(C.byref(p), _stream()), 'dlsg_node_attn_fwd')

    def node_attn_bwd(self, Kp, Vp, qp, alpha, dctx, dqp, dKp, dVp, dalpha_ext=None):
        self._ck(Kp)
        p = L.AttnBwdT()
        nh, nodes, P, H = Kp.shape
        assert nodes == qp.shape[0] and dqp.is_contiguous() and dKp.is_contiguous() and dVp.is_contiguous()
        p.Kp, p.Vp, p.qp, p.alpha, p.dctx = Kp.data_ptr(), Vp.data_ptr(), qp.data_ptr(), alpha.data_ptr(), dctx.data_ptr()
        p.dalpha_ext, p.dqp, p.dKp, p.dVp = _ptr(dalpha_ext), dqp.data_ptr(), dKp.data_ptr(), dVp.data_ptr()
        p.rows, p.nh, p.P, p.H, p.lddctx, p.ldalpha, p.dqp_dtype = qp.shape[0], nh, P, H, dctx.stride(0), alpha.stride(0), _dt(dqp)
        if dalpha_ext is not None:
            assert dalpha_ext.stride(0) == alpha.stride(0)
        self.launches += 1
        L.check(self.lib.dlsg_node_attn_bwd(C.byref(p), _stream()), 'dlsg_node_attn_bwd')

    # ------------------------------------------------------------------ hoisted attention
    @staticmethod
    def attn2_supported(nh, P, Hk, Hv):
        return 1 <= nh <= 2 and 1 <= P <= 8 and Hk <= 1024 and Hv <= 1024 and Hk % 4 == 0 and Hv % 4 == 0

    def attn2_fwd(self, KW, VW, q, alpha, co, scale, rows_per_node=1, ln=None):
        """KW (nh,nodes,P,Hk), VW (nh,nodes,P,Hv) fp32 contiguous; q (rows,Hk) view; alpha (rows,nh*P); co (rows,nh*Hv) view.

        ln = dict(gamma=[..nh], beta=[..nh], y=(rows,nh*Hv) view, stats=(nh,rows,2), drop=(p,seed,offset)|None,
        drop_head_stride=int) fuses the context output layer tanh -> LayerNorm -> dropout into the same kernel."""
        self._ck(KW)
        p = L.Attn2FwdT()
        self._fill_attn2_fwd(p, KW, VW, q, alpha, co, scale, rows_per_node, ln)
        self.launches += 1
        L.check(self.lib.dlsg_attn2_fwd(C.byref(p), _stream()), 'dlsg_attn2_fwd')

    @staticmethod
    def _fill_attn2_fwd(p, KW, VW, q, alpha, co, scale, rows_per_node=1, ln=None):
        nh, nodes, P, Hk = KW.shape
        assert KW.is_contiguous() and VW.is_contiguous() and q.stride(1) == 1 and co.stride(1) == 1
        p.KW, p.VW, p.q, p.alpha, p.co = KW.data_ptr(), VW.data_ptr(), q.data_ptr(), _ptr(alpha), co.data_ptr()
        p.ldq, p.ldalpha, p.ldco = q.stride(0), (alpha.stride(0) if alpha is not None else 0), co.stride(0)
        p.rows, p.nh, p.P, p.Hk, p.Hv, p.rows_per_node, p.nodes, p.scale = q.shape[0], nh, P, Hk, VW.shape[3], rows_per_node, nodes, scale
        if ln is not None:
            y, st = ln['y'], ln['stats']
            assert y.stride(1) == 1 and st.is_contiguous() and tuple(st.shape) == (nh, q.shape[0], 2)
            p.y, p.ldy, p.y_dtype = y.data_ptr(), y.stride(0), _dt(y)
            for k in range(nh):
                p.gamma[k], p.beta[k] = ln['gamma'][k].data_ptr(), ln['beta'][k].data_ptr()
            p.stats, p.stats_head_stride = st.data_ptr(), st.stride(0)
            if ln.get('drop') is not None:
                p.drop_p, p.seed, p.offset = ln['drop']
                p.offset_head_stride = ln['drop_head_stride']

    def attn2_bwd(self, KW, VW, q, alpha, dco, dq, dKW, dVW, scale, dalpha_ext=None, ln=None, save=None):
        """ln = dict(dy=(rows,nh*Hv) fp32 view, co=(rows,nh*Hv), gamma=[..nh], stats=(nh,rows,2), dgamma_rows, dbeta_rows
        (rows,nh*Hv) views, drop, drop_head_stride): the backward of the fused output layer produces dco in-kernel.
        save = (dl (rows, nh*P), dco (rows, nh*Hv)) fp32 views: record this step's d(logits) / d(co) instead of accumulating
        dKW / dVW (attn2_bwd_nodes does that once after the time loop)."""
        self._ck(KW)
        p = L.Attn2BwdT()
        nh, nodes, P, Hk = KW.shape
        assert nodes == q.shape[0] and dKW.is_contiguous() and dVW.is_contiguous() and dq.stride(1) == 1
        p.KW, p.VW, p.q, p.alpha, p.dco, p.dalpha_ext = KW.data_ptr(), VW.data_ptr(), q.data_ptr(), alpha.data_ptr(), _ptr(dco), _ptr(dalpha_ext)
        p.dq, p.dKW, p.dVW = dq.data_ptr(), dKW.data_ptr(), dVW.data_ptr()
        p.ldq, p.ldalpha, p.lddco, p.lddq = q.stride(0), alpha.stride(0), (dco.stride(0) if dco is not None else 0), dq.stride(0)
        p.rows, p.nh, p.P, p.Hk, p.Hv, p.scale = q.shape[0], nh, P, Hk, VW.shape[3], scale
        if dalpha_ext is not None:
            assert dalpha_ext.stride(0) == alpha.stride(0)
        if ln is not None:
            dy, co, st, dgr, dbr = ln['dy'], ln['co'], ln['stats'], ln['dgamma_rows'], ln['dbeta_rows']
            assert dy.dtype == torch.float32 and dy.stride(1) == 1 and co.stride(1) == 1 and st.is_contiguous()
            assert dgr.stride(1) == 1 and dbr.stride(1) == 1 and dgr.stride(0) == dbr.stride(0)
            p.dy, p.lddy, p.co, p.ldco = dy.data_ptr(), dy.stride(0), co.data_ptr(), co.stride(0)
            for k in range(nh):
                p.gamma[k] = ln['gamma'][k].data_ptr()
            p.stats, p.stats_head_stride = st.data_ptr(), st.stride(0)
            p.dgamma_rows, p.dbeta_rows, p.ld_dparam = dgr.data_ptr(), dbr.data_ptr(), dgr.stride(0)
            if ln.get('drop') is not None:
                p.drop_p, p.seed, p.offset = ln['drop']
                p.offset_head_stride = ln['drop_head_stride']
        else:
            assert dco is not None
        if save is not None:
            dl_s, dco_s = save
            assert dl_s.stride(1) == 1 and dco_s.stride(1) == 1 and dl_s.dtype == torch.float32 and dco_s.dtype == torch.float32
            p.dl_save, p.ld_dl_save, p.dco_save, p.ld_dco_save = dl_s.data_ptr(), dl_s.stride(0), dco_s.data_ptr(), dco_s.stride(0)
        self.launches += 1
        L.check(self.lib.dlsg_attn2_bwd(C.byref(p), _stream()), 'dlsg_attn2_bwd')

    def attn2_bwd_nodes(self, q_all, dl_all, alpha_all, dco_all, dKW, dVW, accumulate=False):
        """q_all (T, rows, Hk), dl_all / alpha_all (T, rows, nh*P), dco_all (T, rows, nh*Hv) fp32 (unit inner strides);
        dKW / dVW (nh, rows, P, H) contiguous: the node gradients of the hoisted attention summed over the T steps."""
        self._ck(q_all)
        T, rows, Hk = q_all.shape
        nh, rows2, P, Hk2 = dKW.shape
        Hv = dVW.shape[3]
        assert rows2 == rows and Hk2 == Hk and dKW.is_contiguous() and dVW.is_contiguous()
        for x in (q_all, dl_all, alpha_all, dco_all):
            assert x.dtype == torch.float32 and x.stride(2) == 1 and x.shape[0] == T and x.shape[1] == rows
        self.launches += 1
        L.check(self.lib.dlsg_attn2_bwd_nodes(q_all.data_ptr(), q_all.stride(1), q_all.stride(0), dl_all.data_ptr(), dl_all.stride(1),
                                              dl_all.stride(0), alpha_all.data_ptr(), alpha_all.stride(1), alpha_all.stride(0),
                                              dco_all.data_ptr(), dco_all.stride(1), dco_all.stride(0), dKW.data_ptr(), dVW.data_ptr(),
                                              T, rows, nh, P, Hk, Hv, int(accumulate), _stream()), 'dlsg_attn2_bwd_nodes')

    # ------------------------------------------------------------------ LatentPSL
    @staticmethod
    def latent_psl_supported(T, P, H):
        return P <= 8 and T <= 32 and H % 4 == 0

    def latent_psl_fwd(self, X, theta, Gs, N):
        B, T, H = X.shape
        P = theta.shape[0]
        assert X.is_contiguous() and theta.is_contiguous() and Gs.is_contiguous() and N.is_contiguous()
        self.launches += 1
        L.check(self.lib.dlsg_latent_psl_fwd(X.data_ptr(), theta.data_ptr(), Gs.data_ptr(), N.data_ptr(), B, T, P, H, _stream()),
                'latent_psl_fwd')

    def latent_psl_bwd(self, X, theta, Gs, dN, dX, dtheta):
        B, T, H = X.shape
        P = theta.shape[0]
        assert X.is_contiguous() and dN.is_contiguous() and dX.is_contiguous() and dtheta.is_contiguous()
        self.launches += 1
        L.check(self.lib.dlsg_latent_psl_bwd(X.data_ptr(), theta.data_ptr(), Gs.data_ptr(), dN.data_ptr(), dX.data_ptr(),
                                             dtheta.data_ptr(), B, T, P, H, _stream()), 'latent_psl_bwd')

    def latent_psl_fwd_multi(self, X, theta, Gs, N):
        """Lists (<= 2 entries: the two encoders) of the latent_psl_fwd arguments: one launch."""
        E = len(X)
        B, T, H = X[0].shape
        P = theta[0].shape[0]
        arr = lambda ts: (C.c_void_p * E)(*[t.data_ptr() for t in ts])
        for t in list(X) + list(theta) + list(Gs) + list(N):
            assert t.is_contiguous() and t.dtype == torch.float32
        self.launches += 1
        L.check(self.lib.dlsg_latent_psl_fwd_multi(arr(X), arr(theta), arr(Gs), arr(N), E, B, T, P, H, _stream()), 'latent_psl_fwd_multi')

    def latent_psl_bwd_multi(self, X, theta, Gs, dN, dX, dtheta):
        E = len(X)
        B, T, H = X[0].shape
        P = theta[0].shape[0]
        arr = lambda ts: (C.c_void_p * E)(*[t.data_ptr() for t in ts])
        for t in list(X) + list(theta) + list(Gs) + list(dN) + list(dX) + list(dtheta):
            assert t.is_contiguous() and t.dtype == torch.float32
        self.launches += 1
        L.check(self.lib.dlsg_latent_psl_bwd_multi(arr(X), arr(theta), arr(Gs), arr(dN), arr(dX), arr(dtheta), E, B, T, P, H, _stream()),
                'latent_psl_bwd_multi')

    # ------------------------------------------------------------------ one LSTM step (recurrent product + cell) in one launch
    def lstm_step_supported(self, B, H):
        return bool(self.lib.dlsg_lstm_step_supported(B, H))

    def lstm_step_fwd(self, W, h_in, gin, c_in, c_out, acts, h_out=None, h_op=None):
        """Lists with one entry per direction (<= 2).  W[d] bf16 (4H, H) contiguous; h_in[d] bf16 (B, H) view or h_in=None for
        the first step; gin[d] fp32 (B, 4H) view; c_in / c_out fp32 (B, H) contiguous; acts[d] fp32 (B, 4H) contiguous
        (activated gates out); h_out[d] fp32 (B, H) view; h_op[d] bf16 (B, H) view (the next step's h_in) or None."""
        nd = len(W)
        p = L.LstmStepT()
        B, H4 = gin[0].shape
        H = H4 // 4
        self._ck(gin[0])
        for d in range(nd):
            assert W[d].dtype == torch.bfloat16 and W[d].is_contiguous() and tuple(W[d].shape) == (H4, H)
            assert gin[d].dtype == torch.float32 and gin[d].stride(1) == 1 and gin[d].stride(0) == gin[0].stride(0)
            assert c_out[d].is_contiguous() and acts[d].is_contiguous() and acts[d].dtype == torch.float32
            p.W[d], p.gin[d], p.c_out[d], p.acts[d] = W[d].data_ptr(), gin[d].data_ptr(), c_out[d].data_ptr(), acts[d].data_ptr()
            if c_in is not None and c_in[d] is not None:
                assert c_in[d].is_contiguous()
                p.c_in[d] = c_in[d].data_ptr()
            if h_in is not None:
                assert h_in[d].dtype == torch.bfloat16 and h_in[d].stride(1) == 1 and h_in[d].stride(0) == h_in[0].stride(0)
                p.h_in[d] = h_in[d].data_ptr()
            if h_out is not None:
                assert h_out[d].dtype == torch.float32 and h_out[d].stride(1) == 1 and h_out[d].stride(0) == h_out[0].stride(0)
                p.h_out[d] = h_out[d].data_ptr()
            if h_op is not None:
                assert h_op[d].dtype == torch.bfloat16 and h_op[d].stride(1) == 1 and h_op[d].stride(0) == h_op[0].stride(0)
                p.h_op[d] = h_op[d].data_ptr()
        p.ldgin = gin[0].stride(0)
        p.ldh_in = h_in[0].stride(0) if h_in is not None else 0
        p.ldh_out = h_out[0].stride(0) if h_out is not None else 0
        p.ldh_op = h_op[0].stride(0) if h_op is not None else 0
        p.B, p.H, p.ndir = B, H, nd
        self.launches += 1
        L.check(self.lib.dlsg_lstm_step_fwd(C.byref(p), _stream()), 'dlsg_lstm_step_fwd')

    # ------------------------------------------------------------------ fused region -> frame aggregation (layer.py:184-192)
    def region_aggregate_supported(self, T, TR, H, dtype):
        return dtype == torch.bfloat16 and bool(self.lib.dlsg_region_aggregate_supported(T, TR, H))

    @staticmethod
    def _ra_set(field, tensors, check=None):
        for i, t in enumerate(tensors):
            if t is not None:
                if check is not None:
                    check(t)
                field[i] = t.data_ptr()

    def region_aggregate_fwd(self, Y, F, gamma, beta, scale, T, agg=None, U=None, stats=None, St=None, tconst=None, scores_only=False):
        """Lists (one entry per encoder, E <= 2).  Y[e]: bf16 (B*TR, H) view (unit inner stride); F[e]: fp32 (B*T, H);
        outputs (optional lists) agg / U (B*T, H) fp32, stats (B*TR, 2), St (B, T, TR) raw scores, tconst (B*T, 4) =
        [sum_h bf16(F*gamma), F.beta, m, 1/l] (softmax weights = exp(scale*St - m) / l).
        scores_only: only St = F . LN(Y) and the first two columns of tconst."""
        E = len(Y)
        p = L.RegionAggFwdT()
        rowsY, H, ldy = _rows2d(Y[0])
        rowsF, _, ldf = _rows2d(F[0])
        B = rowsF // T
        TR = rowsY // B
        self._ck(Y[0])
        assert Y[0].dtype == torch.bfloat16 and F[0].dtype == torch.float32

        def same(ref_ld):
            def f(t):
                assert t.stride(-1) == 1 and _rows2d(t)[2] == ref_ld, 'encoders must share the row pitch'
            return f

        def contig(t):
            assert t.is_contiguous() and t.dtype == torch.float32
        self._ra_set(p.Y, Y, same(ldy))
        self._ra_set(p.F, F, same(ldf))
        self._ra_set(p.gamma, gamma, contig)
        self._ra_set(p.beta, beta, contig)
        p.ldy, p.ldf = ldy, ldf
        if agg is not None:
            p.ldagg = _rows2d(agg[0])[2]
            self._ra_set(p.agg, agg, same(p.ldagg))
        if U is not None:
            p.ldu = _rows2d(U[0])[2]
            self._ra_set(p.U, U, same(p.ldu))
        for name, lst in (('stats', stats), ('St', St), ('tconst', tconst)):
            if lst is not None:
                self._ra_set(getattr(p, name), lst, contig)
        p.B, p.E, p.T, p.TR, p.H, p.scores_only, p.scale = B, E, T, TR, H, int(scores_only), float(scale)     # (2: loads only, tools)
        self.launches += 1
        L.check(self.lib.dlsg_region_aggregate_fwd(C.byref(p), _stream()), 'dlsg_region_aggregate_fwd')

    def region_aggregate_bwd_workspace(self, B, T, TR):
        return int(self.lib.dlsg_region_aggregate_bwd_workspace(B, T, TR))

    def region_aggregate_bwd(self, Y, stats, St, dSm, F, dA, U, tcF, tcA, gamma, beta, scale, T, dpre, dF, dgamma, dbeta, dbias=None,
                             work=None):
        """Second backward pass (after region_aggregate_fwd(scores_only=True, F=dA) produced dSm, tcA); work[e]: uint8 scratch
        of region_aggregate_bwd_workspace(B, T, TR) bytes.
        dpre[e]: bf16 (B*TR, H) view written; dF[e] (B*T, H) fp32 written (= dA + score-path gradient); dgamma / dbeta /
        dbias[e] (H) fp32 accumulated (zero them first)."""
        E = len(Y)
        p = L.RegionAggBwdT()
        rowsY, H, ldy = _rows2d(Y[0])
        rowsF, _, ldf = _rows2d(F[0])
        B = rowsF // T
        TR = rowsY // B
        self._ck(Y[0])
        assert Y[0].dtype == torch.bfloat16 and dpre[0].dtype == torch.bfloat16

        def same(ref_ld):
            def f(t):
                assert t.stride(-1) == 1 and _rows2d(t)[2] == ref_ld, 'encoders must share the row pitch'
            return f

        def contig(t):
            assert t.is_contiguous() and t.dtype == torch.float32
        p.ldy, p.ldf, p.ldda, p.ldu = ldy, ldf, _rows2d(dA[0])[2], _rows2d(U[0])[2]
        p.ldd, p.lddf = _rows2d(dpre[0])[2], _rows2d(dF[0])[2]
        self._ra_set(p.Y, Y, same(ldy))
        self._ra_set(p.F, F, same(ldf))
        self._ra_set(p.dA, dA, same(p.ldda))
        self._ra_set(p.U, U, same(p.ldu))
        self._ra_set(p.dpre, dpre, same(p.ldd))
        self._ra_set(p.dF, dF, same(p.lddf))
        for name, lst in (('stats', stats), ('St', St), ('dSm', dSm), ('tcF', tcF), ('tcA', tcA), ('gamma', gamma),
                          ('beta', beta), ('dgamma', dgamma), ('dbeta', dbeta), ('dbias', dbias)):
            if lst is not None:
                self._ra_set(getattr(p, name), lst, contig)
        need = self.region_aggregate_bwd_workspace(B, T, TR)
        for i, w in enumerate(work):
            assert w.dtype == torch.uint8 and w.is_contiguous() and w.numel() >= need
            p.work[i] = w.data_ptr()
        p.B, p.E, p.T, p.TR, p.H, p.scale = B, E, T, TR, H, float(scale)
        self.launches += 2
        L.check(self.lib.dlsg_region_aggregate_bwd(C.byref(p), _stream()), 'dlsg_region_aggregate_bwd')

    # ------------------------------------------------------------------ embedding / small
    def embedding_gather(self, table, ids, out=None, out2=None, drop=None):
        """ids: 1-D int64 view (any stride). out/out2: (rows, W) views."""
        self._ck(table)
        dp, seed, off = drop if (drop is not None and drop[0] > 0) else (0.0, 0, 0)
        ref = out if out is not None else out2
        self.launches += 1
        L.check(self.lib.dlsg_embedding_gather(
            table.data_ptr(), ids.data_ptr(), ids.stride(0), ids.shape[0], table.shape[1],
            _ptr(out), _dt(ref) if out is None else _dt(out), out.stride(0) if out is not None else 0,
            _ptr(out2), _dt(out2) if out2 is not None else 0, out2.stride(0) if out2 is not None else 0,
            dp, seed, off, _stream()), 'dlsg_embedding_gather')

    def embedding_scatter_add(self, dtable, ids, dout, drop=None):
        self._ck(dtable)
        dp, seed, off = drop if (drop is not None and drop[0] > 0) else (0.0, 0, 0)
        self.launches += 1
        L.check(self.lib.dlsg_embedding_scatter_add(dtable.data_ptr(), ids.data_ptr(), ids.stride(0), ids.shape[0],
                                                    dtable.shape[1], dout.data_ptr(), dout.stride(0), dp, seed, off,
                                                    _stream()), 'dlsg_embedding_scatter_add')

    def mean_nodes_fwd(self, x, y):
        B, P, H = x.shape
        assert x.is_contiguous()
        self.launches += 1
        L.check(self.lib.dlsg_mean_nodes_fwd(x.data_ptr(), B, P, H, y.data_ptr(), y.stride(0), _stream()), 'mean_nodes_fwd')

    def mean_nodes_bwd(self, dy, dx):
        B, P, H = dx.shape
        assert dx.is_contiguous()
        self.launches += 1
        L.check(self.lib.dlsg_mean_nodes_bwd(dy.data_ptr(), dy.stride(0), B, P, H, dx.data_ptr(), _stream()), 'mean_nodes_bwd')

    def axpby(self, x, a, y, b):
        assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
        self.launches += 1
        L.check(self.lib.dlsg_axpby(x.data_ptr(), a, y.data_ptr(), b, x.numel(), _stream()), 'axpby')

    def add_rowbcast(self, x, pe, y, drop=None):
        assert x.is_contiguous() and pe.is_contiguous() and y.is_contiguous()
        dp, seed, off = drop if (drop is not None and drop[0] > 0) else (0.0, 0, 0)
        self.launches += 1
        L.check(self.lib.dlsg_add_rowbcast(x.data_ptr(), pe.data_ptr(), y.data_ptr(), x.numel() // pe.numel(), pe.numel(),
                                           dp, seed, off, _stream()), 'add_rowbcast')

    def dropout(self, x, y, drop):
        assert x.is_contiguous() and y.is_contiguous()
        self.launches += 1
        L.check(self.lib.dlsg_dropout(x.data_ptr(), y.data_ptr(), x.numel(), drop[0], drop[1], drop[2], _stream()), 'dropout')

    def relu_(self, x):
        assert x.is_contiguous()
        self.launches += 1
        L.check(self.lib.dlsg_relu(x.data_ptr(), x.numel(), _stream()), 'relu')

    def relu_bwd(self, r, dr, dx):
        self.launches += 1
        L.check(self.lib.dlsg_relu_bwd(r.data_ptr(), dr.data_ptr(), dx.data_ptr(), r.numel(), _stream()), 'relu_bwd')

    def mul(self, a, b, y):
        self.launches += 1
        L.check(self.lib.dlsg_mul(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _stream()), 'mul')

    # ------------------------------------------------------------------ vocab / beam
    def row_argmax(self, logits, ids):
        rows, V = logits.shape
        self.launches += 1
        L.check(self.lib.dlsg_row_argmax(logits.data_ptr(), logits.stride(0), rows, V, ids.data_ptr(), ids.stride(0), _stream()), 'row_argmax')

    def log_softmax(self, logits, out):
        rows, V = logits.shape
        self.launches += 1
        L.check(self.lib.dlsg_log_softmax(logits.data_ptr(), logits.stride(0), rows, V, out.data_ptr(), out.stride(0), _stream()), 'log_softmax')

    def ce_masked(self, logits, targets, lens, loss_sum, dlogits, inv_count, inv_count_dev=None):
        B, Lw, V = logits.shape
        assert logits.is_contiguous() and targets.is_contiguous() and lens.dtype == torch.int32
        self.launches += 1
        row_loss = torch.empty(B * Lw, dtype=torch.float32, device=logits.device)     # deterministic loss summation
        L.check(self.lib.dlsg_ce_masked(logits.data_ptr(), targets.data_ptr(), lens.data_ptr(), B, Lw, V,
                                        loss_sum.data_ptr(), _ptr(dlogits), inv_count, _ptr(inv_count_dev),
                                        row_loss.data_ptr(), _stream()), 'ce_masked')

    def beam_topk(self, logits, last, end_index, k, top_lp, top_id, normalize=True):
        rows, V = logits.shape
        self.launches += 1
        L.check(self.lib.dlsg_beam_topk(logits.data_ptr(), logits.stride(0), rows, V, _ptr(last), end_index, k,
                                        top_lp.data_ptr(), top_id.data_ptr(), 1 if normalize else 0, _stream()), 'beam_topk')

    def beam_merge(self, top_lp, top_id, last_lp, B, beam, k, new_lp, new_cls, backptr, all_end, end_index):
        self.launches += 1
        L.check(self.lib.dlsg_beam_merge(top_lp.data_ptr(), top_id.data_ptr(), last_lp.data_ptr(), B, beam, k,
                                         new_lp.data_ptr(), new_cls.data_ptr(), backptr.data_ptr(), _ptr(all_end),
                                         end_index, _stream()), 'beam_merge')

    def beam_gather(self, src, dst, backptr, B, beam):
        """src/dst: (B*beam, W) contiguous rows (padded operand buffers allowed: pass the full row)."""
        assert src.stride(-1) == 1 and src.stride() == dst.stride() and src.dtype == dst.dtype
        row_bytes = src.stride(0) * src.element_size()
        self.launches += 1
        L.check(self.lib.dlsg_beam_gather(src.data_ptr(), dst.data_ptr(), backptr.data_ptr(), B, beam, row_bytes, _stream()), 'beam_gather')

    def beam_gather_multi(self, pairs, backptr, B, beam):
        """pairs: up to 4 (src, dst) row buffers as for beam_gather; one launch re-indexes all of them."""
        n = len(pairs)
        src, dst, rb = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int32 * n)()
        for k, (s_, d_) in enumerate(pairs):
            assert s_.stride(-1) == 1 and s_.stride() == d_.stride() and s_.dtype == d_.dtype
            src[k], dst[k], rb[k] = s_.data_ptr(), d_.data_ptr(), s_.stride(0) * s_.element_size()
        self.launches += 1
        L.check(self.lib.dlsg_beam_gather_multi(src, dst, rb, n, backptr.data_ptr(), B, beam, _stream()), 'beam_gather_multi')

    def beam_backtrack(self, preds, backs, S, B, beam, out):
        self.launches += 1
        L.check(self.lib.dlsg_beam_backtrack(preds.data_ptr(), _ptr(backs), S, B, beam, out.data_ptr(), _stream()), 'beam_backtrack')


_backend = None


def backend():
    global _backend
    if _backend is None:
        _backend = CudaBackend()
    return _backend


def set_backend(b):
    """Test hook only (tests/cpu_emul.py)."""
    global _backend
    _backend = b
