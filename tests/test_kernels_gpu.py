"""Per-kernel parity on the GPU: every libdlsg entry (through the C-ABI / ctypes) against a plain
torch fp32 restatement of the same op (tests/cpu_emul.py) on the same seeded inputs.

Tolerances: fp32 kernels 1e-5 relative to the output scale; bf16-operand GEMMs are compared against the
fp32 product of the SAME bf16-rounded operands (accumulation-order noise only, 2e-3 of the output scale);
integer outputs (ids, back-pointers, token tables) bit-exact.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from dlsg import ops
from cpu_emul import CpuEmulBackend

DEV = 'cuda'


@pytest.fixture(scope='module')
def be():
    if not torch.cuda.is_available():
        pytest.skip('no GPU')
    return ops.CudaBackend()


EM = CpuEmulBackend()


def R(*shape, dtype=torch.float32, scale=1.0, seed=[0]):
    seed[0] += 1
    g = torch.Generator().manual_seed(seed[0])
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


def both(fn_name, be, cpu_args, cpu_kwargs, outs, tol=1e-5, int_outs=()):
    """Run emulator on CPU tensors and the kernel on CUDA clones; compare the named output tensors."""
    def to_dev(x):
        if not torch.is_tensor(x):
            return x
        # clone the whole storage so that views keep their offsets / strides / padding on the device
        base = torch.empty(0, dtype=x.dtype).set_(x.untyped_storage()).to(DEV)
        return torch.as_strided(base, x.shape, x.stride(), x.storage_offset())
    g_args = [to_dev(a) for a in cpu_args]
    g_kwargs = {k: to_dev(v) for k, v in cpu_kwargs.items()}
    getattr(EM, fn_name)(*cpu_args, **cpu_kwargs)
    getattr(be, fn_name)(*g_args, **g_kwargs)
    torch.cuda.synchronize()
    allv = dict(zip(range(len(cpu_args)), zip(cpu_args, g_args)))
    allv.update({k: (cpu_kwargs[k], g_kwargs[k]) for k in cpu_kwargs})
    for o in outs:
        c, g = allv[o]
        c, g = c.float(), g.float().cpu()
        scale = max(1.0, float(c.abs().max()))
        err = float((c - g).abs().max())
        assert err <= tol * scale, (fn_name, o, err, scale)
    for o in int_outs:
        c, g = allv[o]
        assert torch.equal(c, g.cpu()), (fn_name, o)


def bf(x):
    return x.to(torch.bfloat16)


# ----------------------------------------------------------------------------------------------- GEMM (tcgen05)
TC_SHAPES = [(300, 200, 256), (128, 128, 64), (129, 257, 72), (64, 4096, 2880), (26, 1024, 936), (1664, 1055, 1536),
             (5, 24, 8), (2000, 2048, 2048)]


@pytest.mark.parametrize('M,N,K', TC_SHAPES)
def test_gemm_tc_plain(be, M, N, K):
    a, b = bf(R(M, K)), bf(R(N, K))
    out = torch.zeros(M, N)
    both('gemm', be, [a, b, out], {}, [2], tol=2e-3)


@pytest.mark.parametrize('M,N,K', [(200, 300, 128), (64, 512, 200), (936, 26, 1024), (200, 320, 128), (300, 2048, 64)])
def test_gemm_tc_epilogues(be, M, N, K):
    a, b = bf(R(M, K, scale=0.2)), bf(R(N, K, scale=0.2))
    bias_n, bias_m = R(N), R(M)
    both('gemm', be, [a, b, torch.zeros(M, N)], dict(bias=bias_n, bias_axis='n', tanh=True), [2], tol=2e-3)
    both('gemm', be, [a, b, torch.zeros(M, N)], dict(bias=bias_m, bias_axis='m', alpha=0.5), [2], tol=2e-3)
    both('gemm', be, [a, b, R(M, N)], dict(accum=True), [2], tol=2e-3)
    both('gemm', be, [a, b, R(M, N)], dict(accum=True, bias=bias_n), [2], tol=2e-3)
    both('gemm', be, [a, b, R(N, M).t()], dict(accum=True, bias=bias_m, bias_axis='m'), [2], tol=2e-3)   # accumulate + STORE_T
    both('gemm', be, [a, b, torch.zeros(M, N, dtype=torch.bfloat16)], dict(bias=bias_n), [2], tol=1e-2)
    both('gemm', be, [a, b, torch.zeros(N, M).t()], dict(bias=bias_n), [2], tol=2e-3)      # STORE_T
    both('gemm', be, [a, b, torch.zeros(M, N + 5)[:, :N]], dict(), [2], tol=2e-3)            # ldd > N
    both('gemm', be, [a, b, torch.zeros(M, N + 8, dtype=torch.bfloat16)[:, :N]], dict(bias=bias_n, tanh=True), [2], tol=1e-2)
    both('gemm', be, [a, b, torch.zeros(M, N + 8, dtype=torch.bfloat16)[:, 8:]], dict(bias=bias_m, bias_axis='m', alpha=0.5), [2], tol=1e-2)
    both('gemm', be, [a, b, R(M, N + 4)[:, 4:]], dict(accum=True, bias=bias_n), [2], tol=2e-3)


@pytest.mark.parametrize('M,N,K', [(8192, 2048, 192), (128 * 75 + 17, 1024, 136), (5000, 1000, 72), (4736, 2048, 2048)])
def test_gemm_tc_cta_pair_tiles(be, M, N, K):
    """Shapes large enough for the CTA-pair path (256 x bn tiles, tcgen05.mma.cta_group::2, thread-block cluster of two):
    an odd number of 128-row tiles (the last pair is half empty), N not a multiple of the tile, K tails, the bias + tanh bf16
    epilogue of the region projection, fp32 accumulate, a transposed store, and both operands MN-major (the weight-gradient
    form); each against the fp32 product of the same bf16 operands."""
    a, b = bf(R(M, K, scale=0.2)), bf(R(N, K, scale=0.2))
    bias_n = R(N)
    both('gemm', be, [a, b, torch.zeros(M, N)], {}, [2], tol=2e-3)
    both('gemm', be, [a, b, torch.zeros(M, N, dtype=torch.bfloat16)], dict(bias=bias_n, tanh=True), [2], tol=1e-2)
    both('gemm', be, [a, b, R(M, N)], dict(accum=True, bias=bias_n), [2], tol=2e-3)
    both('gemm', be, [a, b, torch.zeros(N, M).t()], dict(bias=bias_n), [2], tol=2e-3)                       # STORE_T
    if K % 8 == 0 and M % 8 == 0 and N % 8 == 0:
        at, bt = bf(R(K, M, scale=0.2)), bf(R(K, N, scale=0.2))                                            # MN-major views
        both('gemm', be, [at.t(), bt.t(), torch.zeros(M, N)], {}, [2], tol=2e-3)
        both('gemm', be, [a, bt.t(), torch.zeros(M, N)], {}, [2], tol=2e-3)


def test_gemm_tc_batched_and_strided_views(be):
    B_, M, N, K = 6, 150, 40, 136
    a, b = bf(R(B_, M, K)), bf(R(B_, N, K))
    both('gemm', be, [a, b, torch.zeros(B_, M, N)], {}, [2], tol=2e-3)
    both('gemm', be, [a, b, torch.zeros(B_, N, M).transpose(1, 2)], dict(alpha=0.25), [2], tol=2e-3)
    # operand views with a column offset / padded pitch (the decoder's Xl[:, oq:oq+Hq] pattern)
    big_a, big_b = bf(R(70, 512)), bf(R(96, 640))
    both('gemm', be, [big_a[:, 128:384], big_b[:, 64:320], torch.zeros(70, 96)], {}, [2], tol=2e-3)
    # heads as batch with interleaved rows (ctx (R, nh, H) -> (nh, R, H))
    x = bf(R(33, 2, 64))
    w = bf(R(2, 48, 64))
    both('gemm', be, [x.transpose(0, 1), w, torch.zeros(33, 2, 48).transpose(0, 1)], {}, [2], tol=2e-3)


@pytest.mark.parametrize('M,N,K', [(304, 200, 256), (128, 128, 64), (136, 264, 72), (64, 4096, 2880), (24, 1024, 936), (1664, 1056, 1536),
                                   (8, 24, 8), (2048, 2048, 1664), (64, 40, 520)])
@pytest.mark.parametrize('ta,tb', [(True, False), (False, True), (True, True)])
def test_gemm_tc_mn_major_operands(be, M, N, K, ta, tb):
    """Transposed (MN-major) operand views are read in place: a = A^T view of a (K,M) buffer, b likewise."""
    a = bf(R(K, M)).t() if ta else bf(R(M, K))
    b = bf(R(K, N)).t() if tb else bf(R(N, K))
    both('gemm', be, [a, b, torch.zeros(M, N)], {}, [2], tol=2e-3)
    both('gemm', be, [a, b, torch.zeros(N, M).t()], dict(bias=R(N), alpha=0.5), [2], tol=2e-3)


def test_gemm_tc_mn_major_batched_and_padded(be):
    B_, M, N, K = 3, 150, 72, 136
    a = bf(R(B_, K, M + 2))[:, :, :M].transpose(1, 2)       # MN-major with a padded pitch
    b = bf(R(B_, K, N)).transpose(1, 2)
    both('gemm', be, [a, b, torch.zeros(B_, M, N)], {}, [2], tol=2e-3)
    # wgrad shape: both operands MN-major, long K (split-K path for the small output)
    dy, x = bf(R(1664, 256)), bf(R(1664, 320))
    both('gemm', be, [dy.t(), x.t(), torch.zeros(256, 320)], {}, [2], tol=4e-3)


@pytest.mark.parametrize('splitk', [2, 3, 5])
def test_gemm_tc_splitk(be, splitk):
    M, N, K = 64, 1024, 64 * 15
    a, b = bf(R(M, K)), bf(R(N, K))
    both('gemm', be, [a, b, torch.zeros(splitk, M, N)], dict(splitk=splitk, bias=R(N)), [2], tol=2e-3)


@pytest.mark.parametrize('M,N,K,splitk', [(192, 512, 2048, 4), (192, 2048, 512, 2), (130, 512, 2048, 3)])
def test_gemm_tc_explicit_splitk_more_than_64_rows(be, M, N, K, splitk):
    """Explicit split-K partials for a recurrent GEMM with more than 64 rows (the critic's stacked LSTM, 192 rows:
    dlsg.linalg.splitk_rows), weight read K-major and as a transposed (MN-major) view."""
    a, w = bf(R(M, K, scale=0.3)), bf(R(N, K, scale=0.3))
    both('gemm', be, [a, w, torch.zeros(splitk, M, N)], dict(splitk=splitk), [2], tol=2e-3)
    wt = bf(R(K, N, scale=0.3))
    both('gemm', be, [a, wt.t(), torch.zeros(splitk, M, N)], dict(splitk=splitk), [2], tol=2e-3)


@pytest.mark.parametrize('M,N,K', [(64, 4096, 2880), (64, 2880, 4096), (17, 1000, 1544), (256, 384, 1664), (3, 130, 520)])
def test_gemm_tc_auto_splitk_fixup(be, M, N, K):
    """Skinny problems split K automatically (partials in the workspace + a reduce/epilogue kernel, fixed summation
    order): every epilogue form, repeated launches and run-to-run determinism."""
    a, b = bf(R(M, K, scale=0.2)), bf(R(N, K, scale=0.2))
    bias_n, bias_m = R(N), R(M)
    for _ in range(2):
        both('gemm', be, [a, b, torch.zeros(M, N)], dict(bias=bias_n, tanh=True), [2], tol=2e-3)
        both('gemm', be, [a, b, torch.zeros(M, N)], dict(bias=bias_m, bias_axis='m', alpha=0.5), [2], tol=2e-3)
        both('gemm', be, [a, b, R(M, N)], dict(accum=True), [2], tol=2e-3)
        both('gemm', be, [a, b, torch.zeros(M, N, dtype=torch.bfloat16)], dict(bias=bias_n), [2], tol=1e-2)
        both('gemm', be, [a, b, torch.zeros(N, M).t()], dict(bias=bias_n), [2], tol=2e-3)
        both('gemm', be, [a, b, R(M, N)], dict(atomic=True, bias=bias_n), [2], tol=2e-3)     # split-K landing in D by atomic adds
        both('gemm', be, [a, b, R(N, M).t()], dict(atomic=True), [2], tol=2e-3)
    ad, bd = a.to(DEV), b.to(DEV)
    outs = []
    for _ in range(4):
        o = torch.empty(M, N, device=DEV)
        be.gemm(ad, bd, o)
        outs.append(o)
    torch.cuda.synchronize()
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    if (M, N, K) == (3, 130, 520):
        B_ = 5                                             # batched skinny problem: per-batch tile counters
        a3, b3 = bf(R(B_, M, K)), bf(R(B_, N, K))
        both('gemm', be, [a3, b3, torch.zeros(B_, M, N)], {}, [2], tol=2e-3)


# ----------------------------------------------------------------------------------------------- GEMM (FFMA)
def test_gemm_simt_strided(be):
    a, b = R(70, 90), R(50, 90)
    both('gemm', be, [a, b, torch.zeros(70, 50)], dict(bias=R(50), tanh=True), [2])
    at, bt = R(90, 70).t(), R(90, 50).t()
    both('gemm', be, [at, bt, torch.zeros(70, 50)], dict(alpha=0.3), [2])
    both('gemm', be, [R(4, 26, 5).transpose(1, 2), R(4, 26, 64).transpose(1, 2), torch.zeros(4, 5, 64)], {}, [2])
    both('gemm', be, [R(33, 17), bf(R(9, 17)), torch.zeros(33, 9)], dict(accum=False), [2])
    both('gemm', be, [R(3, 10), R(1, 10), torch.zeros(3, 1)], {}, [2])


# ----------------------------------------------------------------------------------------------- conversions
def test_convert_and_colsum(be):
    x = R(77, 130)
    both('convert', be, [x], dict(dst=torch.zeros(77, 136, dtype=torch.bfloat16)[:, :130], dstT=torch.zeros(130, 80, dtype=torch.bfloat16)[:, :77]),
         ['dst', 'dstT'], tol=1e-2)
    both('convert', be, [R(64, 256)], dict(dst=torch.zeros(64, 256, dtype=torch.bfloat16)), ['dst'], tol=1e-2)
    xb = R(5, 26, 40)
    both('convert', be, [xb], dict(dstT=torch.zeros(5, 40, 32)[:, :, :26]), ['dstT'])
    both('convert', be, [bf(R(40, 24))], dict(dst=torch.zeros(40, 24)), ['dst'])
    both('colsum', be, [R(1000, 70), torch.zeros(70)], {}, [1], tol=1e-5)
    both('colsum', be, [bf(R(300, 64)), R(64)], {}, [1], tol=1e-3)


def test_adam_multi_segments(be):
    """dlsg_adam_multi over a device table: whole parameters, column segments with bf16 operand copies (own pitch), a
    1-row bias, an odd width (scalar path); three steps against the emulator's torch.optim.Adam arithmetic."""
    shapes = [(300, 3884), (64, 301), (1, 4096), (513, 1024)]
    cpu = [dict(p=R(*s_), m=torch.zeros(*s_), v=torch.zeros(*s_)) for s_ in shapes]
    dst_cpu = [torch.zeros(300, 1536 + 8, dtype=torch.bfloat16), torch.zeros(300, 2352, dtype=torch.bfloat16),
               torch.zeros(513, 1024, dtype=torch.bfloat16)]
    gpu = [{k: v.to(DEV) for k, v in c.items()} for c in cpu]
    dst_gpu = [d.to(DEV) for d in dst_cpu]
    step_c, step_g = torch.zeros(()), torch.zeros((), device=DEV)
    lr_g = torch.full((), 1.6e-4, device=DEV)

    def segs(ts, ds, grads):
        out = []
        w, gw = ts[0], grads[0]
        out.append(dict(p=w['p'][:, :1536], g=gw[:, :1536], m=w['m'][:, :1536], v=w['v'][:, :1536], dst=ds[0][:, :1536]))
        out.append(dict(p=w['p'][:, 1536:], g=gw[:, 1536:], m=w['m'][:, 1536:], v=w['v'][:, 1536:], dst=ds[1][:, :2348]))
        for i in (1, 2):
            out.append(dict(p=ts[i]['p'], g=grads[i], m=ts[i]['m'], v=ts[i]['v'], dst=None))
        out.append(dict(p=ts[3]['p'], g=grads[3], m=ts[3]['m'], v=ts[3]['v'], dst=ds[2]))
        return out
    for it in range(3):
        grads_c = [R(*s_, scale=0.3) for s_ in shapes]
        grads_g = [g.to(DEV) for g in grads_c]
        step_c += 1
        step_g += 1
        EM.adam_multi(EM.make_adam_plan(segs(cpu, dst_cpu, grads_c)), step_c, 1.6e-4, 0.5, 0.9, 1e-8)
        be.adam_multi(be.make_adam_plan(segs(gpu, dst_gpu, grads_g)), step_g, 123.0, 0.5, 0.9, 1e-8, lr_dev=lr_g)
        torch.cuda.synchronize()
    for c, g in zip(cpu, gpu):
        for k in ('p', 'm', 'v'):
            assert (c[k] - g[k].cpu()).abs().max() <= 2e-6 * max(1.0, float(c[k].abs().max())), k
    for c, g in zip(dst_cpu, dst_gpu):
        assert (c.float() - g.float().cpu()).abs().max() <= 1e-2
    # per-parameter step counters (torch keeps one `step` per parameter; they diverge for a parameter that sat out some
    # steps): each segment's bias correction follows its own counter, not the launch-wide one
    own = [3.0, 3.0, 11.0, 1.0, 40.0]
    grads_c = [R(*s_, scale=0.3) for s_ in shapes]
    grads_g = [g.to(DEV) for g in grads_c]
    sc, sg_ = segs(cpu, dst_cpu, grads_c), segs(gpu, dst_gpu, grads_g)
    for k_, (a_, b_) in enumerate(zip(sc, sg_)):
        a_['step'] = torch.tensor(own[k_])
        b_['step'] = torch.tensor(own[k_], device=DEV)
    EM.adam_multi(EM.make_adam_plan(sc), torch.tensor(999.0), 1.6e-4, 0.5, 0.9, 1e-8)
    be.adam_multi(be.make_adam_plan(sg_), torch.tensor(999.0, device=DEV), 123.0, 0.5, 0.9, 1e-8, lr_dev=lr_g)
    torch.cuda.synchronize()
    for c, g in zip(cpu, gpu):
        assert (c['p'] - g['p'].cpu()).abs().max() <= 2e-6 * max(1.0, float(c['p'].abs().max()))


def test_multi_convert_segments(be):
    """One launch over a device table of 2-D segments: bf16 / fp32 destinations with their own pitch, column slices of a
    wider source (the packed LSTM gate matrices), summed bias pairs, odd widths (scalar path), multi-chunk segments."""
    w1, w2, w3 = R(300, 3884), R(4096, 1024), R(37, 301)
    b1, b2 = R(4096), R(4096)
    d1 = torch.zeros(300, 1536 + 8, dtype=torch.bfloat16)
    d2 = torch.zeros(300, 304, dtype=torch.bfloat16)
    d3 = torch.zeros(2 * 4096, 1024, dtype=torch.bfloat16)
    d4 = torch.zeros(37, 304, dtype=torch.bfloat16)
    d5, d6 = torch.zeros(4096), torch.zeros(1, 301)
    pairs = [(w1[:, :1536], None, d1[:, :1536]), (w1[:, 3584:], None, d2[:, :300]), (w2, None, d3[4096:]),
             (w3, None, d4[:, :301]), (b1.view(1, -1), b2.view(1, -1), d5.view(1, -1)), (w3[5:6], None, d6)]
    dev_pairs = []
    for src, src2, dst in pairs:
        def to_dev(x):
            base = torch.empty(0, dtype=x.dtype).set_(x.untyped_storage()).to(DEV)
            return torch.as_strided(base, x.shape, x.stride(), x.storage_offset())
        dev_pairs.append((to_dev(src), to_dev(src2) if src2 is not None else None, to_dev(dst)))
    plan = be.make_convert_plan(dev_pairs)
    assert plan['n'] > len(pairs)                     # the big segments are split into several chunks
    be.multi_convert(plan)
    torch.cuda.synchronize()
    for (src, src2, dst), (_, _, g) in zip(pairs, dev_pairs):
        want = (src if src2 is None else src + src2).to(dst.dtype)
        assert torch.equal(want, g.cpu()), (src.shape, dst.dtype)


# ----------------------------------------------------------------------------------------------- norm family
@pytest.mark.parametrize('D', [64, 52, 1024, 1536, 2048])
@pytest.mark.parametrize('pre,post', [(True, False), (False, False), (False, True)])
def test_norm_fwd_bwd(be, D, pre, post):
    rows = 37
    x, res = R(rows, D), R(rows, D)
    gamma, beta = 1 + 0.1 * R(D), 0.1 * R(D)
    stats = torch.zeros(rows, 2)
    both('norm_fwd', be, [x, gamma, beta], dict(y=torch.zeros(rows, D), y2=torch.zeros(rows, D + 8, dtype=torch.bfloat16)[:, :D],
                                                 res=res, stats=stats, pre_tanh=pre, post_tanh=post), ['y', 'stats'], tol=2e-5)
    EM.norm_fwd(x, gamma, beta, res=res, stats=stats, pre_tanh=pre, post_tanh=post)
    dy = R(rows, D)
    both('norm_bwd', be, [dy, x, gamma, beta, stats], dict(dx=R(rows, D), res=res, dgamma=R(D), dbeta=R(D), pre_tanh=pre,
                                                           post_tanh=post, dx_accum=True), ['dx', 'dgamma', 'dbeta'], tol=3e-5)
    # fused-tanh input (GEMM epilogue produced tanh): backward still applies (1-x^2)
    t = torch.tanh(x)
    EM.norm_fwd(t, gamma, beta, stats=stats)
    both('norm_bwd', be, [dy, bf(t), gamma, beta, stats], dict(dx=torch.zeros(rows, D, dtype=torch.bfloat16), dgamma=torch.zeros(D),
                                                               dbeta=torch.zeros(D), in_is_tanh=True), ['dx'], tol=2e-2)


@pytest.mark.parametrize('rows,D,Dpad', [(2500, 1024, 2048), (2049, 512, 512), (4100, 136, 144)])
def test_norm_bf16_streaming(be, rows, D, Dpad):
    """The big-activation form (bf16 in / out, rows >= 2048): streaming kernels of norm_bf16.cu, column slices of a wider
    buffer (the two encoders' halves of the region projection), backward through the GEMM's fused tanh."""
    t = bf(torch.tanh(R(rows, Dpad)))[:, Dpad - D:]
    gamma, beta = 1 + 0.1 * R(D), 0.1 * R(D)
    stats = torch.zeros(rows, 2)
    both('norm_fwd', be, [t, gamma, beta], dict(y=torch.zeros(rows, D, dtype=torch.bfloat16), stats=stats), ['y', 'stats'], tol=1e-2)
    EM.norm_fwd(t, gamma, beta, stats=stats)
    dy = bf(R(rows, D))
    for tanh_in in (True, False):
        both('norm_bwd', be, [dy, t, gamma, beta, stats],
             dict(dx=torch.zeros(rows, Dpad, dtype=torch.bfloat16)[:, :D], dgamma=R(D), dbeta=R(D), in_is_tanh=tanh_in), ['dx'], tol=2e-2)
        both('norm_bwd', be, [dy, t, gamma, beta, stats],
             dict(dx=torch.zeros(rows, D, dtype=torch.bfloat16), dgamma=torch.zeros(D), dbeta=torch.zeros(D), in_is_tanh=tanh_in),
             ['dgamma', 'dbeta'], tol=2e-3)
        # fused column sums of dx (the producing Linear's bias gradient), accumulated onto existing values; the emulator
        # sums the un-rounded fp32 dx, the kernel too (before the bf16 store)
        both('norm_bwd', be, [dy, t, gamma, beta, stats],
             dict(dx=torch.zeros(rows, D, dtype=torch.bfloat16), dgamma=torch.zeros(D), dbeta=torch.zeros(D), in_is_tanh=tanh_in,
                  dxsum=R(D)), ['dgamma', 'dbeta', 'dxsum'], tol=2e-3)
    # not eligible for the streaming form (fp32 dx): the wrapper adds a colsum launch instead
    x32 = R(64, D)
    st32 = torch.zeros(64, 2)
    EM.norm_fwd(x32, gamma, beta, stats=st32)
    both('norm_bwd', be, [R(64, D), x32, gamma, beta, st32], dict(dx=torch.zeros(64, D), dgamma=torch.zeros(D), dbeta=torch.zeros(D),
                                                                  dxsum=R(D)), ['dx', 'dxsum'], tol=1e-4)


@pytest.mark.parametrize('rows,D', [(37, 512), (1664, 512), (9, 96), (130, 1024)])
def test_norm_double_backward(be, rows, D):
    """Closed-form backward-of-backward of LayerNorm (dlsg_norm_bwd2) against automatic differentiation."""
    x, dy, u = R(rows, D), R(rows, D), R(rows, D)
    gamma = 1 + 0.1 * R(D)
    stats = torch.zeros(rows, 2)
    EM.norm_fwd(x, gamma, torch.zeros(D), stats=stats)
    both('norm_bwd2', be, [x, dy, u, gamma, stats], dict(g_dy=torch.zeros(rows, D), g_x=torch.zeros(rows, D), g_gamma=R(D)),
         ['g_dy', 'g_x', 'g_gamma'], tol=3e-5)


def test_norm_strided_slices(be):
    D, rows = 96, 12
    big = R(rows, 4 * D)
    gamma, beta = 1 + 0.1 * R(D), 0.1 * R(D)
    both('norm_fwd', be, [big[:, D:2 * D], gamma, beta], dict(y=torch.zeros(rows, 3, D)[:, 1], stats=torch.zeros(rows, 2), post_tanh=True),
         ['y', 'stats'], tol=2e-5)


# ----------------------------------------------------------------------------------------------- LSTM cell
@pytest.mark.parametrize('H', [64, 1024, 1536])
def test_lstm_cell(be, H):
    B_ = 9
    gates = R(3, B_, 4 * H)
    c_prev = R(B_, H)
    kw = dict(h_out=torch.zeros(B_, H), row_bias=R(B_, 2, 4 * H)[:, 1], bias=R(4 * H), h2=torch.zeros(B_, H + 40, dtype=torch.bfloat16)[:, 8:8 + H],
              h3=torch.zeros(B_, 2 * H)[:, H:])
    both('lstm_cell_fwd', be, [gates, c_prev, torch.zeros(B_, H)], kw, [0, 2, 'h_out', 'h3'], tol=1e-5)
    acts = gates[0].clone()
    EM.lstm_cell_fwd(gates.clone(), c_prev, torch.zeros(B_, H))
    g2 = gates.clone()
    c_new = torch.zeros(B_, H)
    EM.lstm_cell_fwd(g2, c_prev, c_new)
    acts = g2[0].contiguous()
    kw = dict(dgates=torch.zeros(B_, 4 * H), dgates2=torch.zeros(B_, 4 * H + 16, dtype=torch.bfloat16)[:, :4 * H],
              dgatesT=torch.zeros(4 * H, 5 * B_)[:, 2 * B_:3 * B_], dh2=R(B_, 3 * H)[:, H:2 * H])
    both('lstm_cell_bwd', be, [acts, c_prev, c_new, R(B_, 2 * H)[:, :H], R(B_, H), torch.zeros(B_, H)], kw, [5, 'dgates', 'dgatesT'], tol=1e-5)


@pytest.mark.parametrize('H', [64, 512])
@pytest.mark.parametrize('nulls', [False, True])
def test_lstm_cell_double_backward(be, H, nulls):
    """Closed-form backward-of-backward kernel against automatic differentiation of the restated cell backward."""
    B_ = 6
    gates = R(1, B_, 4 * H)
    c_prev = R(B_, H)
    c_new = torch.zeros(B_, H)
    EM.lstm_cell_fwd(gates, c_prev, c_new)
    acts = gates[0].contiguous()
    outs = [torch.zeros(B_, H), torch.zeros(B_, H), torch.zeros(B_, 4 * H), torch.zeros(B_, H)]
    args = [acts, c_prev, c_new, R(B_, H), None if nulls else R(B_, H), R(B_, 4 * H), None if nulls else R(B_, H)] + outs
    both('lstm_cell_bwd2', be, args, {}, [7, 8, 9, 10], tol=2e-5)


@pytest.mark.parametrize('H,S', [(64, 1), (512, 1), (512, 3), (50, 2), (1024, 9), (64, 16)])
def test_lstm_cell_second_order_loop_fields(be, H, S):
    """The fields the critic's fused second-order loops use (dlsg.generic._LstmBptt2): cell backward with the injections
    dc_next2 / dgates_add and the dh_total output; cell backward-of-backward with split-K partials added to u and a
    second (bf16, pitched) copy of g_dh."""
    B_ = 7
    gates = R(1, B_, 4 * H)
    c_prev, c_new = R(B_, H), torch.zeros(B_, H)
    EM.lstm_cell_fwd(gates, c_prev, c_new)
    acts = gates[0].contiguous()
    kw = dict(dgates=torch.zeros(B_, 4 * H), dgates2=torch.zeros(B_, 4 * H + 16, dtype=torch.bfloat16)[:, :4 * H], dh2=R(S, B_, H),
              dc_next2=R(B_, H), dgates_add=R(B_, 4 * H), dh_total=torch.zeros(B_, H))
    both('lstm_cell_bwd', be, [acts, c_prev, c_new, R(B_, H), R(B_, H), torch.zeros(B_, H)], kw, [5, 'dgates', 'dh_total'], tol=1e-5)
    both('lstm_cell_bwd', be, [acts, c_prev, c_new, R(B_, H), R(B_, H), torch.zeros(B_, H)], kw, ['dgates2'], tol=1e-2)
    kw = dict(dgates=torch.zeros(B_, 3, 4 * H)[:, 1], dh2=R(S, B_, H), dc_next2=None, dgates_add=R(B_, 4 * H))     # pitched dgates
    both('lstm_cell_bwd', be, [acts, c_prev, c_new, torch.zeros(B_, H), None, torch.zeros(B_, H)], kw, [5, 'dgates'], tol=1e-5)
    outs = [torch.zeros(B_, H), torch.zeros(B_, H), torch.zeros(B_, 4 * H), torch.zeros(B_, H)]
    args = [acts, c_prev, c_new, R(B_, H), R(B_, H), R(B_, 4 * H), R(B_, H)] + outs
    kw = dict(u2=R(S, B_, 4 * H) if S > 1 else R(B_, 4 * H), g_dh2=torch.zeros(B_, H + 24, dtype=torch.bfloat16)[:, 8:8 + H])
    both('lstm_cell_bwd2', be, args, kw, [7, 8, 9, 10], tol=2e-5)
    both('lstm_cell_bwd2', be, args, kw, ['g_dh2'], tol=1e-2)
    args[5] = None                                            # u only from the partials
    both('lstm_cell_bwd2', be, args, dict(u2=R(S, B_, 4 * H)), [7, 8, 9, 10], tol=2e-5)
    args[5], args[7] = R(B_, 2, 4 * H)[:, 1], torch.zeros(B_, 3, H)[:, 2]        # row-pitched u and g_dh (batch-major slices)
    both('lstm_cell_bwd2', be, args, {}, [7, 8, 9, 10], tol=2e-5)


@pytest.mark.parametrize('H,post', [(64, False), (1024, False), (1536, True)])
def test_fused_cell_norm(be, H, post):
    B_ = 7
    gates = R(3, B_, 4 * H)
    c_prev = R(B_, H)
    gamma, beta = 1 + 0.1 * R(H), 0.1 * R(H)
    kw = dict(h_out=torch.zeros(B_, H), row_bias=R(B_, 4 * H), bias=R(4 * H), h2=torch.zeros(B_, H + 40, dtype=torch.bfloat16)[:, 8:8 + H],
              y2=torch.zeros(B_, 2 * H, dtype=torch.bfloat16)[:, H:], stats=torch.zeros(B_, 2), post_tanh=post)
    both('lstm_cell_norm_fwd', be, [gates, c_prev, torch.zeros(B_, H), gamma, beta, torch.zeros(B_, 3 * H)[:, H:2 * H]], kw,
         [0, 2, 5, 'h_out', 'stats'], tol=2e-5)
    g2, c_new, hh, st = gates.clone(), torch.zeros(B_, H), torch.zeros(B_, H), torch.zeros(B_, 2)
    EM.lstm_cell_norm_fwd(g2, c_prev, c_new, gamma, beta, torch.zeros(B_, H), h_out=hh, row_bias=kw['row_bias'], bias=kw['bias'], stats=st,
                          post_tanh=post)
    acts = g2[0].contiguous()
    kw = dict(dh=R(B_, 2 * H)[:, :H], dh2=R(B_, 3 * H)[:, H:2 * H], dgates=torch.zeros(B_, 4 * H),
              dgates2=torch.zeros(B_, 4 * H + 16, dtype=torch.bfloat16)[:, :4 * H], dgatesT=torch.zeros(4 * H, 5 * B_)[:, 2 * B_:3 * B_],
              dgates_sum=R(B_, 4 * H), post_tanh=post)
    both('norm_lstm_cell_bwd', be, [acts, c_prev, c_new, R(B_, H), torch.zeros(B_, H), R(B_, 2 * H)[:, H:], hh, gamma, beta, st,
                                    torch.zeros(B_, H), torch.zeros(B_, H)],
         kw, [4, 10, 11, 'dgates', 'dgatesT', 'dgates_sum'], tol=3e-5)


# ----------------------------------------------------------------------------------------------- softmax
@pytest.mark.parametrize('shape,dim', [((4, 936, 26), 1), ((4, 26, 936), 2), ((3, 26, 5), 1), ((6, 26, 26), 2), ((7, 5, 1), 1)])
@pytest.mark.parametrize('mask_mode', [0, 1, 2])
def test_softmax(be, shape, dim, mask_mode):
    x = R(*shape, scale=3.0)
    mask = (R(*shape) > -0.5).float()
    mask[0] = 0          # a fully masked slice -> uniform rows (sublayer.py:70-72)
    kw = dict(scale=0.37, mask=mask if mask_mode else None, mask_mode=mask_mode)
    both('softmax_fwd', be, [x, torch.zeros(*shape), dim], kw, [1], tol=1e-6)
    both('softmax_bwd', be, [x, R(*shape), torch.zeros(*shape), dim], kw, [2], tol=1e-5)


@pytest.mark.parametrize('shape,dim', [((4, 70, 26), 1), ((3, 26, 5), 1), ((6, 26, 26), 2), ((7, 5, 1), 1), ((9, 2), 1)])
@pytest.mark.parametrize('mask_mode', [0, 1, 2])
def test_softmax_double_backward(be, shape, dim, mask_mode):
    """Closed-form backward of the softmax backward (dlsg_softmax_bwd2) against automatic differentiation of its restatement."""
    if len(shape) == 2:
        shape = (shape[0], shape[1], 1)
    x = R(*shape, scale=3.0)
    mask = (R(*shape) > -0.5).float()
    mask[0] = 0
    kw = dict(g_dy=torch.zeros(*shape), g_x=torch.zeros(*shape), scale=0.37, mask=mask if mask_mode else None, mask_mode=mask_mode)
    both('softmax_bwd2', be, [x, R(*shape), R(*shape), dim], kw, ['g_dy', 'g_x'], tol=2e-5)
    kw['g_dy'] = None
    both('softmax_bwd2', be, [x, R(*shape), R(*shape), dim], kw, ['g_x'], tol=2e-5)


@pytest.mark.parametrize('n,cols', [(6 * 26 * 512, 26 * 512), (3 * 35, 35)])
def test_elementwise_forms(be, n, cols):
    """dlsg_ew: every op, the float4 path (aligned, n % 4 == 0) and the scalar path."""
    from dlsg import _lib as L
    t = lambda: R(n)
    z = lambda: torch.zeros(n)
    e = torch.rand(n // cols)
    cases = [(L.EW_TANH_BWD, [t(), torch.tanh(t())], 1, 0), (L.EW_TANH_BWD2, [t(), torch.tanh(t()), t()], 2, 0),
             (L.EW_MUL_BWD, [t(), t(), t()], 2, 0), (L.EW_MUL_BWD2, [t(), t(), t(), t(), t()], 3, 0),
             (L.EW_LERP_ROWS, [t(), t(), e], 1, cols), (L.EW_LERP_ROWS_BWD, [t(), e], 2, cols)]
    for op, ins, n_out, c in cases:
        ref = [z() for _ in range(n_out)]
        EM.ew(op, ins, ref, c)
        outs = [torch.zeros(n, device=DEV) for _ in range(n_out)]
        be.ew(op, [i_.to(DEV) for i_ in ins], outs, c)
        for a_, b_ in zip(ref, outs):
            assert float((a_ - b_.cpu()).abs().max()) <= 1e-6 * max(1.0, float(a_.abs().max())), op
        if n_out > 1:                                         # an output left out
            outs2 = [None] + [torch.zeros(n, device=DEV) for _ in range(n_out - 1)]
            be.ew(op, [i_.to(DEV) for i_ in ins], outs2, c)
            assert torch.equal(outs2[1], outs[1]), op


# ----------------------------------------------------------------------------------------------- node attention
@pytest.mark.parametrize('nh,P,H,rpn', [(2, 5, 1024, 1), (2, 8, 64, 1), (1, 26, 64, 1), (2, 5, 64, 5)])
def test_node_attn(be, nh, P, H, rpn):
    nodes, rows = 6, 6 * rpn
    Kp, Vp, qp = R(nh, nodes, P, H), R(nh, nodes, P, H), R(rows, nh * H)
    alpha = torch.zeros(rows, nh * P)
    both('node_attn_fwd', be, [Kp, Vp, qp, alpha, torch.zeros(rows, nh * H + 8)[:, :nh * H], rpn], {}, [3, 4], tol=2e-5)
    if rpn == 1:
        EM.node_attn_fwd(Kp, Vp, qp, alpha, torch.zeros(rows, nh * H), 1)
        both('node_attn_bwd', be, [Kp, Vp, qp, alpha, R(rows, nh * H), torch.zeros(rows, nh * H), R(nh, nodes, P, H), R(nh, nodes, P, H)],
             dict(dalpha_ext=R(rows, nh * P)), [5, 6, 7], tol=3e-5)


@pytest.mark.parametrize('nh,P,Hk,Hv,rpn', [(2, 5, 1024, 1024, 1), (2, 8, 64, 64, 1), (1, 6, 64, 96, 1), (2, 5, 1024, 1024, 5)])
def test_attn2_hoisted(be, nh, P, Hk, Hv, rpn):
    nodes, rows = 6, 6 * rpn
    KW, VW, q = R(nh, nodes, P, Hk, scale=0.3), R(nh, nodes, P, Hv), R(rows, Hk + 8)[:, :Hk]
    alpha = torch.zeros(rows, nh * P)
    both('attn2_fwd', be, [KW, VW, q, alpha, torch.zeros(rows, nh * Hv + 4)[:, :nh * Hv], 0.11, rpn], {}, [3, 4], tol=2e-5)
    if rpn == 1:
        EM.attn2_fwd(KW, VW, q, alpha, torch.zeros(rows, nh * Hv), 0.11, 1)
        both('attn2_bwd', be, [KW, VW, q, alpha, R(rows, nh * Hv), R(rows, Hk + 12)[:, 4:4 + Hk], R(nh, nodes, P, Hk), R(nh, nodes, P, Hv), 0.11],
             dict(dalpha_ext=R(rows, nh * P)), [5, 6, 7], tol=3e-5)


@pytest.mark.parametrize('nh,P,Hk,Hv,T', [(2, 5, 1024, 1024, 4), (1, 8, 64, 96, 3), (2, 3, 128, 64, 26)])
def test_attn2_deferred_node_gradients(be, nh, P, Hk, Hv, T):
    """Decoder attention backward with the node gradients deferred: every step records d(logits) / d(co)
    (`save=`) instead of read-modify-writing dKW / dVW, and ONE dlsg_attn2_bwd_nodes launch sums them over time.
    Equal to the per-step accumulation (kernel vs kernel, and vs the emulator), step slices taken from padded buffers."""
    rows = 7
    KW, VW = R(nh, rows, P, Hk, scale=0.3), R(nh, rows, P, Hv)
    q_all = R(T, rows, Hk + 4)[:, :, :Hk]
    dco_in = R(T, rows, nh * Hv)
    al_all = torch.zeros(T, rows, nh * P + 3)[:, :, :nh * P]
    for t in range(T):
        EM.attn2_fwd(KW, VW, q_all[t], al_all[t], torch.zeros(rows, nh * Hv), 0.2, 1)
    dev = lambda x: x.to(DEV)
    KWd, VWd, qd, ald, dcd = dev(KW), dev(VW), dev(q_all.contiguous()), dev(al_all.contiguous()), dev(dco_in)
    # (a) per-step accumulation on the device
    dq_a = torch.zeros(T, rows, Hk, device=DEV)
    dK_a, dV_a = torch.zeros(nh, rows, P, Hk, device=DEV), torch.zeros(nh, rows, P, Hv, device=DEV)
    for t in range(T):
        be.attn2_bwd(KWd, VWd, qd[t], ald[t], dcd[t], dq_a[t], dK_a, dV_a, 0.2)
    # (b) deferred
    dq_b = torch.zeros(T, rows, Hk, device=DEV)
    dl_s, dco_s = torch.zeros(T, rows, nh * P, device=DEV), torch.zeros(T, rows, nh * Hv, device=DEV)
    dK_b, dV_b = torch.full((nh, rows, P, Hk), 7.0, device=DEV), torch.full((nh, rows, P, Hv), 7.0, device=DEV)
    untouched = dK_b.clone()
    for t in range(T):
        be.attn2_bwd(KWd, VWd, qd[t], ald[t], dcd[t], dq_b[t], dK_b, dV_b, 0.2, save=(dl_s[t], dco_s[t]))
    assert torch.equal(dK_b, untouched)                       # the step kernel no longer touches the node gradients
    be.attn2_bwd_nodes(qd, dl_s, ald, dco_s, dK_b, dV_b)
    torch.cuda.synchronize()
    assert torch.equal(dq_a, dq_b)
    for a, b_ in ((dK_a, dK_b), (dV_a, dV_b)):
        assert float((a - b_).abs().max()) <= 2e-5 * max(1.0, float(a.abs().max()))
    # accumulate form and the emulator
    be.attn2_bwd_nodes(qd, dl_s, ald, dco_s, dK_b, dV_b, accumulate=True)
    torch.cuda.synchronize()
    assert float((dK_b - 2 * dK_a).abs().max()) <= 4e-5 * max(1.0, float(dK_a.abs().max()))
    eK, eV = torch.zeros(nh, rows, P, Hk), torch.zeros(nh, rows, P, Hv)
    EM.attn2_bwd_nodes(q_all, dl_s.cpu(), al_all, dco_s.cpu(), eK, eV)
    assert float((eK - dK_a.cpu()).abs().max()) <= 3e-5 * max(1.0, float(eK.abs().max()))
    assert float((eV - dV_a.cpu()).abs().max()) <= 3e-5 * max(1.0, float(eV.abs().max()))


@pytest.mark.parametrize('nh,P,Hk,Hv,ydt', [(2, 5, 1024, 1024, torch.bfloat16), (2, 8, 64, 64, torch.float32), (1, 6, 64, 96, torch.float32)])
def test_attn2_fused_output_layer(be, nh, P, Hk, Hv, ydt):
    """attn2 with the context output layer (tanh -> LayerNorm -> dropout) fused in: forward and backward against the
    emulator without dropout, and against the unfused GPU kernels (attn2 + norm per head, same Philox sites) with it."""
    rows = 7
    KW, VW, q = R(nh, rows, P, Hk, scale=0.3), R(nh, rows, P, Hv), R(rows, Hk)
    gam, bet = [1 + 0.1 * R(Hv) for _ in range(nh)], [0.1 * R(Hv) for _ in range(nh)]
    dy, dq0, dK0, dV0, da = R(rows, nh * Hv + 8)[:, :nh * Hv], R(rows, Hk), R(nh, rows, P, Hk), R(nh, rows, P, Hv), R(rows, nh * P)
    D = lambda x: x.to(DEV)
    for drop in (None, (0.3, 77, 4 << 32)):
        # ---- reference: emulator (no dropout) or the unfused GPU kernels (dropout)
        rb, dev_ref = (EM, (lambda x: x.clone())) if drop is None else (be, D)
        r = dict(alpha=dev_ref(torch.zeros(rows, nh * P)), co=dev_ref(torch.zeros(rows, nh * Hv)), y=dev_ref(torch.zeros(rows, nh * Hv)),
                 st=dev_ref(torch.zeros(nh, rows, 2)), dco=dev_ref(torch.zeros(rows, nh * Hv)), dq=dev_ref(dq0), dK=dev_ref(dK0), dV=dev_ref(dV0),
                 dg=[dev_ref(torch.zeros(Hv)) for _ in range(nh)], db=[dev_ref(torch.zeros(Hv)) for _ in range(nh)])
        a_ = [dev_ref(x) for x in (KW, VW, q, dy, da)]
        g_, b_ = [dev_ref(x) for x in gam], [dev_ref(x) for x in bet]
        rb.attn2_fwd(a_[0], a_[1], a_[2], r['alpha'], r['co'], 0.11, 1)
        for k in range(nh):
            sl = slice(k * Hv, (k + 1) * Hv)
            dk = None if drop is None else (drop[0], drop[1], drop[2] + (k << 28))
            rb.norm_fwd(r['co'][:, sl], g_[k], b_[k], y=r['y'][:, sl], stats=r['st'][k], pre_tanh=True, drop=dk)
            rb.norm_bwd(a_[3][:, sl], r['co'][:, sl], g_[k], b_[k], r['st'][k], dx=r['dco'][:, sl], dgamma=r['dg'][k], dbeta=r['db'][k],
                        pre_tanh=True, drop=dk)
        rb.attn2_bwd(a_[0], a_[1], a_[2], r['alpha'], r['dco'], r['dq'], r['dK'], r['dV'], 0.11, dalpha_ext=a_[4])
        # ---- fused kernels
        f = dict(alpha=D(torch.zeros(rows, nh * P)), co=D(torch.zeros(rows, nh * Hv)), y=D(torch.zeros(rows, nh * Hv + 8, dtype=ydt))[:, :nh * Hv],
                 st=D(torch.zeros(nh, rows, 2)), dq=D(dq0), dK=D(dK0), dV=D(dV0),
                 dgr=D(torch.zeros(rows, nh * Hv)), dbr=D(torch.zeros(rows, nh * Hv)))
        d_ = [D(x) for x in (KW, VW, q, dy, da)]
        dg_, db_ = [D(x) for x in gam], [D(x) for x in bet]
        be.attn2_fwd(d_[0], d_[1], d_[2], f['alpha'], f['co'], 0.11, 1,
                     ln=dict(gamma=dg_, beta=db_, y=f['y'], stats=f['st'], drop=drop, drop_head_stride=1 << 28))
        be.attn2_bwd(d_[0], d_[1], d_[2], f['alpha'], None, f['dq'], f['dK'], f['dV'], 0.11, dalpha_ext=d_[4],
                     ln=dict(dy=d_[3], co=f['co'], gamma=dg_, stats=f['st'], dgamma_rows=f['dgr'], dbeta_rows=f['dbr'], drop=drop,
                             drop_head_stride=1 << 28))
        torch.cuda.synchronize()
        cmp = lambda name, a, b, tol: _close(name, a, b, tol)
        ytol = 2e-2 if ydt == torch.bfloat16 else 3e-5
        cmp('alpha', r['alpha'], f['alpha'], 2e-5); cmp('co', r['co'], f['co'], 2e-5); cmp('y', r['y'], f['y'], ytol)
        cmp('stats', r['st'], f['st'], 1e-4); cmp('dq', r['dq'], f['dq'], 1e-4); cmp('dKW', r['dK'], f['dK'], 1e-4)
        cmp('dVW', r['dV'], f['dV'], 1e-4)
        for k in range(nh):
            sl = slice(k * Hv, (k + 1) * Hv)
            cmp('dgamma', r['dg'][k], f['dgr'][:, sl].sum(0), 1e-4); cmp('dbeta', r['db'][k], f['dbr'][:, sl].sum(0), 1e-4)


@pytest.mark.parametrize('nh,P,H,Hv,S,rpn', [(2, 5, 1024, 1024, 3, 1), (2, 5, 1024, 1024, 1, 5), (1, 8, 256, 128, 2, 1), (2, 3, 64, 1024, 4, 1)])
def test_cell_norm_attn2_one_launch_equals_two_launches(be, nh, P, H, Hv, S, rpn):
    """dlsg_cell_norm_attn2_fwd (query-LSTM cell + LayerNorm + hoisted attention + context output layer of a decode step in
    one launch) against dlsg_lstm_cell_norm_fwd followed by dlsg_attn2_fwd on the same inputs: same arithmetic and same
    Philox sites, with and without the three dropouts; rows sharing node tensors (beam search: rows_per_node)."""
    rows = 10
    nodes = rows // rpn
    D = lambda x: x.to(DEV)
    gates0, c_prev, rb = R(S, rows, 4 * H), R(rows, H), R(rows, 2, 4 * H)[:, 1]
    gam, bet = 1 + 0.1 * R(H), 0.1 * R(H)
    KW, VW = R(nh, nodes, P, H, scale=0.3), R(nh, nodes, P, Hv)
    og, ob = [1 + 0.1 * R(Hv) for _ in range(nh)], [0.1 * R(Hv) for _ in range(nh)]
    for drops in ((None, None, None), ((0.5, 11, 1 << 20), (0.3, 12, 2 << 20), (0.3, 13, 4 << 32))):
        outs = []
        for one in (False, True):
            g = D(gates0.clone())
            o = dict(c_out=D(torch.zeros(rows, H)), q32=D(torch.zeros(rows, H)), h_out=D(torch.zeros(rows, H)),
                     h2=D(torch.zeros(rows, H + 40, dtype=torch.bfloat16))[:, 8:8 + H], y2=D(torch.zeros(rows, 2 * H, dtype=torch.bfloat16))[:, H:],
                     stq=D(torch.zeros(rows, 2)), alpha=D(torch.zeros(rows, nh * P)), co=D(torch.zeros(rows, nh * Hv)),
                     y=D(torch.zeros(rows, nh * Hv + 8, dtype=torch.bfloat16))[:, :nh * Hv], stc=D(torch.zeros(nh, rows, 2)))
            cell = dict(gates=g, c_prev=D(c_prev), c_out=o['c_out'], gamma=D(gam), beta=D(bet), y=o['q32'], h_out=o['h_out'], row_bias=D(rb),
                        h2=o['h2'], y2=o['y2'], stats=o['stq'], drop=drops[0], ydrop=drops[1])
            attn = dict(KW=D(KW), VW=D(VW), q=o['q32'], alpha=o['alpha'], co=o['co'], scale=0.09, rows_per_node=rpn,
                        ln=dict(gamma=[D(x) for x in og], beta=[D(x) for x in ob], y=o['y'], stats=o['stc'], drop=drops[2],
                                drop_head_stride=1 << 28))
            if one:
                assert be.cell_norm_attn2_fwd(cell, attn)
            else:
                c2 = dict(cell)
                be.lstm_cell_norm_fwd(c2.pop('gates'), c2.pop('c_prev'), c2.pop('c_out'), c2.pop('gamma'), c2.pop('beta'), c2.pop('y'), **c2)
                be.attn2_fwd(**attn)
            torch.cuda.synchronize()
            o['acts'] = g[0]
            outs.append(o)
        for k in outs[0]:
            a_, b_ = outs[0][k].float().cpu(), outs[1][k].float().cpu()
            tol = 1e-2 if outs[0][k].dtype == torch.bfloat16 else 2e-6
            assert float((a_ - b_).abs().max()) <= tol * max(1.0, float(a_.abs().max())), (k, drops[0] is not None)
        if drops[0] is None:
            plain = outs[1]
    if S == 1:                                                   # the emulator path of the same entry
        g = gates0[0].clone()
        o = dict(c_out=torch.zeros(rows, H), q32=torch.zeros(rows, H), alpha=torch.zeros(rows, nh * P), co=torch.zeros(rows, nh * Hv),
                 y=torch.zeros(rows, nh * Hv), stc=torch.zeros(nh, rows, 2))
        assert EM.cell_norm_attn2_fwd(dict(gates=g, c_prev=c_prev, c_out=o['c_out'], gamma=gam, beta=bet, y=o['q32'], row_bias=rb),
                                      dict(KW=KW, VW=VW, q=o['q32'], alpha=o['alpha'], co=o['co'], scale=0.09, rows_per_node=rpn,
                                           ln=dict(gamma=og, beta=ob, y=o['y'], stats=o['stc'], drop=None, drop_head_stride=1 << 28)))
        _close('alpha vs emulator', o['alpha'], plain['alpha'], 2e-5)
        _close('y vs emulator', o['y'], plain['y'].float(), 2e-2)


def _close(name, a, b, tol):
    a, b = a.float().cpu(), b.float().cpu()
    err, scale = float((a - b).abs().max()), max(1.0, float(a.abs().max()))
    assert err <= tol * scale, (name, err, scale)


@pytest.mark.parametrize('B_,T,P,H', [(6, 26, 5, 1024), (3, 26, 8, 64), (4, 26, 1, 512), (2, 6, 5, 64)])
def test_latent_psl(be, B_, T, P, H):
    X, theta = R(B_, T, H), R(P, H, scale=0.05)
    Gs, N = torch.zeros(B_, T, P), torch.zeros(B_, P, H)
    both('latent_psl_fwd', be, [X, theta, Gs, N], {}, [2, 3], tol=2e-5)
    EM.latent_psl_fwd(X, theta, Gs, N)
    both('latent_psl_bwd', be, [X, theta, Gs, R(B_, P, H), torch.zeros(B_, T, H), R(P, H)], {}, [4, 5], tol=3e-5)


# ----------------------------------------------------------------------------------------------- embedding & co
def test_embedding_mean_elementwise(be):
    V_, W = 50, 20
    table = R(V_, W)
    ids = torch.randint(0, V_, (4, 9))
    both('embedding_gather', be, [table, ids[:, 3]], dict(out=torch.zeros(4, 64)[:, 8:28], out2=torch.zeros(4, 24, dtype=torch.bfloat16)[:, :W]),
         ['out'], tol=1e-6)
    both('embedding_scatter_add', be, [torch.zeros(V_, W), ids.reshape(-1), R(36, 40)[:, 10:30]], {}, [0], tol=1e-5)
    x = R(5, 8, 64)
    both('mean_nodes_fwd', be, [x, torch.zeros(5, 128)[:, 64:]], {}, [1], tol=1e-6)
    both('mean_nodes_bwd', be, [R(5, 128)[:, :64], R(5, 8, 64)], {}, [1], tol=1e-6)
    both('axpby', be, [R(1000), 0.5, R(1000), 2.0], {}, [2], tol=1e-6)
    both('add_rowbcast', be, [R(6, 26, 32), R(26, 32), torch.zeros(6, 26, 32)], {}, [2], tol=1e-6)
    both('relu_', be, [R(999)], {}, [0], tol=0)
    both('mul', be, [R(100), R(100), torch.zeros(100)], {}, [2], tol=1e-6)


def test_dropout_statistics_and_mask_reuse(be):
    n, p = 1 << 20, 0.3
    x = torch.ones(n, device=DEV)
    y = torch.empty_like(x)
    be.dropout(x, y, (p, 1234, 0))
    keep = float((y > 0).float().mean())
    assert abs(keep - (1 - p)) < 5e-3
    assert torch.allclose(y[y > 0], torch.full_like(y[y > 0], 1 / (1 - p)))
    y2 = torch.empty_like(x)
    be.dropout(x, y2, (p, 1234, 0))
    assert torch.equal(y, y2)                                   # same (seed, offset) -> same mask (backward reuses it)
    be.dropout(x, y2, (p, 1235, 0))
    assert not torch.equal(y, y2)
    # norm_fwd / norm_bwd regenerate identical masks: d(sum y)/dx is zero exactly where y was dropped
    D = 256
    xx = torch.randn(64, D, device=DEV)
    g, b = torch.ones(D, device=DEV), torch.zeros(D, device=DEV)
    yy, st = torch.empty_like(xx), torch.empty(64, 2, device=DEV)
    be.norm_fwd(xx, g, b, y=yy, stats=st, drop=(p, 77, 5 << 32))
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    be.norm_bwd(torch.ones_like(xx), xx, g, b, st, dx=torch.empty_like(xx), dgamma=dg, dbeta=db, drop=(p, 77, 5 << 32))
    assert torch.allclose(db, (yy != 0).float().sum(0) / (1 - p), atol=1e-3)


# ----------------------------------------------------------------------------------------------- vocab / beam
@pytest.mark.parametrize('V_', [37, 10547])
def test_vocab_rows(be, V_):
    rows = 20
    lg = R(rows, V_, scale=2.0)
    lg[3, 5] = lg[3, 9] = lg[3].max() + 1            # tie -> lowest index
    both('row_argmax', be, [lg, torch.zeros(rows, 3, dtype=torch.int64)[:, 1]], {}, [], int_outs=[1])
    both('log_softmax', be, [lg, torch.zeros(rows, V_)], {}, [1], tol=2e-6)
    B_, L = 4, 5
    lg3 = R(B_, L, V_, scale=2.0)
    tg = torch.randint(0, V_, (B_, L))
    lens = torch.tensor([5, 3, 1, 4], dtype=torch.int32)
    both('ce_masked', be, [lg3, tg, lens, torch.zeros(1), torch.zeros(B_, L, V_), 1.0 / 13], {}, [3, 4], tol=2e-6)


@pytest.mark.parametrize('V_,k', [(37, 5), (10547, 5), (10547, 3), (41, 1), (500, 8)])
def test_beam_kernels(be, V_, k):
    B_, beam = 6, k
    rows = B_ * beam
    lg = R(rows, V_, scale=2.0)
    last = torch.randint(0, 6, (rows,))
    end = 2
    tl, ti = torch.zeros(rows, k), torch.zeros(rows, k, dtype=torch.int64)
    getattr(EM, 'beam_topk')(lg, last, end, k, tl, ti)
    gtl, gti = torch.zeros(rows, k, device=DEV), torch.zeros(rows, k, dtype=torch.int64, device=DEV)
    be.beam_topk(lg.to(DEV), last.to(DEV), end, k, gtl, gti)
    live = last != end
    assert torch.equal(ti[live], gti.cpu()[live])
    assert (tl[live] - gtl.cpu()[live]).abs().max() < 1e-5
    assert torch.equal(gti.cpu()[~live][:, 0], torch.full((int((~live).sum()),), end))
    assert torch.all(gtl.cpu()[~live][:, 0] == 0) and (k == 1 or torch.all(torch.isinf(gtl.cpu()[~live][:, 1:])))
    # merge on the device's own top-k (the -inf tail of forced rows may pick different dummy ids than torch.topk)
    tl, ti = gtl.cpu(), gti.cpu()
    last_lp = R(B_, beam)
    outs = [torch.zeros(B_, beam), torch.zeros(B_, beam, dtype=torch.int64), torch.zeros(B_, beam, dtype=torch.int64)]
    both('beam_merge', be, [tl, ti, last_lp, B_, beam, k, outs[0], outs[1], outs[2], None, end], {}, [6], int_outs=[7, 8])
    src = R(rows, 72)
    bp = torch.randint(0, beam, (B_, beam))
    both('beam_gather', be, [src, torch.zeros(rows, 72), bp, B_, beam], {}, [1], tol=0)
    both('beam_gather', be, [bf(R(rows, 80))[:, :72], torch.zeros(rows, 80, dtype=torch.bfloat16)[:, :72], bp, B_, beam], {}, [1], tol=0)
    S = 7
    preds = torch.randint(0, V_, (S, B_, beam))
    backs = torch.randint(0, beam, (S - 1, B_, beam))
    both('beam_backtrack', be, [preds, backs, S, B_, beam, torch.zeros(B_, beam, S, dtype=torch.int64)], {}, [], int_outs=[5])
    both('beam_backtrack', be, [preds, backs, 4, B_, beam, torch.zeros(B_, beam, 4, dtype=torch.int64)], {}, [], int_outs=[5])


@pytest.mark.parametrize('V_', [10547, 20011])
def test_beam_topk_degenerate_and_strided_rows(be, V_):
    """Rows that defeat the threshold shortcut of the top-k kernels (all logits equal / thousands of ties at the maximum /
    -inf entries), rows taken as a column slice of a wider buffer (pitch != V, every 16-byte alignment), a vocabulary larger
    than the register-resident variant holds (20011): ids and log-probs must equal torch's log_softmax + (value desc, index
    asc) selection."""
    k, rows = 5, 12
    wide = R(rows, V_ + 13, scale=2.0)
    lg = wide[:, 3:3 + V_]                       # pitch V_+13, misaligned base
    lg[0] = 0.25                                 # all equal: top-k = the k lowest indices
    lg[1, :] = -1.0
    lg[1, 100:4100] = 7.0                        # 4000 ties at the maximum
    lg[2, 50:] = float('-inf')                   # only 50 finite entries
    lg[3, 7] = lg[3, 9000] = lg[3].max() + 2     # tie between a low and a high index
    last = torch.full((rows,), 9, dtype=torch.int64)
    ref_lp = torch.log_softmax(lg.double(), -1)
    order = torch.argsort(torch.arange(V_).expand(rows, V_) - 1e9 * 0, 1)      # placeholder for clarity: indices ascending
    # (value desc, index asc): sort by index first (stable), then by value descending (stable)
    idx = torch.argsort(lg.double(), dim=1, descending=True, stable=True)[:, :k]
    gtl, gti = torch.zeros(rows, k, device=DEV), torch.zeros(rows, k, dtype=torch.int64, device=DEV)
    be.beam_topk(wide.to(DEV)[:, 3:3 + V_], last.to(DEV), 2, k, gtl, gti)
    assert torch.equal(gti.cpu(), idx), (gti.cpu()[:4], idx[:4])
    assert (gtl.cpu().double() - ref_lp.gather(1, idx)).abs().max() < 1e-4


# ----------------------------------------------------------------------------------------------- fused region aggregation
@pytest.mark.parametrize('B_,T,TR,E', [(3, 26, 936, 2), (2, 26, 37, 2), (2, 5, 20, 1), (1, 1, 1, 1), (2, 13, 64, 2)])
def test_region_aggregate_fused(be, B_, T, TR, E):
    """csrc/region_agg.cu (layer.py:184-192 in one pass over the region activations) against the plain composition
    LayerNorm -> scores -> softmax over all T*R regions -> weighted sum and its hand-derived backward (cpu_emul).
    Y is a column slice of the two-encoder projection buffer (row pitch 2H), ragged last tile, T < 26, one encoder only.
    Tolerances: the kernels round the softmax weights (x rstd) and the gamma-folded frame vectors to bf16 for the mma;
    the emulator rounds the same frame vectors, so scores are tight (2e-3) and aggregates / gradients are 1e-2 of scale."""
    H = 1024
    scale = 1.0 / math.sqrt(2048)
    Ybuf = bf(torch.tanh(R(B_ * TR, 2 * H)))
    Y = [Ybuf[:, e * H:(e + 1) * H] for e in range(E)]
    F = [R(B_ * T, H) for _ in range(E)]
    gamma = [1 + 0.1 * R(H) for _ in range(E)]
    beta = [0.1 * R(H) for _ in range(E)]
    z = lambda *s, dt=torch.float32: [torch.zeros(*s, dtype=dt) for _ in range(E)]
    cpu = dict(agg=z(B_ * T, H), U=z(B_ * T, H), stats=z(B_ * TR, 2), St=z(B_, T, TR), tconst=z(B_ * T, 4))
    dev = lambda lst: None if lst is None else [x.to(DEV) for x in lst]
    Ybuf_d = Ybuf.to(DEV)
    Yd = [Ybuf_d[:, e * H:(e + 1) * H] for e in range(E)]
    gpu = {k: dev(v) for k, v in cpu.items()}
    # amplify the scores so that the softmax is far from uniform (scale * S of order 1-5)
    sc = scale * 40
    EM.region_aggregate_fwd(Y, F, gamma, beta, sc, T, **cpu)
    be.region_aggregate_fwd(Yd, dev(F), dev(gamma), dev(beta), sc, T, **gpu)
    torch.cuda.synchronize()
    tols = dict(agg=1e-2, U=1e-2, stats=1e-4, St=2e-3, tconst=2e-3)
    for k in cpu:
        for e in range(E):
            c, g_ = cpu[k][e], gpu[k][e].cpu()
            s_ = max(1.0, float(c.abs().max()))
            assert float((c - g_).abs().max()) <= tols[k] * s_, (k, e, float((c - g_).abs().max()), s_)
    # inference form: no optional outputs
    agg2 = dev(z(B_ * T, H))
    be.region_aggregate_fwd(Yd, dev(F), dev(gamma), dev(beta), sc, T, agg=agg2)
    for e in range(E):
        assert torch.equal(agg2[e], gpu['agg'][e])
    # ---- backward: scores pass with F := dA, then the prep + streaming pass
    dA = [R(B_ * T, H) for _ in range(E)]
    c1 = dict(St=z(B_, T, TR), tconst=z(B_ * T, 4))
    g1 = {k: dev(v) for k, v in c1.items()}
    EM.region_aggregate_fwd(Y, dA, gamma, beta, sc, T, scores_only=True, **c1)
    be.region_aggregate_fwd(Yd, dev(dA), dev(gamma), dev(beta), sc, T, scores_only=True, **g1)
    torch.cuda.synchronize()
    for k, tol in (('St', 2e-3), ('tconst', 2e-3)):
        for e in range(E):
            c, g_ = c1[k][e], g1[k][e].cpu()
            assert float((c - g_).abs().max()) <= tol * max(1.0, float(c.abs().max())), (k, e)
    dbuf = torch.zeros(B_ * TR, 2 * H, dtype=torch.bfloat16)
    dbuf_d = dbuf.to(DEV)
    c2 = dict(dpre=[dbuf[:, e * H:(e + 1) * H] for e in range(E)], dF=z(B_ * T, H), dgamma=[R(H) for _ in range(E)],
              dbeta=[R(H) for _ in range(E)], dbias=[R(H) for _ in range(E)])
    g2 = dict(dpre=[dbuf_d[:, e * H:(e + 1) * H] for e in range(E)], dF=dev(c2['dF']), dgamma=dev(c2['dgamma']),
              dbeta=dev(c2['dbeta']), dbias=dev(c2['dbias']))
    # both sides start from the emulator's forward tensors, so only the backward kernels are compared
    args = [cpu['stats'], cpu['St'], c1['St'], F, dA, cpu['U'], cpu['tconst'], c1['tconst'], gamma, beta]
    EM.region_aggregate_bwd(Y, *args, sc, T, **c2)
    work = [torch.empty(be.region_aggregate_bwd_workspace(B_, T, TR), dtype=torch.uint8, device=DEV) for _ in range(E)]
    be.region_aggregate_bwd(Yd, *[dev(a) for a in args], sc, T, work=work, **g2)
    torch.cuda.synchronize()
    for k, tol in (('dpre', 2e-2), ('dF', 1e-2), ('dgamma', 1e-2), ('dbeta', 1e-4), ('dbias', 1e-2)):
        for e in range(E):
            c, g_ = c2[k][e].float(), g2[k][e].float().cpu()
            s_ = max(1e-6, float(c.abs().max()))
            assert float((c - g_).abs().max()) <= tol * s_, (k, e, float((c - g_).abs().max()), s_)
    if E < 2:       # the other encoder's half of the gradient buffer is untouched
        assert float(dbuf_d[:, H:].abs().max()) == 0.0


@pytest.mark.parametrize('M,N,K,splitk,atomic', [(64, 4096, 2864, 4, False), (64, 6144, 4608, 1, True), (64, 1024, 1024, 1, False),
                                                 (200, 384, 520, 1, False), (640, 10547, 1536, 1, False)])
def test_gemm_static_operand_early_fetch_is_bit_identical(be, M, N, K, splitk, atomic):
    """DLSG_GEMM_B_STATIC (weight tiles requested before the programmatic-dependent-launch wait) changes WHEN the tiles
    travel, not what is computed: bit-identical output with and without it, directly behind a kernel that rewrites the
    activations (the dependent operand), in the swap-AB (M <= 64: the weight rides the 128-row side), plain, split-K
    and atomic-accumulate forms."""
    Kp = (K + 7) // 8 * 8
    w = bf(R(N, Kp))[:, :K]
    x_src = R(M, Kp)
    outs = []
    monkey = ops.STATIC_PREFETCH
    ops.STATIC_PREFETCH = True                      # off by default (no measured gain); the capability stays tested
    for static in (False, True):
        x = torch.empty(M, Kp, dtype=torch.bfloat16, device=DEV)
        wd = w.to(DEV)
        if splitk > 1:
            o = torch.zeros(splitk, M, N, device=DEV)
        else:
            o = torch.zeros(M, N, device=DEV)
        xs = x_src.to(DEV)
        d1, d2 = torch.zeros(64, device=DEV), torch.zeros(64, device=DEV)
        for rep in range(3):                              # the activations (operand A) are rewritten right before every GEMM
            be.convert(xs * (rep + 1), dst=x)
            if atomic:
                o.zero_()
            # the backend drops the request directly behind a conversion launch (it could be writing the weight copy):
            # put an ordinary kernel in between, as the recurrent loops have (cell / attention kernel before each GEMM)
            be.axpby(d1, 1.0, d2, 0.0)
            assert be._writer_last.get(torch.cuda.current_stream().cuda_stream) is False
            be.gemm(x[:, :K], wd, o, splitk=splitk, atomic=atomic, b_static=static)
        torch.cuda.synchronize()
        outs.append(o.sum(0) if splitk > 1 else o)
    ops.STATIC_PREFETCH = monkey
    if atomic:                                            # split order is not fixed: last-bit noise in either run
        assert float((outs[0] - outs[1]).abs().max()) <= 1e-3 * float(outs[0].abs().max())
    else:
        assert torch.equal(outs[0], outs[1])
    ref = (x_src * 3).to(torch.bfloat16).float()[:, :K] @ w.float().t()
    assert float((outs[1].cpu() - ref).abs().max()) <= 2e-3 * float(ref.abs().max())


@pytest.mark.parametrize('B_,H,nd,first', [(64, 1024, 2, False), (64, 1024, 2, True), (5, 128, 1, False), (33, 256, 2, False),
                                           (64, 512, 3, False), (48, 512, 4, True)])
def test_lstm_step_fused(be, B_, H, nd, first):
    """csrc/lstm_step.cu: recurrent product + LSTM cell of one time step in one launch (two directions), against the
    emulator's fp32 product of the same bf16 operands; strided views as the BiLSTM passes them (h of a (B, T, H) operand
    buffer, the direction's half of the (B, T, 8H) input projection, the layer-output slice), batch < 64, first step."""
    T = 3
    W = [bf(R(4 * H, H, scale=0.05)) for _ in range(nd)]
    hbuf = bf(R(nd, B_, T, H))
    Gin = R(B_, T, nd * 4 * H)
    c_in = [R(B_, H) for _ in range(nd)]
    outbuf, opbuf = torch.zeros(B_, T, nd * H), torch.zeros(nd, B_, T, H, dtype=torch.bfloat16)

    def args(dev_):
        mv = (lambda x: x.to(DEV)) if dev_ else (lambda x: x)
        hb, gi, ob, pb = mv(hbuf), mv(Gin), mv(outbuf.clone()), mv(opbuf.clone())
        out = dict(c_out=[mv(torch.zeros(B_, H)) for _ in range(nd)], acts=[mv(torch.zeros(B_, 4 * H)) for _ in range(nd)],
                   h_out=[ob[:, 1, d * H:(d + 1) * H] for d in range(nd)], h_op=[pb[d, :, 2] for d in range(nd)])
        a = ([mv(w) for w in W], None if first else [hb[d, :, 1] for d in range(nd)],
             [gi[:, 1, d * 4 * H:(d + 1) * 4 * H] for d in range(nd)], [mv(c) for c in c_in])
        return a, out, (ob, pb)
    (ac, oc, bc), (ag, og, bg) = args(False), args(True)
    EM.lstm_step_fwd(*ac, oc['c_out'], oc['acts'], oc['h_out'], oc['h_op'])
    be.lstm_step_fwd(*ag, og['c_out'], og['acts'], og['h_out'], og['h_op'])
    torch.cuda.synchronize()
    for k in ('c_out', 'acts'):
        for d in range(nd):
            assert float((oc[k][d] - og[k][d].cpu()).abs().max()) <= 2e-4, (k, d)
    assert float((bc[0] - bg[0].cpu()).abs().max()) <= 2e-4                     # h into the layer-output slice, rest untouched
    assert float((bc[1].float() - bg[1].float().cpu()).abs().max()) <= 1e-2      # bf16 operand copy


def test_beam_gather_multi(be):
    """One launch re-indexes up to four state buffers by the beam back-pointers (rows of different widths / dtypes, one of
    them not 16-byte sized): equal to four single-buffer gathers."""
    B_, beam = 7, 5
    g = torch.Generator().manual_seed(3)
    back = torch.randint(0, beam, (B_, beam), generator=g)
    bufs = [bf(R(B_ * beam, 2864 + 8)), bf(R(B_ * beam, 4608)), R(B_ * beam, 1024), R(B_ * beam, 1537)]
    want = [torch.zeros_like(x) for x in bufs]
    for s_, d_ in zip(bufs, want):
        EM.beam_gather(s_, d_, back, B_, beam)
    got = [torch.zeros_like(x).to(DEV) for x in bufs]
    be.beam_gather_multi([(s_.to(DEV), d_) for s_, d_ in zip(bufs, got)], back.to(DEV), B_, beam)
    torch.cuda.synchronize()
    for w, g_ in zip(want, got):
        assert torch.equal(w, g_.cpu())
