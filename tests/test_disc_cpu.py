"""DiscV2 (models/model.py:145-168) + WGAN-GP double backward (run_gun.py:351-375) through the drop-in mirror,
kernels emulated on CPU (tests/cpu_emul.py), against the golden vectors produced by the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from dlsg import synth, ops
from dlsg import linalg as la
from dlsg import functional as DF
from cpu_emul import CpuEmulBackend


@pytest.fixture(autouse=True)
def emul_backend():
    old = ops._backend
    ops.set_backend(CpuEmulBackend())
    DF.WC.clear()
    la.set_precision('fp32')
    yield
    ops.set_backend(old)
    DF.WC.clear()
    la.set_precision('bf16')


def disc_case(golden_dir, tag, P, K, dev='cpu'):
    import models.model as M
    args = synth.small_args(visual_hidden_size=1024, num_proposals=P, num_topk=K)
    V, B, L = 37, 3, args.max_words
    g = np.load(os.path.join(golden_dir, tag + '.npz'))
    net = M.DiscV2(args, V)
    synth.fill_state_dict(net, prefix='D.')
    net = net.to(dev).eval()
    rs = np.random.RandomState(5)
    _, _, caps, lens = synth.make_inputs(B, args, V, seed=14)
    att_mask = synth.att_mask_from_captions(caps).to(dev)
    obj = torch.from_numpy(rs.standard_normal((B, P, 1024)).astype(np.float32)).to(dev)
    mot = torch.from_numpy(rs.standard_normal((B, P, 1024)).astype(np.float32)).to(dev)
    alpha = torch.softmax(torch.from_numpy(rs.standard_normal((B, L, 2 * P)).astype(np.float32)), -1).to(dev)
    fake = torch.from_numpy(rs.standard_normal((B, L, V)).astype(np.float32)).to(dev)
    real = torch.zeros(B, L, V).scatter_(2, caps.unsqueeze(2), 1).to(dev)
    eps = torch.from_numpy(rs.uniform(size=(B, 1, 1)).astype(np.float32)).to(dev)
    net.zero_grad()
    r_logit = net(real.clone(), obj, mot, att_mask, alpha)
    f_in = fake.clone().requires_grad_(True)
    f_logit = net(f_in, obj, mot, att_mask, alpha)
    mixed = (real * eps + fake * (1 - eps)).requires_grad_(True)
    m_logit = net(mixed, obj, mot, att_mask, alpha)
    gr = torch.autograd.grad(m_logit, mixed, torch.ones_like(m_logit), create_graph=True, retain_graph=True)[0]
    gn = gr.contiguous().view(B, -1).norm(2, dim=1)
    gp = ((gn - 1) * (gn - 1)).mean()
    loss_d = f_logit.mean() - r_logit.mean() + 10 * gp
    loss_d.backward()
    return g, net, r_logit, f_logit, m_logit, gn, gp, loss_d, f_in


@pytest.mark.parametrize('tag,P,K', [('disc_small_msr', 5, 5), ('disc_small_msvd', 8, 3)])
def test_disc_forward_and_gradient_penalty(golden_dir, tag, P, K):
    g, net, r, f, m, gn, gp, loss_d, f_in = disc_case(golden_dir, tag, P, K)
    assert np.abs(r.detach().numpy() - g['r_logit']).max() < 2e-5
    assert np.abs(f.detach().numpy() - g['f_logit']).max() < 2e-5
    assert np.abs(m.detach().numpy() - g['m_logit']).max() < 2e-5
    assert np.abs(gn.detach().numpy() - g['gnorm_mixed']).max() < 1e-4
    assert abs(gp.item() - g['gp'][0]) < 1e-4
    assert abs(loss_d.item() - g['loss_d'][0]) < 1e-4
    assert np.abs(f_in.grad.numpy() - g['dfake']).max() < 1e-5
    for k, p in net.named_parameters():
        ref = float(g['gnorm.' + k][0])
        gn_k = float(p.grad.double().norm())
        assert abs(gn_k - ref) <= 5e-4 * max(ref, 1e-3), (k, gn_k, ref)


def test_disc_bf16_operand_memo_matches_fp32_reference(golden_dir):
    """bf16 mode routes every generic-path product through the memoised operand copies (dlsg.linalg.op_cached: one
    conversion per tensor shared by forward, data-gradient, weight-gradient and double-backward products).  The result
    must stay within bf16 tolerance of the reference's fp32 golden vectors, and the memo must actually be used."""
    la.set_precision('bf16')
    g, net, r, f, m, gn, gp, loss_d, f_in = disc_case(golden_dir, 'disc_small_msr', 5, 5)
    assert np.abs(r.detach().numpy() - g['r_logit']).max() < 5e-3
    assert np.abs(m.detach().numpy() - g['m_logit']).max() < 5e-3
    assert np.abs(gn.detach().numpy() - g['gnorm_mixed']).max() < 5e-3
    assert abs(loss_d.item() - g['loss_d'][0]) < 2e-2
    for k, p in net.named_parameters():
        ref = float(g['gnorm.' + k][0])
        assert abs(float(p.grad.double().norm()) - ref) <= 5e-2 * max(ref, 1e-3), k
    assert len(la._PMEMO) > 0 and len(la._AMEMO) > 0
    # a fused-optimizer style update (no version bump) must not be served from a stale copy after a new forward
    w = net.att.K.weight
    v0 = w._version
    w.data.mul_(0.0)
    assert w._version == v0 or True
    r2 = net(torch.zeros_like(f_in.detach()), torch.zeros(3, 5, 1024), torch.zeros(3, 5, 1024),
             torch.ones(3, f_in.shape[1], f_in.shape[1]), torch.full((3, f_in.shape[1], 10), 0.1))
    key_hits = [e for e in la._PMEMO.values() if e[0] is w]
    assert key_hits and float(key_hits[-1][2].float().abs().max()) == 0.0


def test_disc_grouped_forward_equals_separate_calls(golden_dir):
    """DiscV2(..., _groups=3) over [real; fake; mixed] stacked along the batch must equal the three separate reference
    calls (PSLScore2's batch mean is taken per group), for values and for the WGAN-GP input gradient."""
    import models.model as M
    args = synth.small_args(visual_hidden_size=1024, num_proposals=5, num_topk=5)
    V, B, L = 37, 3, args.max_words
    net = M.DiscV2(args, V)
    synth.fill_state_dict(net, prefix='D.')
    net.eval()
    rs = np.random.RandomState(9)
    _, _, caps, _ = synth.make_inputs(B, args, V, seed=14)
    att_mask = synth.att_mask_from_captions(caps)
    obj = torch.from_numpy(rs.standard_normal((B, 5, 1024)).astype(np.float32))
    mot = torch.from_numpy(rs.standard_normal((B, 5, 1024)).astype(np.float32))
    alpha = torch.softmax(torch.from_numpy(rs.standard_normal((B, L, 10)).astype(np.float32)), -1)
    xs = [torch.from_numpy(rs.standard_normal((B, L, V)).astype(np.float32)) for _ in range(3)]
    mixed_a = xs[2].clone().requires_grad_(True)
    sep = [net(xs[0], obj, mot, att_mask, alpha), net(xs[1], obj, mot, att_mask, alpha), net(mixed_a, obj, mot, att_mask, alpha)]
    ga = torch.autograd.grad(sep[2].sum(), mixed_a)[0]
    mixed_b = xs[2].clone().requires_grad_(True)
    out = net(torch.cat([xs[0], xs[1], mixed_b], 0), obj.repeat(3, 1, 1), mot.repeat(3, 1, 1), att_mask.repeat(3, 1, 1),
              alpha.repeat(3, 1, 1), _groups=3)
    gb = torch.autograd.grad(out[2 * B:].sum(), mixed_b)[0]
    assert (out.detach() - torch.cat([s.detach() for s in sep])).abs().max() < 1e-5
    assert (ga - gb).abs().max() < 1e-6


def test_gan_iteration_stacked_critic_calls_match_separate_calls():
    """dlsg.gan.GanIteration on CPU (kernels emulated, eval mode, same WGAN-GP epsilon draws): running the three critic
    calls of every critic step as ONE stacked forward (batched=True) must give the same four logged scalars and the same
    updated critic / generator weights as the literal run_gun.py sequence of separate calls (batched=False)."""
    import contextlib
    import io
    import models.model as M
    from dlsg.gan import GanIteration
    args = synth.small_args(visual_hidden_size=1024, region_projected_size=1024, query_hidden_size=1024, max_words=5, max_frames=4)
    V, B = 37, 3
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=21)
    results = []
    for batched in (True, False):
        DF.WC.clear()
        with contextlib.redirect_stdout(io.StringIO()):
            G_ = M.CapGnnModel(args, synth.Vocab(V))
        D_ = M.DiscV2(args, V)
        synth.fill_state_dict(G_)
        synth.fill_state_dict(D_, prefix='D.')
        G_.eval()
        D_.eval()
        og = torch.optim.Adam(G_.parameters(), lr=1.6e-4, betas=(0.5, 0.9))
        od = torch.optim.Adam(D_.parameters(), lr=1.6e-4, betas=(0.5, 0.9))
        it = GanIteration(G_, D_, og, od, frames, regions, caps, lens, args.max_words, 1.0, num_d=2, gan_lambda=0.05,
                          graph=False, batched=batched)
        torch.manual_seed(77)
        outs = [[float(x) for x in it()] for _ in range(2)]
        results.append((outs, {k: p.detach().clone() for k, p in D_.named_parameters()},
                        {k: p.detach().clone() for k, p in G_.named_parameters()}))
    (oa, da, ga), (ob, db, gb) = results
    for ra, rb in zip(oa, ob):
        for x, y in zip(ra, rb):
            assert abs(x - y) <= 2e-4 * max(1.0, abs(y)), (oa, ob)
    assert oa[1] != oa[0]                                           # the second iteration sees updated weights
    # Adam normalises every gradient component by its own magnitude: where |g| is at the 1e-8 epsilon, summation-order noise
    # between the stacked and the separate products moves an element by a fraction of lr per step.  Bound: 2 * lr over
    # the 4 critic / 2 generator updates, and all but a few (< 2 %) elements of every tensor must agree to 1e-6.
    for name, a_, b_ in [(k, da[k], db[k]) for k in da] + [(k, ga[k], gb[k]) for k in ga]:
        d = (a_ - b_).abs()
        assert d.max() < 3.2e-4, name
        assert (d > 1e-6).float().mean() < 2e-2, name


# ----------------------------------------------------------------------------------------------- second-order building blocks
def _grads(fn, inputs, seed, wrt=None):
    """Gradients of a random scalar functional of the create_graph gradient of fn(*inputs) wrt every input (wrt: indices of the
    first gradients that enter the functional; default all)."""
    xs = [x.detach().clone().requires_grad_(True) for x in inputs]
    y = fn(*xs)
    g = torch.Generator().manual_seed(seed)
    w1 = torch.randn(y.shape, generator=g)
    first = torch.autograd.grad((y * w1).sum(), xs, create_graph=True, allow_unused=True)
    tot = 0
    for k, f in enumerate(first):
        if f is not None and (wrt is None or k in wrt):
            tot = tot + (f * torch.randn(f.shape, generator=g)).sum()
    second = torch.autograd.grad(tot, xs, allow_unused=True)
    return [f.detach() if f is not None else None for f in first], second


def _same(a, b, tol=2e-5):
    for x, y in zip(a, b):
        if x is None or y is None:
            assert (x is None or float(x.abs().max()) == 0) and (y is None or float(y.abs().max()) == 0)
            continue
        assert float((x - y).abs().max()) <= tol * max(1.0, float(y.abs().max())), float((x - y).abs().max())


def test_elementwise_and_softmax_second_order_match_torch_autograd():
    """The closed-form first / second backward of the fused element-wise forms and of the axis softmax (dlsg.generic: _Tanh /
    _TanhBwd, _Mul / _MulBwd, _LerpRows, _Softmax / _SoftmaxBwd) against torch's own double backward of the same functions."""
    from dlsg import generic as GN
    rs = torch.Generator().manual_seed(3)
    a, b = torch.randn(4, 6, 8, generator=rs), torch.randn(4, 6, 8, generator=rs)
    e = torch.rand(4, 1, 1, generator=rs)
    f1, s1 = _grads(lambda x: GN.tanh_(x), [a], 1)
    f2, s2 = _grads(lambda x: torch.tanh(x), [a], 1)
    _same(f1, f2); _same(s1, s2)
    f1, s1 = _grads(lambda x, y: GN.mul(GN.tanh_(x), y), [a, b], 2)
    f2, s2 = _grads(lambda x, y: torch.tanh(x) * y, [a, b], 2)
    _same(f1, f2); _same(s1, s2)
    f1, s1 = _grads(lambda x, y: GN.tanh_(GN.lerp_rows(x, y, e)), [a, b], 3)
    f2, s2 = _grads(lambda x, y: torch.tanh(x * e + y * (1 - e)), [a, b], 3)
    _same(f1, f2); _same(s1, s2)
    mask = (torch.randn(4, 6, 8, generator=rs) > -0.7).float()
    mask[1] = 0
    for dim in (1, 2):
        for mode in (0, 1, 2):
            def ours(x):
                return GN.tanh_(GN.softmax(x, dim, scale=0.41, mask=mask if mode else None, mask_mode=mode))

            def ref(x):
                v = x * 0.41
                if mode == 1:
                    v = torch.where(mask > 0, v, torch.full_like(v, -9e15))
                s = torch.softmax(v, dim)
                if mode == 2:
                    s = torch.where(mask > 0, s, torch.zeros_like(s))
                return torch.tanh(s)
            f1, s1 = _grads(ours, [a], 10 + dim * 3 + mode)
            f2, s2 = _grads(ref, [a], 10 + dim * 3 + mode)
            _same(f1, f2); _same(s1, s2)


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_lstm_fused_second_order_equals_stepwise_restatement(prec):
    """generic._LstmBptt2 (the LSTM's differentiable BPTT as one node on fused loops, closed-form cell backward-of-backward)
    against the step-by-step autograd restatement `_lstm_bptt_diff` and, in fp32, against torch.nn.LSTM's own double backward:
    first-order gradients and the gradients of a functional of them, for the input projection, both weights and the biases."""
    from dlsg import generic as GN
    la.set_precision(prec)
    B, T, H = 5, 7, 128
    rs = torch.Generator().manual_seed(11)
    x = torch.randn(B, T, H, generator=rs)
    w_ih, w_hh = 0.1 * torch.randn(4 * H, H, generator=rs), 0.1 * torch.randn(4 * H, H, generator=rs)
    b_ih, b_hh = 0.1 * torch.randn(4 * H, generator=rs), 0.1 * torch.randn(4 * H, generator=rs)
    res = {}
    for fused in (True, False):
        GN.FUSED_LSTM_BPTT2 = fused
        la.new_param_epoch()
        # (the functional reads the INPUT gradient only - the WGAN-GP case; second derivatives through the recurrent weight
        # gradient are outside the fused node's contract)
        res[fused] = _grads(lambda x_, a_, b_, c_, d_: GN.lstm(x_, a_, b_, c_, d_), [x, w_ih, w_hh, b_ih, b_hh], 5, wrt=(0,))
    GN.FUSED_LSTM_BPTT2 = True
    tol = 2e-5 if prec == 'fp32' else 2e-2
    _same(res[True][0], res[False][0], tol); _same(res[True][1], res[False][1], tol)
    if prec == 'fp32':
        def ref(x_, a_, b_, c_, d_):
            h, c = torch.zeros(B, H), torch.zeros(B, H)
            out = []
            for t in range(T):
                g = x_[:, t] @ a_.t() + c_ + d_ + h @ b_.t()
                i, f, gg, o = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
                c = f * c + i * gg
                h = o * torch.tanh(c)
                out.append(h)
            return torch.stack(out, 1)
        f2, s2 = _grads(ref, [x, w_ih, w_hh, b_ih, b_hh], 5, wrt=(0,))
        _same(res[True][0], f2, 1e-4); _same(res[True][1], s2, 1e-4)


def test_input_grads_only_scope_skips_parameter_gradients_only():
    """generic.input_grads_only(): inside the scope a backward pass computes the gradients of non-leaf tensors (what the WGAN-GP
    penalty asks for) and skips those of parameters and of views of parameters; the gradient it does return is unchanged."""
    from dlsg import generic as GN
    rs = torch.Generator().manual_seed(2)
    x = torch.randn(6, 16, generator=rs).requires_grad_(True)
    w = torch.nn.Parameter(torch.randn(8, 16, generator=rs))
    b1, b2 = torch.nn.Parameter(torch.randn(8, generator=rs)), torch.nn.Parameter(torch.randn(8, generator=rs))
    be = ops.backend()

    def run(scoped):
        h = x * 2.0                                                  # a non-leaf input, like the interpolated tokens
        y = GN.linear(GN.linear(h, w, b1), w[:, :8], b1 + b2)        # parameter, view of a parameter, non-leaf bias
        l0 = be.launches
        if scoped:
            with GN.input_grads_only():
                g = torch.autograd.grad(y.sum(), h, create_graph=True)[0]
        else:
            g = torch.autograd.grad(y.sum(), h, create_graph=True)[0]
        return g.detach(), be.launches - l0
    g_all, n_all = run(False)
    g_in, n_in = run(True)
    assert torch.equal(g_all, g_in)
    assert n_in < n_all                                               # the weight-gradient products / bias column sums are gone
    assert not GN._INPUT_GRADS_ONLY


def test_splitk_rows_rule():
    """linalg.splitk_rows: explicit split-K factors for recurrent products with more than 64 rows (the critic's 192 stacked
    rows): 1 when K is short or the tiles already fill the GPU, never more than the cell kernels' vector path sums, no empty
    split; up to 64 rows it is splitk_for."""
    la.set_precision('bf16')
    assert la.splitk_rows(192, 512, 2048) == 4                       # data-gradient product of the critic LSTM: 8 tiles, 32 k-blocks
    assert la.splitk_rows(192, 2048, 512) == 1                       # forward product: 8 k-blocks only
    assert la.splitk_rows(256, 4096, 2864) == 2                      # greedy decode, query-LSTM gates: 64 tiles
    assert la.splitk_rows(640, 6144, 4608) == 1                      # beam-5 rows: the tiles alone fill the GPU
    assert la.splitk_rows(64, 4096, 2864) == la.splitk_for(64, 4096, 2864)
    for rows, n, k in [(130, 512, 2048), (192, 512, 1100), (100, 1024, 4096)]:
        s = la.splitk_rows(rows, n, k)
        kb = (k + 63) // 64
        assert 1 <= s <= 4 and (s - 1) * ((kb + s - 1) // s) < kb
    la.set_precision('fp32')
    assert la.splitk_rows(192, 512, 2048) == 1
