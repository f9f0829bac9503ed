"""End-to-end parity of the drop-in models (real sm_100a kernels through the C-ABI) against
(a) the golden vectors produced by the unmodified reference and (b) the CPU oracle at larger sizes.

Tolerances:
  fp32 mode : logits / nodes max-abs <= 1e-4 (x output scale), loss |d| <= 1e-4, grads rel-L2 <= 1e-3,
              greedy / beam token ids bit-exact.
  bf16 mode : logits rel-L2 <= 2e-2, loss |d| <= 2e-2, grads rel-L2 <= 6e-2 (26-step recurrence);
              tokens compared up to the first position whose reference top-2 log-prob gap < 5e-2.
"""
import contextlib
import io
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from dlsg import synth, linalg as la
from oracle import dlsg_oracle as O

DEV = 'cuda'


@pytest.fixture(autouse=True)
def _gpu_only():
    if not torch.cuda.is_available():
        pytest.skip('no GPU')
    yield
    la.set_precision('bf16')


def build(cls, args, V):
    import models.model as M
    with contextlib.redirect_stdout(io.StringIO()):
        net = getattr(M, cls)(args, synth.Vocab(V))
    synth.fill_state_dict(net)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net.to(DEV), sd


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


SMALL = [('capgnn_small_msr', synth.small_args(), 37, 3),
         ('capgnn_small_msvd', synth.small_args(num_proposals=8, num_topk=3, decode_hidden_size=64, dataset='msvd', num_obj=5), 41, 2)]


@pytest.mark.parametrize('tag,args,V,B', SMALL)
def test_golden_fp32(golden_dir, tag, args, V, B):
    """fp32 mode against the reference's own outputs (tests/golden/*.npz)."""
    la.set_precision('fp32')
    g = np.load(os.path.join(golden_dir, tag + '.npz'))
    net, _ = build('CapGnnModel', args, V)
    net.eval()
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=12)
    fr, rg, cp = frames.to(DEV), regions.to(DEV), caps.to(DEV)
    out, obj, mot, alpha = net(fr, rg, cp, args.max_words, 1.0)
    assert np.abs(out.detach().cpu().numpy() - g['logits']).max() < 1e-4
    assert np.abs(obj.detach().cpu().numpy() - g['obj']).max() < 1e-4
    assert np.abs(mot.detach().cpu().numpy() - g['mot']).max() < 1e-4
    assert np.abs(alpha.detach().cpu().numpy() - g['alpha']).max() < 1e-4
    loss = O.packed_ce_loss(out, cp, lens)
    assert abs(loss.item() - g['loss'][0]) < 1e-4
    loss.backward()
    for k, p in net.named_parameters():
        if 'gnone.' + k in g.files:
            assert p.grad is None, k
            continue
        ref = float(g['gnorm.' + k][0])
        gn = float(p.grad.double().norm())
        head = p.grad.reshape(-1)[:8].cpu().numpy()
        assert abs(gn - ref) <= 1e-3 * max(ref, 1e-4) or abs(gn - ref) < 1e-7, (k, gn, ref)
        assert np.abs(head - g['ghead.' + k]).max() <= 1e-3 * max(1e-3, np.abs(g['ghead.' + k]).max()) + 1e-7, k
    with torch.no_grad():
        random.seed(12)
        out6 = net(fr, rg, cp, args.max_words, 0.6)[0]
        assert np.abs(out6.cpu().numpy() - g['logits_tf06']).max() < 1e-4
        net.update_beam_size(1)
        assert np.array_equal(net(fr, rg, None)[0].cpu().numpy(), g['greedy'])
        for bm in (5, 3):
            net.update_beam_size(bm)
            assert np.array_equal(net(fr, rg, None)[0].cpu().numpy(), g['beam%d' % bm])
            # the generic AllenNLP-contract entry (BeamSearch.search + Decoder.beam_step) too
            dec = net.decoder
            o, m = net.encoder(fr, rg)
            dec.batch_size = B
            glob = torch.cat([o.mean(1), m.mean(1)], -1)
            z = lambda h: o.new_zeros(B, h)
            st = {'query_lstm_h': z(args.query_hidden_size), 'query_lstm_c': z(args.query_hidden_size),
                  'lang_lstm_h': z(args.decode_hidden_size), 'lang_lstm_c': z(args.decode_hidden_size),
                  'cnn_feats': o, 'global_feat': glob, 'cnn_feats_2': m}
            start = torch.full((B,), 1, dtype=torch.long, device=DEV)
            preds, lp = dec.beam_search.search(start, st, dec.beam_step)
            assert np.array_equal(preds.cpu().numpy(), g['beam%d_all' % bm])
            assert np.abs(lp.cpu().numpy() - g['beam%d_lp' % bm]).max() < 1e-3


def test_golden_baseline1_fp32(golden_dir):
    la.set_precision('fp32')
    args, V, B = synth.small_args(decode_hidden_size=52), 37, 3
    g = np.load(os.path.join(golden_dir, 'baseline1_small.npz'))
    net, _ = build('CapBaseline1', args, V)
    net.eval()
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=13)
    fr, rg, cp = frames.to(DEV), regions.to(DEV), caps.to(DEV)
    out = net(fr, rg, cp, args.max_words, 1.0)[0]
    assert np.abs(out.detach().cpu().numpy() - g['logits']).max() < 1e-4
    O.packed_ce_loss(out, cp, lens).backward()
    for k, p in net.named_parameters():
        if 'gnone.' + k in g.files:
            continue
        ref = float(g['gnorm.' + k][0])
        assert abs(float(p.grad.double().norm()) - ref) <= 1e-3 * max(ref, 1e-4) + 1e-7, k
    with torch.no_grad():
        net.update_beam_size(1)
        assert np.array_equal(net(fr, rg, None)[0].cpu().numpy(), g['greedy'])
        net.update_beam_size(5)
        assert np.array_equal(net(fr, rg, None)[0].cpu().numpy(), g['beam5'])


def _token_parity(ours, ref_ids, ref_logp_fn, tol_gap):
    """ids equal up to the first position where the reference's own top-2 gap is below tol_gap."""
    ours, ref_ids = ours.cpu(), ref_ids.cpu()
    n_exact = 0
    for b in range(ref_ids.shape[0]):
        L = min(ours.shape[1], ref_ids.shape[1])
        for t in range(L):
            if ours[b, t] != ref_ids[b, t]:
                gap = ref_logp_fn(b, t)
                assert gap < tol_gap, ('token mismatch with a decisive reference gap', b, t, gap)
                break
        else:
            n_exact += 1
    return n_exact


FULL = [('msr', synth.msr_args(), 10547, 2), ('msvd', synth.msvd_args(), 9468, 2),
        ('msr_dm1024', synth.msr_args(m_feature_size=1024), 10547, 2)]


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
@pytest.mark.parametrize('tag,args,V,B', FULL)
def test_full_width_vs_oracle(tag, args, V, B, prec):
    """Reference widths (1536/2048/1024-d, 26 frames, 36 or 16 regions, V~1e4) against the CPU oracle."""
    la.set_precision(prec)
    net, sd = build('CapGnnModel', args, V)
    net.eval()
    for v in sd.values():
        v.requires_grad_(True)
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=12)
    fr, rg, cp = frames.to(DEV), regions.to(DEV), caps.to(DEV)
    out, obj, mot, alpha = net(fr, rg, cp, args.max_words, 1.0)
    ro, robj, rmot, ralpha = O.cap_gnn_forward(sd, frames, regions, caps, args.max_words, 1.0, args.a_feature_size)
    loss = O.packed_ce_loss(out, cp, lens)
    rloss = O.packed_ce_loss(ro, caps, lens)
    loss.backward()
    rloss.backward()
    o_ = out.detach().cpu()
    if prec == 'fp32':
        assert (o_ - ro).abs().max() < 1e-4 * max(1.0, float(ro.abs().max()))
        assert (obj.detach().cpu() - robj).abs().max() < 1e-4
        assert abs(loss.item() - rloss.item()) < 1e-4
        gtol = 1e-3
    else:
        assert rel(o_, ro.detach()) < 2e-2
        assert rel(obj.detach().cpu(), robj.detach()) < 2e-2 and rel(mot.detach().cpu(), rmot.detach()) < 2e-2
        assert abs(loss.item() - rloss.item()) < 2e-2
        gtol = 6e-2
    bad = []
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        ref = sd[k].grad
        e = rel(p.grad.cpu(), ref)
        small = float((p.grad.cpu() - ref).abs().max()) < (1e-7 if prec == 'fp32' else 1e-5)
        # attention K/Q weight gradients are sums of softmax-Jacobian terms that cancel (sum_p dlogit_p = 0):
        # under bf16 operands their relative error is ~1.5x the other parameters'
        kq = (k.endswith('.K.weight') or k.endswith('.Q.weight')) and prec == 'bf16'
        if not (e < gtol * (1.7 if kq else 1.0) or small):
            bad.append((k, e))
    assert not bad, bad
    # decoding: greedy + beam-5 tokens vs the oracle
    with torch.no_grad():
        net.update_beam_size(1)
        g_ids = net(fr, rg, None)[0]
        r_ids = O.cap_gnn_forward(sd, frames, regions, None, args.max_words, 1.0, args.a_feature_size, beam_size=1)[0]
        net.update_beam_size(5)
        b_ids = net(fr, rg, None)[0]
        rb_ids, _, _ = O.decoder_beam(sd, 'decoder', robj.detach(), rmot.detach(), args.max_words, 5)
        if prec == 'fp32':
            assert torch.equal(g_ids.cpu(), r_ids)
            assert torch.equal(b_ids.cpu(), rb_ids)
        else:
            # waiver rule: a mismatch is only allowed where the reference's top-2 logit gap is below the bf16 tolerance
            sd_ng = {k: v.detach() for k, v in sd.items()}

            def gap_fn(b, t, ids=r_ids):
                caps_ref = ids[b:b + 1]
                lo = O.decoder_forward(sd_ng, 'decoder', robj.detach()[b:b + 1], rmot.detach()[b:b + 1], caps_ref, t + 1, 1.0)[0]
                top2 = torch.topk(torch.log_softmax(lo[0, t], -1), 2)[0]
                return float(top2[0] - top2[1])
            _token_parity(g_ids, r_ids, gap_fn, 5e-2)
            # beam-5 in bf16 (the precision the captions/s figure is quoted in): identical best sequences, except where the
            # fp32 oracle itself scores our sequence within the tolerance of its own best one (near-tie of two hypotheses)
            from test_parity_baseline_sizes_gpu import _seq_logprob
            b_cpu = b_ids.cpu()
            n = min(b_cpu.shape[1], rb_ids.shape[1])
            for b in range(B):
                if torch.equal(b_cpu[b, :n], rb_ids[b, :n]):
                    continue
                lp_o = _seq_logprob(sd_ng, robj.detach()[b:b + 1], rmot.detach()[b:b + 1], b_cpu[b])
                lp_r = _seq_logprob(sd_ng, robj.detach()[b:b + 1], rmot.detach()[b:b + 1], rb_ids[b])
                assert lp_o > lp_r - 5e-2, ('bf16 beam-5 sequence the reference scores clearly below its own best', b, lp_o, lp_r)


def test_batch64_rows_are_independent_and_train_mode_runs():
    """BASELINE config-2 size (B=64, MSR widths): rows of the batch are independent clips, so the first 4
    rows of a B=64 run must equal a B=4 run (size-independent property at full size); then a train-mode
    (dropout on) step must give finite loss/grads and different logits from eval mode."""
    la.set_precision('bf16')
    args, V = synth.msr_args(), 10547
    net, _ = build('CapGnnModel', args, V)
    net.eval()
    frames, regions, caps, lens = synth.make_inputs(64, args, V, seed=3)
    fr, rg, cp = frames.to(DEV), regions.to(DEV), caps.to(DEV)
    with torch.no_grad():
        o64 = net(fr, rg, cp, 26, 1.0)[0]
        o4 = net(fr[:4].contiguous(), rg[:4].contiguous(), cp[:4].contiguous(), 26, 1.0)[0]
    # not bit-equal: tile / split-K choices depend on the batch size, so fp32 summation order and hence the bf16
    # roundings of intermediate activations differ (bf16 eps = 3.9e-3)
    assert rel(o64[:4], o4) < 1e-2
    net.train()
    out = net(fr, rg, cp, 26, 1.0)[0]
    loss = O.packed_ce_loss(out, cp, lens)
    loss.backward()
    assert torch.isfinite(loss)
    for k, p in net.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), k
    assert rel(out.detach()[:4], o4) > 1e-3


def test_fused_masked_ce_matches_packed_ce():
    from dlsg import losses
    B, L, V = 8, 26, 10547
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(B, L, V, generator=g).to(DEV).requires_grad_(True)
    _, _, caps, lens = synth.make_inputs(B, synth.msr_args(), V, seed=5)
    cp = caps.to(DEV)
    loss = losses.packed_cross_entropy(logits, cp, lens)
    loss.backward()
    l2 = logits.detach().clone().requires_grad_(True)
    ref = O.packed_ce_loss(l2, cp, lens)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5
    assert (logits.grad - l2.grad).abs().max() < 1e-7


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
@pytest.mark.parametrize('tag,P,K', [('disc_small_msr', 5, 5), ('disc_small_msvd', 8, 3)])
def test_disc_v2_and_wgan_gp_golden(golden_dir, tag, P, K, prec):
    """DiscV2 forward (3 input kinds), first-order grads and the WGAN-GP double backward against the reference's
    golden vectors.  fp32: 1e-4; bf16: the critic scores are O(0.1), tolerance 2e-2 absolute, penalty 5e-2."""
    from test_disc_cpu import disc_case
    la.set_precision(prec)
    g, net, r, f, m, gn, gp, loss_d, f_in = disc_case(golden_dir, tag, P, K, dev=DEV)
    tol = 1e-4 if prec == 'fp32' else 2e-2
    assert np.abs(r.detach().cpu().numpy() - g['r_logit']).max() < tol
    assert np.abs(f.detach().cpu().numpy() - g['f_logit']).max() < tol
    assert np.abs(m.detach().cpu().numpy() - g['m_logit']).max() < tol
    assert np.abs(gn.detach().cpu().numpy() - g['gnorm_mixed']).max() < (1e-3 if prec == 'fp32' else 5e-2)
    assert abs(loss_d.item() - g['loss_d'][0]) < (1e-3 if prec == 'fp32' else 0.3)
    if prec == 'fp32':
        assert np.abs(f_in.grad.cpu().numpy() - g['dfake']).max() < 1e-5
        for k, p in net.named_parameters():
            ref = float(g['gnorm.' + k][0])
            assert abs(float(p.grad.double().norm()) - ref) <= 2e-3 * max(ref, 1e-3), k
