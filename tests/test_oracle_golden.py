"""Pin the CPU oracle (oracle/dlsg_oracle.py) against golden vectors produced by the reference.

The .npz files were written by tests/golden/make_golden.py, which imports the UNMODIFIED
reference modules.  fp32 tolerance: max-abs 2e-5 on logits/nodes, token ids bit-exact.
"""
import os
import random

import numpy as np
import pytest
import torch

from dlsg import synth
from oracle import dlsg_oracle as O

TOL = 2e-5


def _sd_capgnn(args, V):
    """state_dict keys+shapes of CapGnnModel, filled by dlsg.synth."""
    H, Da, Dm, Dr = args.visual_hidden_size, args.a_feature_size, args.m_feature_size, args.region_feature_size
    P, Hq, Hd, W = args.num_proposals, args.query_hidden_size, args.decode_hidden_size, args.word_size
    shapes = {}

    def lnp(k, d):
        shapes[k + '.weight'] = (d,)
        shapes[k + '.bias'] = (d,)

    def lin(k, o, i, bias=True):
        shapes[k + '.weight'] = (o, i)
        if bias:
            shapes[k + '.bias'] = (o,)

    for enc, emb in (('encoder.obj_encoder', True), ('encoder.motion_encoder', False)):
        lin(enc + '.obj_embed', H, Dr)
        lnp(enc + '.obj_norm.1', H)
        if emb:
            lin(enc + '.visual_embed', H, Da)
        lnp(enc + '.visual_norm.1', H)
        lnp(enc + '.obj_visual_norm.1', H)
        shapes[enc + '.v2l_layer.theta'] = (P, H)
        lnp(enc + '.v2l_layer.out_norm.1', H)
        lnp(enc + '.att_l2l_norm', H)
    e = 'encoder.motion_pre_encoder'
    lin(e + '.linear_embed', H, Da + Dm)
    for sfx in ('', '_reverse'):
        shapes[e + '.lstm.weight_ih_l0' + sfx] = (4 * H, H)
        shapes[e + '.lstm.weight_hh_l0' + sfx] = (4 * H, H)
        shapes[e + '.lstm.bias_ih_l0' + sfx] = (4 * H,)
        shapes[e + '.lstm.bias_hh_l0' + sfx] = (4 * H,)
    lnp(e + '.layernorm_lstm', 2 * H)
    for k in 'KQV':
        lin(e + '.self_attention.%s' % k, 2 * H, 2 * H, False)
    lin(e + '.self_attention.output_layer.0', H, 2 * H, False)
    lnp(e + '.layernorm_sa', H)
    d = 'decoder'
    shapes[d + '.word_embed.weight'] = (V, W)
    qin = H + W + Hd + H
    for k, (hh, xin) in (('query_lstm', (Hq, qin)), ('lang_lstm', (Hd, 2 * H + Hq))):
        shapes['%s.%s.weight_ih' % (d, k)] = (4 * hh, xin)
        shapes['%s.%s.weight_hh' % (d, k)] = (4 * hh, hh)
        shapes['%s.%s.bias_ih' % (d, k)] = (4 * hh,)
        shapes['%s.%s.bias_hh' % (d, k)] = (4 * hh,)
    lnp(d + '.query_lstm_layernorm', Hq)
    lnp(d + '.lang_lstm_layernorm', Hd)
    for a in ('context_att', 'context_att_2'):
        for k in 'KQV':
            lin('%s.%s.%s' % (d, a, k), H, H if k != 'Q' else Hq, False)
        lin('%s.%s.output_layer.0' % (d, a), H, H, False)
        lnp('%s.%s.output_layer.2' % (d, a), H)
    lnp(d + '.context_layernorm', Hd)
    lin(d + '.word_restore', V, Hd)
    return {k: synth.fill_tensor(k, s) for k, s in shapes.items()}


CASES = [('capgnn_small_msr', synth.small_args(), 37, 3),
         ('capgnn_small_msvd', synth.small_args(num_proposals=8, num_topk=3, decode_hidden_size=64,
                                                dataset='msvd', num_obj=5), 41, 2)]


@pytest.mark.parametrize('tag,args,V,B', CASES)
def test_capgnn_forward_loss_grads(golden_dir, tag, args, V, B):
    g = np.load(os.path.join(golden_dir, tag + '.npz'))
    sd = _sd_capgnn(args, V)
    for v in sd.values():
        v.requires_grad_(True)
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=12)
    out, obj, mot, alpha = O.cap_gnn_forward(sd, frames, regions, caps, args.max_words, 1.0, args.a_feature_size)
    assert np.abs(out.detach().numpy() - g['logits']).max() < TOL
    assert np.abs(obj.detach().numpy() - g['obj']).max() < TOL
    assert np.abs(mot.detach().numpy() - g['mot']).max() < TOL
    assert np.abs(alpha.detach().numpy() - g['alpha']).max() < TOL
    loss = O.packed_ce_loss(out, caps, lens)
    assert abs(loss.item() - g['loss'][0]) < TOL
    loss.backward()
    n_checked = 0
    for k, v in sd.items():
        if 'gnone.' + k in g.files:
            assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
            continue
        gn = float(v.grad.double().norm())
        ref = float(g['gnorm.' + k][0])
        assert abs(gn - ref) <= 1e-4 * max(ref, 1e-3), (k, gn, ref)
        assert np.abs(v.grad.reshape(-1)[:8].numpy() - g['ghead.' + k]).max() < 1e-4 * max(1.0, np.abs(g['ghead.' + k]).max()), k
        n_checked += 1
    assert n_checked > 50


@pytest.mark.parametrize('tag,args,V,B', CASES)
def test_capgnn_decode_tokens(golden_dir, tag, args, V, B):
    g = np.load(os.path.join(golden_dir, tag + '.npz'))
    sd = _sd_capgnn(args, V)
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=12)
    with torch.no_grad():
        random.seed(12)
        out6 = O.cap_gnn_forward(sd, frames, regions, caps, args.max_words, 0.6, args.a_feature_size)[0]
        assert np.abs(out6.numpy() - g['logits_tf06']).max() < TOL
        greedy = O.cap_gnn_forward(sd, frames, regions, None, args.max_words, 1.0, args.a_feature_size, beam_size=1)[0]
        assert np.array_equal(greedy.numpy(), g['greedy'])
        obj, mot = O.cap_gnn_encoder(sd, frames, regions, args.a_feature_size)
        for bm in (5, 3):
            best, preds, lp = O.decoder_beam(sd, 'decoder', obj, mot, args.max_words, bm)
            assert np.array_equal(best.numpy(), g['beam%d' % bm])
            assert np.array_equal(preds.numpy(), g['beam%d_all' % bm])
            assert np.abs(lp.numpy() - g['beam%d_lp' % bm]).max() < 1e-4


def _sd_baseline1(args, V):
    full = _sd_capgnn(args, V)
    H, Hd, Hq, W = args.visual_hidden_size, args.decode_hidden_size, args.query_hidden_size, args.word_size
    sd = {}
    for k, v in full.items():
        if k.startswith('encoder.motion_pre_encoder.') and 'self_attention' not in k and 'layernorm_sa' not in k:
            sd[k.replace('encoder.motion_pre_encoder.', 'encoder.')] = None
        elif k.startswith('decoder.') and 'context_att_2' not in k:
            sd[k] = None
    shapes = {k: tuple(full[k.replace('encoder.', 'encoder.motion_pre_encoder.', 1)].shape) if k.startswith('encoder.')
              else tuple(full[k].shape) for k in sd}
    shapes['encoder.out_try.weight'] = (H, 2 * H)
    shapes['encoder.out_try.bias'] = (H,)
    shapes['decoder.query_lstm.weight_ih'] = (4 * Hq, H + W + Hd)
    shapes['decoder.lang_lstm.weight_ih'] = (4 * Hd, H + Hq)
    return {k: synth.fill_tensor(k, s) for k, s in shapes.items()}


def test_baseline1(golden_dir):
    args, V, B = synth.small_args(decode_hidden_size=52), 37, 3
    g = np.load(os.path.join(golden_dir, 'baseline1_small.npz'))
    sd = _sd_baseline1(args, V)
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=13)
    with torch.no_grad():
        out = O.cap_baseline1_forward(sd, frames, caps, args.max_words, 1.0)
        assert np.abs(out.numpy() - g['logits']).max() < TOL
        assert abs(O.packed_ce_loss(out, caps, lens).item() - g['loss'][0]) < TOL
        assert np.array_equal(O.cap_baseline1_forward(sd, frames, None, args.max_words, beam_size=1).numpy(), g['greedy'])
        assert np.array_equal(O.cap_baseline1_forward(sd, frames, None, args.max_words, beam_size=5).numpy(), g['beam5'])


def _sd_disc(V):
    shapes = {'fusion': (2, 512), 'block.0.res_block.1.weight': (512, 512, 3), 'block.0.res_block.1.bias': (512,),
              'conv1d.weight': (512, V, 1), 'conv1d.bias': (512,),
              'lstm.weight_ih_l0': (2048, 512), 'lstm.weight_hh_l0': (2048, 512),
              'lstm.bias_ih_l0': (2048,), 'lstm.bias_hh_l0': (2048,),
              'layer_norm.weight': (512,), 'layer_norm.bias': (512,),
              'att.K.weight': (512, 512), 'att.Q.weight': (512, 512), 'att.V.weight': (512, 512),
              'att.output_layer.0.weight': (512, 512), 'att_norm.1.weight': (512,), 'att_norm.1.bias': (512,),
              'text_sum.theta': (1, 512), 'text_sum.out_norm.1.weight': (512,), 'text_sum.out_norm.1.bias': (512,)}
    for s in ('motion_psl_score', 'obj_psl_score'):
        shapes[s + '.psl_scorer.classify.weight'] = (1, 512)
        shapes[s + '.psl_scorer.classify.bias'] = (1,)
        for e in ('visual_embed.0', 'sent_embed.0'):
            shapes['%s.psl_scorer.%s.weight' % (s, e)] = (512, 512)
            shapes['%s.psl_scorer.%s.bias' % (s, e)] = (512,)
        shapes[s + '.psl_embed.0.weight'] = (512, 1024)
        shapes[s + '.psl_embed.0.bias'] = (512,)
        shapes[s + '.att_norm.0.weight'] = (512, 512)
        shapes[s + '.att_norm.0.bias'] = (512,)
        for n in ('psl_embed.2', 'psl_norm.1', 'att_norm.2'):
            shapes['%s.%s.weight' % (s, n)] = (512,)
            shapes['%s.%s.bias' % (s, n)] = (512,)
    return {k: synth.fill_tensor('D.' + k, s) for k, s in shapes.items()}


@pytest.mark.parametrize('tag,P,K', [('disc_small_msr', 5, 5), ('disc_small_msvd', 8, 3)])
def test_disc_v2_and_gradient_penalty(golden_dir, tag, P, K):
    args = synth.small_args(visual_hidden_size=1024, num_proposals=P, num_topk=K)
    V, B, L = 37, 3, args.max_words
    g = np.load(os.path.join(golden_dir, tag + '.npz'))
    sd = _sd_disc(V)
    for v in sd.values():
        v.requires_grad_(True)
    rs = np.random.RandomState(5)
    _, _, caps, lens = synth.make_inputs(B, args, V, seed=14)
    att_mask = synth.att_mask_from_captions(caps)
    obj = torch.from_numpy(rs.standard_normal((B, P, 1024)).astype(np.float32))
    mot = torch.from_numpy(rs.standard_normal((B, P, 1024)).astype(np.float32))
    alpha = torch.softmax(torch.from_numpy(rs.standard_normal((B, L, 2 * P)).astype(np.float32)), -1)
    fake = torch.from_numpy(rs.standard_normal((B, L, V)).astype(np.float32)).requires_grad_(True)
    real = torch.zeros(B, L, V).scatter_(2, caps.unsqueeze(2), 1)
    eps = torch.from_numpy(rs.uniform(size=(B, 1, 1)).astype(np.float32))
    r = O.disc_v2(sd, real, obj, mot, att_mask, alpha, P, K)
    f = O.disc_v2(sd, fake, obj, mot, att_mask, alpha, P, K)
    mixed = (real * eps + fake.detach() * (1 - eps)).requires_grad_(True)
    m = O.disc_v2(sd, mixed, obj, mot, att_mask, alpha, P, K)
    assert np.abs(r.detach().numpy() - g['r_logit']).max() < TOL
    assert np.abs(f.detach().numpy() - g['f_logit']).max() < TOL
    assert np.abs(m.detach().numpy() - g['m_logit']).max() < TOL
    gr = torch.autograd.grad(m, mixed, torch.ones_like(m), create_graph=True, retain_graph=True)[0]
    gn = gr.reshape(B, -1).norm(2, dim=1)
    gp = ((gn - 1) ** 2).mean()
    assert np.abs(gn.detach().numpy() - g['gnorm_mixed']).max() < 1e-4
    loss = f.mean() - r.mean() + 10 * gp
    assert abs(loss.item() - g['loss_d'][0]) < 1e-4
    loss.backward()
    assert np.abs(fake.grad.numpy() - g['dfake']).max() < 1e-5
    for k, v in sd.items():
        ref = float(g['gnorm.' + k][0])
        gn_k = float(v.grad.double().norm())
        assert abs(gn_k - ref) <= 2e-4 * max(ref, 1e-3), (k, gn_k, ref)
