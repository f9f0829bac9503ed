"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Run once in the build container:  python tests/golden/make_golden.py
The reference is imported with a stub for its one missing third-party symbol
(allennlp.common.checks.ConfigurationError, models/allennlp_beamsearch.py:12).
Weights/inputs come from dlsg.synth (closed-form of name+shape), so only OUTPUTS are stored.
/root/reference does not exist on the GPU box: tests only read the committed .npz files.
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('DLSG_REFERENCE', '/root/reference')

for name in ('allennlp', 'allennlp.common', 'allennlp.common.checks'):
    sys.modules[name] = types.ModuleType(name)


class ConfigurationError(Exception):
    pass


sys.modules['allennlp.common.checks'].ConfigurationError = ConfigurationError
sys.path.insert(0, REF)
sys.path.insert(1, os.path.join(ROOT, 'd-lsg-video-caption_b200'))

import io
import contextlib

with contextlib.redirect_stdout(io.StringIO()):
    from models.model import CapGnnModel, CapBaseline1, DiscV2  # noqa: E402  (the reference's)
import models as _m  # noqa: E402

assert _m.__file__.startswith(REF), _m.__file__
from dlsg import synth  # noqa: E402


def grad_summary(module):
    out = {}
    for k, p in module.named_parameters():
        if p.grad is None:
            out['gnone.' + k] = np.zeros(0, np.float32)
        else:
            g = p.grad.detach().reshape(-1)
            out['gnorm.' + k] = np.array([float(g.double().norm())], np.float64)
            out['ghead.' + k] = g[:8].numpy().copy()
    return out


def packed_ce(outputs, captions, cap_lens):
    o = torch.cat([outputs[j][:cap_lens[j]] for j in range(len(cap_lens))], 0)
    t = torch.cat([captions[j][:cap_lens[j]] for j in range(len(cap_lens))], 0)
    return torch.nn.CrossEntropyLoss()(o, t)


def gen_capgnn(tag, args, V, B):
    vocab = synth.Vocab(V)
    with contextlib.redirect_stdout(io.StringIO()):
        net = CapGnnModel(args, vocab)
    synth.fill_state_dict(net)
    net.eval()
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=12)
    res = {}
    net.zero_grad()
    out, obj, mot, alpha = net(frames, regions, caps, args.max_words, 1.0)
    loss = packed_ce(out, caps, lens)
    loss.backward()
    res.update(logits=out.detach().numpy(), obj=obj.detach().numpy(), mot=mot.detach().numpy(),
               alpha=alpha.detach().numpy(), loss=np.array([loss.item()], np.float64))
    res.update(grad_summary(net))
    with torch.no_grad():
        random.seed(12)
        out6, _, _, _ = net(frames, regions, caps, args.max_words, 0.6)
        res['logits_tf06'] = out6.numpy()
        net.update_beam_size(1)
        res['greedy'] = net(frames, regions, None)[0].numpy()
        for bm in (5, 3):
            net.update_beam_size(bm)
            res['beam%d' % bm] = net(frames, regions, None)[0].numpy()
            # all beams + scores straight from BeamSearch.search
            dec = net.decoder
            o, m = net.encoder(frames, regions)
            dec.batch_size = B
            glob = torch.cat([o.mean(1), m.mean(1)], -1)
            z = lambda h: o.new_zeros(B, h)
            st = {'query_lstm_h': z(args.query_hidden_size), 'query_lstm_c': z(args.query_hidden_size),
                  'lang_lstm_h': z(args.decode_hidden_size), 'lang_lstm_c': z(args.decode_hidden_size),
                  'cnn_feats': o, 'global_feat': glob, 'cnn_feats_2': m}
            start = torch.full((B,), 1, dtype=torch.long)
            preds, lp = dec.beam_search.search(start, st, dec.beam_step)
            res['beam%d_all' % bm] = preds.numpy()
            res['beam%d_lp' % bm] = lp.numpy()
    np.savez_compressed(os.path.join(HERE, tag + '.npz'), **res)
    print(tag, 'loss', loss.item(), 'greedy', res['greedy'][0][:8], 'beam5', res['beam5'].shape)


def gen_baseline1(tag, args, V, B):
    vocab = synth.Vocab(V)
    with contextlib.redirect_stdout(io.StringIO()):
        net = CapBaseline1(args, vocab)
    synth.fill_state_dict(net)
    net.eval()
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=13)
    net.zero_grad()
    out = net(frames, regions, caps, args.max_words, 1.0)[0]
    loss = packed_ce(out, caps, lens)
    loss.backward()
    res = dict(logits=out.detach().numpy(), loss=np.array([loss.item()], np.float64))
    res.update(grad_summary(net))
    with torch.no_grad():
        net.update_beam_size(1)
        res['greedy'] = net(frames, regions, None)[0].numpy()
        net.update_beam_size(5)
        res['beam5'] = net(frames, regions, None)[0].numpy()
    np.savez_compressed(os.path.join(HERE, tag + '.npz'), **res)
    print(tag, 'loss', loss.item())


def gen_disc(tag, args, V, B):
    """DiscV2 on (one-hot | soft logits | mixed) inputs + WGAN-GP penalty (run_gun.py:351-375)."""
    net = DiscV2(args, V)
    synth.fill_state_dict(net, prefix='D.')
    net.eval()
    L, P = args.max_words, args.num_proposals
    rs = np.random.RandomState(5)
    _, _, caps, lens = synth.make_inputs(B, args, V, seed=14)
    att_mask = synth.att_mask_from_captions(caps)
    obj = torch.from_numpy(rs.standard_normal((B, P, 1024)).astype(np.float32))
    mot = torch.from_numpy(rs.standard_normal((B, P, 1024)).astype(np.float32))
    alpha = torch.softmax(torch.from_numpy(rs.standard_normal((B, L, 2 * P)).astype(np.float32)), -1)
    fake = torch.from_numpy(rs.standard_normal((B, L, V)).astype(np.float32))
    real = torch.zeros(B, L, V).scatter_(2, caps.unsqueeze(2), 1)
    eps = torch.from_numpy(rs.uniform(size=(B, 1, 1)).astype(np.float32))
    res = {}
    net.zero_grad()
    r_logit = net(real.clone(), obj, mot, att_mask, alpha)
    f_in = fake.clone().requires_grad_(True)
    f_logit = net(f_in, obj, mot, att_mask, alpha)
    mixed = (real * eps + fake * (1 - eps)).requires_grad_(True)
    m_logit = net(mixed, obj, mot, att_mask, alpha)
    g = torch.autograd.grad(m_logit, mixed, torch.ones_like(m_logit), create_graph=True, retain_graph=True)[0]
    gn = g.contiguous().view(B, -1).norm(2, dim=1)
    gp = ((gn - 1) * (gn - 1)).mean()
    loss_d = f_logit.mean() - r_logit.mean() + 10 * gp
    loss_d.backward()
    res.update(r_logit=r_logit.detach().numpy(), f_logit=f_logit.detach().numpy(),
               m_logit=m_logit.detach().numpy(), gp=np.array([gp.item()], np.float64),
               gnorm_mixed=gn.detach().numpy(), loss_d=np.array([loss_d.item()], np.float64),
               dfake=f_in.grad.numpy())
    res.update(grad_summary(net))
    np.savez_compressed(os.path.join(HERE, tag + '.npz'), **res)
    print(tag, 'r', r_logit.detach().numpy(), 'gp', gp.item())


def gen_init_order():
    """Registration order, shapes and seeded default initialisation of every parameter / buffer of the reference's live
    models at the real (MSR-VTT / MSVD) widths: the mirror must reproduce all three (state_dict AND optimizer checkpoints
    index parameters by position; a freshly initialised model must train like the reference's)."""
    import json
    out = {}
    for name, args in (('msr', synth.msr_args()), ('msvd', synth.msvd_args())):
        for cls_name, make in (('CapGnnModel', lambda: CapGnnModel(args, synth.Vocab(101))),
                               ('CapBaseline1', lambda: CapBaseline1(args, synth.Vocab(101))),
                               ('DiscV2', lambda: DiscV2(args, 101))):
            torch.manual_seed(3)
            with contextlib.redirect_stdout(io.StringIO()):
                net = make()
            out['%s.%s' % (name, cls_name)] = [[k, list(v.shape), float(v.double().abs().sum())] for k, v in net.state_dict().items()]
    with open(os.path.join(HERE, 'init_order.json'), 'w') as f:
        json.dump(out, f)
    print('init_order.json:', {k: len(v) for k, v in out.items()})


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'init_order':
        gen_init_order()
        sys.exit(0)
    torch.manual_seed(12)
    torch.set_num_threads(8)
    gen_capgnn('capgnn_small_msr', synth.small_args(), V=37, B=3)
    gen_capgnn('capgnn_small_msvd', synth.small_args(num_proposals=8, num_topk=3, decode_hidden_size=64,
                                                     dataset='msvd', num_obj=5), V=41, B=2)
    gen_baseline1('baseline1_small', synth.small_args(decode_hidden_size=52), V=37, B=3)
    gen_disc('disc_small_msr', synth.small_args(visual_hidden_size=1024, num_proposals=5, num_topk=5), V=37, B=3)
    gen_disc('disc_small_msvd', synth.small_args(visual_hidden_size=1024, num_proposals=8, num_topk=3), V=37, B=3)
    gen_init_order()
