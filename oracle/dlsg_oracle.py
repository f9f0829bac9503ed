"""CPU fp32 oracle for the D-LSG hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this file; the product path (d-lsg-video-caption_b200/) never does.

It is a functional restatement (plain torch CPU ops over a state_dict-keyed dict of
tensors, no nn.Module, no autograd.Function, eval-mode = dropout off) of the reference's
models/model.py, models/layer.py, models/sublayer.py and models/allennlp_beamsearch.py.
Each function cites the reference file:line it follows.  Parity is PINNED against the
reference itself: tests/golden/make_golden.py imports the unmodified reference modules in
the build container, runs them on dlsg.synth inputs/weights and commits the outputs as
tests/golden/*.npz; tests/test_oracle_golden.py checks this file against those vectors.
(The reference ships no tests / golden vectors of its own - SURVEY.md 8c.)

Gradients: run these functions under torch autograd (they are plain differentiable ops).
"""
import math
import random
import warnings

import numpy as np
import torch
import torch.nn.functional as F

PAD, START, END = 0, 1, 2


# ----------------------------------------------------------------------------- primitives
def layer_norm(x, w, b, eps=1e-5):
    """nn.LayerNorm: biased variance over the last dim, eps=1e-5."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def ln(sd, key, x):
    return layer_norm(x, sd[key + '.weight'], sd[key + '.bias'])


def linear(sd, key, x, bias=True):
    y = x @ sd[key + '.weight'].t()
    if bias:
        y = y + sd[key + '.bias']
    return y


def scaled(x, d):
    """torch.div(x, torch.tensor(np.sqrt(d))): float64 0-dim divisor, fp32 result."""
    return torch.div(x, torch.tensor(np.sqrt(d)))


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """nn.LSTMCell / nn.LSTM step: gate order i,f,g,o."""
    g = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    H = h.shape[-1]
    i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def lstm_seq(sd, pfx, x, reverse=False, suffix=''):
    """One direction of nn.LSTM(batch_first) with zero initial state (layer.py:40-44,52)."""
    B, T, _ = x.shape
    w_ih, w_hh = sd[pfx + '.weight_ih_l0' + suffix], sd[pfx + '.weight_hh_l0' + suffix]
    b_ih, b_hh = sd[pfx + '.bias_ih_l0' + suffix], sd[pfx + '.bias_hh_l0' + suffix]
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    out = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        h, c = lstm_cell(x[:, t], h, c, w_ih, w_hh, b_ih, b_hh)
        out[t] = h
    return torch.stack(out, 1)


def positional_encoding(d_model, n):
    """sublayer.py:85-98 (buffer 'pe'); recomputed, not read from the state_dict."""
    pe = torch.zeros(n, d_model)
    pos = torch.arange(0., n).unsqueeze(1)
    div = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


# ----------------------------------------------------------------------------- encoder
def latent_psl(sd, pfx, x):
    """sublayer.py:189-198: softmax over the sequence axis (dim=1), pool, tanh->LN."""
    g = F.softmax(x @ sd[pfx + '.theta'].t(), dim=1)
    n = g.transpose(-1, -2) @ x
    return ln(sd, pfx + '.out_norm.1', torch.tanh(n))


def encoder_tun(sd, pfx, visual, regions, use_embed=True, baseline=False):
    """layer.py:172-201 EncoderVisualGraphTUN.forward."""
    bs, T, R, Dr = regions.shape
    v = linear(sd, pfx + '.visual_embed', visual) if use_embed else visual
    fr = ln(sd, pfx + '.visual_norm.1', torch.tanh(v))
    if R < 5:
        x = fr
    else:
        o = linear(sd, pfx + '.obj_embed', regions).view(bs, T * R, -1)
        o = ln(sd, pfx + '.obj_norm.1', torch.tanh(o))
        s = F.softmax(scaled(o @ fr.transpose(-1, -2), Dr), dim=1)      # over the T*R objects
        agg = (o.transpose(-1, -2) @ s).transpose(-1, -2)
        x = ln(sd, pfx + '.obj_visual_norm.1', torch.tanh(agg + fr))
    if baseline:
        return x
    return latent_psl(sd, pfx + '.v2l_layer', x)


def self_attention(sd, pfx, x, att_mask=None, get_pe=False):
    """sublayer.py:63-82: row i uses K_i as the query side; mask fill -9e15."""
    d = sd[pfx + '.K.weight'].shape[0]
    if get_pe:
        x = x + positional_encoding(d, x.shape[1]).to(x.dtype)
    k = x @ sd[pfx + '.K.weight'].t()
    q = x @ sd[pfx + '.Q.weight'].t()
    v = x @ sd[pfx + '.V.weight'].t()
    logits = scaled(k @ q.transpose(-1, -2), d)
    if att_mask is not None:
        logits = torch.where(att_mask > 0, logits, -9e15 * torch.ones_like(logits))
    w = F.softmax(logits, dim=-1)
    return (w @ v) @ sd[pfx + '.output_layer.0.weight'].t()


def encoder_visual(sd, pfx, frames, baseline=False):
    """layer.py:46-61 EncoderVisual.forward (embed=True)."""
    x = linear(sd, pfx + '.linear_embed', frames)
    out = torch.cat([lstm_seq(sd, pfx + '.lstm', x), lstm_seq(sd, pfx + '.lstm', x, True, '_reverse')], -1)
    out = ln(sd, pfx + '.layernorm_lstm', out)
    if baseline:
        return linear(sd, pfx + '.out_try', out)
    out = self_attention(sd, pfx + '.self_attention', out, None, get_pe=True)
    return ln(sd, pfx + '.layernorm_sa', out)


def cap_gnn_encoder(sd, frames, regions, a_feature_size, pfx='encoder'):
    """model.py:69-73."""
    obj = encoder_tun(sd, pfx + '.obj_encoder', frames[:, :, :a_feature_size], regions, True)
    mot_in = encoder_visual(sd, pfx + '.motion_pre_encoder', frames)
    mot = encoder_tun(sd, pfx + '.motion_encoder', mot_in, regions, False)
    return obj, mot


# ----------------------------------------------------------------------------- decoder
def attention_share(sd, pfx, nodes, q):
    """sublayer.py:28-43: softmax over the node axis (dim=1); tanh in the output layer."""
    d = sd[pfx + '.K.weight'].shape[0]
    k = nodes @ sd[pfx + '.K.weight'].t()
    v = (nodes @ sd[pfx + '.V.weight'].t()).transpose(-1, -2)
    qq = (q @ sd[pfx + '.Q.weight'].t()).unsqueeze(2)
    w = F.softmax(scaled(k @ qq, d), dim=1)
    ctx = (v @ w).squeeze(2)
    out = ln(sd, pfx + '.output_layer.2', torch.tanh(ctx @ sd[pfx + '.output_layer.0.weight'].t()))
    return out, w


def decode_step(sd, pfx, word, qh, qc, lh, lc, glob, n1, n2):
    """layer.py:569-602 Decoder.decode (multi_modal iff n2 is not None, use_fusion=False)."""
    qh, qc = lstm_cell(torch.cat([lh, glob, word], 1), qh, qc,
                       sd[pfx + '.query_lstm.weight_ih'], sd[pfx + '.query_lstm.weight_hh'],
                       sd[pfx + '.query_lstm.bias_ih'], sd[pfx + '.query_lstm.bias_hh'])
    q = ln(sd, pfx + '.query_lstm_layernorm', qh)
    ctx, alpha = attention_share(sd, pfx + '.context_att', n1, q)
    if n2 is not None:
        ctx2, alpha2 = attention_share(sd, pfx + '.context_att_2', n2, q)
        lang_in = torch.cat([ctx, ctx2, q], 1)
        alpha = torch.cat([alpha, alpha2], 1)
    else:
        lang_in = torch.cat([ctx, q], 1)
    lh, lc = lstm_cell(lang_in, lh, lc,
                       sd[pfx + '.lang_lstm.weight_ih'], sd[pfx + '.lang_lstm.weight_hh'],
                       sd[pfx + '.lang_lstm.bias_ih'], sd[pfx + '.lang_lstm.bias_hh'])
    dec = torch.tanh(ln(sd, pfx + '.lang_lstm_layernorm', lh))
    logits = linear(sd, pfx + '.word_restore', dec)
    return logits, qh, qc, lh, lc, alpha


def _decoder_init(sd, pfx, n1, n2, multi_modal):
    """layer.py:407-422: global feature, zero states, <start> embedding."""
    B = n1.shape[0]
    glob = n1.mean(1)
    if n2 is not None:
        glob = torch.cat([glob, n2.mean(1)], -1)
        if not multi_modal:
            n1 = torch.cat([n1, n2], 1)
            n2 = None
    Hq = sd[pfx + '.query_lstm.weight_hh'].shape[1]
    Hd = sd[pfx + '.lang_lstm.weight_hh'].shape[1]
    z = lambda h: n1.new_zeros(B, h)
    start = torch.full((B,), START, dtype=torch.long)
    return n1, n2, glob, z(Hq), z(Hq), z(Hd), z(Hd), start


def decoder_forward(sd, pfx, n1, n2, captions, max_words, tf_ratio=1.0, multi_modal=True, rng=random):
    """layer.py:394-447 train / greedy branch.  captions None -> greedy ids (B,max_words)."""
    infer = captions is None
    n1, n2, glob, qh, qc, lh, lc, wid = _decoder_init(sd, pfx, n1, n2, multi_modal)
    emb = sd[pfx + '.word_embed.weight']
    word = emb[wid]
    outs, alphas = [], []
    for i in range(max_words):
        logits, qh, qc, lh, lc, alpha = decode_step(sd, pfx, word, qh, qc, lh, lc, glob, n1, n2)
        use_tf = (not infer) and (rng.random() < tf_ratio)          # layer.py:432 short-circuit
        wid = captions[:, i] if use_tf else logits.max(1)[1]
        word = emb[wid]
        if infer:
            outs.append(wid)
        else:
            outs.append(logits)
            alphas.append(alpha)
    return torch.stack(outs, 1), alphas


def beam_search(step, start_pred, state, end_index, max_steps, beam, per_node):
    """allennlp_beamsearch.py:51-294 restated; `step(last_pred, state)->(logp, state)`."""
    B = start_pred.shape[0]
    logp0, state = step(start_pred, state)
    V = logp0.shape[1]
    if per_node > V:
        raise ValueError('Target vocab size (%d) too small relative to per_node_beam_size (%d).' % (V, per_node))
    last_lp, pred0 = logp0.topk(beam)
    if beam == 1 and bool((pred0 == end_index).all()):
        warnings.warn('Empty sequences predicted.', RuntimeWarning)
        return pred0.unsqueeze(-1), last_lp
    preds, backs = [pred0], []
    after_end = logp0.new_full((B * beam, V), float('-inf'))
    after_end[:, end_index] = 0.0
    state = {k: v.unsqueeze(1).expand(B, beam, *v.shape[1:]).reshape(B * beam, *v.shape[1:])
             for k, v in state.items()}
    for _ in range(max_steps - 1):
        last = preds[-1].reshape(B * beam)
        if bool((last == end_index).all()):
            break
        logp, state = step(last, state)
        cleaned = torch.where(last.unsqueeze(-1).expand(B * beam, V) == end_index, after_end, logp)
        top_lp, top_cls = cleaned.topk(per_node)
        summed = (top_lp + last_lp.unsqueeze(2).expand(B, beam, per_node).reshape(B * beam, per_node))
        summed = summed.reshape(B, beam * per_node)
        cls = top_cls.reshape(B, beam * per_node)
        last_lp, idx = summed.topk(beam)
        preds.append(cls.gather(1, idx))
        bp = (idx / per_node).type(torch.int64)
        backs.append(bp)
        new_state = {}
        for k, v in state.items():
            e = bp.view(B, beam, *([1] * (v.dim() - 1))).expand(B, beam, *v.shape[1:])
            new_state[k] = v.reshape(B, beam, *v.shape[1:]).gather(1, e).reshape(B * beam, *v.shape[1:])
        state = new_state
    rec = [preds[-1].unsqueeze(2)]
    cur = backs[-1]
    for t in range(len(preds) - 2, 0, -1):
        rec.append(preds[t].gather(1, cur).unsqueeze(2))
        cur = backs[t - 1].gather(1, cur)
    rec.append(preds[0].gather(1, cur).unsqueeze(2))
    return torch.cat(list(reversed(rec)), 2), last_lp


def decoder_beam(sd, pfx, n1, n2, max_words, beam, multi_modal=True):
    """layer.py:449-460 + beam_step 489-567 (per-beam decode == batched decode row-wise)."""
    n1, n2, glob, qh, qc, lh, lc, start = _decoder_init(sd, pfx, n1, n2, multi_modal)
    emb = sd[pfx + '.word_embed.weight']
    B = n1.shape[0]
    st = {'qh': qh, 'qc': qc, 'lh': lh, 'lc': lc, 'n1': n1, 'glob': glob}
    if n2 is not None:
        st['n2'] = n2

    def step(last, s):
        logits, qh_, qc_, lh_, lc_, _ = decode_step(sd, pfx, emb[last], s['qh'], s['qc'], s['lh'], s['lc'],
                                                    s['glob'], s['n1'], s.get('n2'))
        s2 = dict(s)
        s2.update(qh=qh_, qc=qc_, lh=lh_, lc=lc_)
        return F.log_softmax(logits, dim=1), s2

    preds, lp = beam_search(step, start, st, END, max_words, beam, beam)
    best = torch.topk(lp, 1)[1].squeeze(1)
    return torch.stack([preds[i, best[i], :] for i in range(B)]), preds, lp


def cap_gnn_forward(sd, frames, regions, captions, max_words, tf_ratio=1.0, a_feature_size=1536,
                    beam_size=5, rng=random):
    """model.py:32-40 CapGnnModel.forward.  Returns (outputs, obj, motion, alpha_all)."""
    obj, mot = cap_gnn_encoder(sd, frames, regions, a_feature_size)
    if captions is None and beam_size > 1:
        out, _, _ = decoder_beam(sd, 'decoder', obj, mot, max_words, beam_size)
        return out, obj, mot, []
    out, alphas = decoder_forward(sd, 'decoder', obj, mot, captions, max_words, tf_ratio, True, rng)
    alpha_all = torch.cat(alphas, dim=-1).transpose(1, 2) if len(alphas) else []
    return out, obj, mot, alpha_all


def cap_baseline1_forward(sd, frames, captions, max_words, tf_ratio=1.0, beam_size=5, rng=random):
    """model.py:101-104 CapBaseline1: EncoderVisual(baseline) + Decoder(multi_modal=False)."""
    enc = encoder_visual(sd, 'encoder', frames, baseline=True)
    if captions is None and beam_size > 1:
        out, _, _ = decoder_beam(sd, 'decoder', enc, None, max_words, beam_size, multi_modal=False)
        return out
    out, _ = decoder_forward(sd, 'decoder', enc, None, captions, max_words, tf_ratio, False, rng)
    return out


def packed_ce_loss(outputs, captions, cap_lens):
    """run_gun.py:189-197: pack by cap_lens, nn.CrossEntropyLoss (mean over tokens)."""
    o = torch.cat([outputs[j][:cap_lens[j]] for j in range(len(cap_lens))], 0)
    t = torch.cat([captions[j][:cap_lens[j]] for j in range(len(cap_lens))], 0)
    return F.cross_entropy(o, t)


# ----------------------------------------------------------------------------- discriminator
def psl_score2(sd, pfx, psl, alpha, att_out, seq_mask, num_top):
    """layer.py:688-715 PSLScore2.forward -> 0-dim scalar (batch mean)."""
    bs, P, _ = psl.shape
    p = ln(sd, pfx + '.psl_embed.2', torch.tanh(linear(sd, pfx + '.psl_embed.0', psl)))
    if P > num_top:
        idx = torch.topk(alpha.sum(dim=1), num_top, -1)[1]
        p = torch.gather(p, 1, idx.unsqueeze(-1).expand(bs, num_top, p.shape[-1]))
    a = ln(sd, pfx + '.att_norm.2', torch.tanh(linear(sd, pfx + '.att_norm.0', att_out)))
    adj = F.softmax(scaled(a @ p.transpose(-1, -2), 512), dim=1)        # over the words
    adj = torch.where(seq_mask > 0, adj, torch.zeros_like(adj))
    adj_alpha = adj.sum(1)
    g = (a.transpose(-1, -2) @ adj).transpose(-1, -2)
    g = ln(sd, pfx + '.psl_norm.1', torch.tanh(g))
    sc = pfx + '.psl_scorer'
    s = linear(sd, sc + '.classify',
               torch.tanh(linear(sd, sc + '.visual_embed.0', p)) * torch.tanh(linear(sd, sc + '.sent_embed.0', g)))
    s = s.squeeze()
    s = (s * adj_alpha).sum(-1) / adj_alpha.sum(-1)
    return s.mean(-1)


def disc_v2(sd, inputs, obj, mot, att_mask, alpha_all, num_psl, num_top):
    """model.py:145-168 DiscV2.forward -> (B,)."""
    x = inputs @ sd['conv1d.weight'][:, :, 0].t() + sd['conv1d.bias']          # conv1d k=1, (B,L,512)
    r = torch.relu(x)                                                          # in-place ReLU quirk
    w3 = sd['block.0.res_block.1.weight']
    rp = F.pad(r, (0, 0, 1, 1))
    conv = (rp[:, :-2] @ w3[:, :, 0].t() + rp[:, 1:-1] @ w3[:, :, 1].t() + rp[:, 2:] @ w3[:, :, 2].t()
            + sd['block.0.res_block.1.bias'])
    y = r + 0.3 * conv
    h = ln(sd, 'layer_norm', lstm_seq(sd, 'lstm', y))
    att = ln(sd, 'att_norm.1', torch.tanh(self_attention(sd, 'att', h, att_mask)))
    seq = att_mask[:, 0, :].unsqueeze(2)
    alpha_all = alpha_all * seq
    m = seq.repeat(1, 1, num_top)
    s_obj = psl_score2(sd, 'obj_psl_score', obj, alpha_all[:, :, :num_psl], att, m, num_top)
    s_mot = psl_score2(sd, 'motion_psl_score', mot, alpha_all[:, :, -num_psl:], att, m, num_top)
    sent = latent_psl(sd, 'text_sum', att).squeeze()
    f = F.softmax(sent @ sd['fusion'].t(), dim=-1)
    return s_obj * f[:, 0] + s_mot * f[:, 1]
