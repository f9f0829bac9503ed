"""Drop-in mirror of the reference's models/model.py (CapGnnModel :25-53, CapGnnEncoder :56-73, DiscV2 :110-168,
CapBaseline1 :94-107, CapBaselineModel :76-91, CapModel :10-22): same constructor arguments, forward signatures, return
tuples and state_dict keys, so run_gun.py / run_graph.py / evaluate.py drive it unchanged.  Arithmetic: libdlsg sm_100a
kernels (dlsg.functional / dlsg.decoder / dlsg.generic); the classes below only wire blocks together.
"""
import torch
import torch.nn as nn

from models.layer import (EncoderVisual, EncoderVisualGraph, Decoder, EncoderVisualGAT, EncoderVisualGraphTUN,   # noqa: F401
                          PSLScore, PSLScore2, tun_pair_forward)
from models.sublayer import SelfAttention, JointEmbedVideoModel2, AttentionShare, ResBlock, LatentGNN, LatentPSL   # noqa: F401
from dlsg import generic as G
from dlsg.modspec import declare, lin, tanh_norm, xavier_uniform_param


class _Captioner(nn.Module):
    """What every captioning model of the reference shares: `update_beam_size` forwards to the decoder (model.py:21,42,90,106)."""

    def update_beam_size(self, beam_size):
        self.decoder.update_beam_size(beam_size)


class CapModel(_Captioner):
    """Stale in the reference (run.py is broken): Decoder(args, vocab) expects a 2H global feature that
    this wiring never provides.  Constructible for import compatibility."""

    def __init__(self, args, vocab):
        super().__init__()
        declare(self, [('encoder', lambda: EncoderVisual(args)), ('decoder', lambda: Decoder(args, vocab))])

    def forward(self, visual_feats, caption, max_words=None, teacher_forcing_ratio=1.0):
        return self.decoder(self.encoder(visual_feats), caption, max_words, teacher_forcing_ratio)[0]


class CapGnnModel(_Captioner):
    def __init__(self, args, vocab):
        super().__init__()
        self.use_visual_gan = args.use_visual_gan
        declare(self, [('encoder', lambda: CapGnnEncoder(args)), ('decoder', lambda: Decoder(args, vocab, multi_modal=True))])

    def forward(self, visual_feats, region_feats, caption, max_words=None, teacher_forcing_ratio=1.0):
        """-> (logits (B,T,V) | token ids, object nodes (B,P,H), motion nodes (B,P,H), attention weights (B,T,2P) | [])."""
        obj_nodes, motion_nodes = self.encoder(visual_feats, region_feats)
        outputs, alpha = self.decoder._run(obj_nodes, caption, max_words, teacher_forcing_ratio, motion_nodes)
        return outputs, obj_nodes, motion_nodes, ([] if alpha is None else alpha)

    def load_encoder(self, model, model_path):
        """Take encoder + word embedding of a pre-trained model and freeze the embedding (model.py:45-53)."""
        model.load_state_dict(torch.load(model_path, map_location='cuda:0'))
        self.encoder, self.decoder.word_embed = model.encoder, model.decoder.word_embed
        self.decoder.word_embed.weight.requires_grad_(False)


class CapGnnEncoder(nn.Module):
    def __init__(self, args, baseline=False):
        super().__init__()
        self.a_feature_size = args.a_feature_size
        declare(self, [('obj_encoder', lambda: EncoderVisualGraphTUN(args, input_type='object', baseline=baseline)),
                       ('motion_pre_encoder', lambda: EncoderVisual(args)),
                       ('motion_encoder', lambda: EncoderVisualGraphTUN(args, input_type='motion', use_embed=False, baseline=baseline))])

    def forward(self, visual_feats, region_feats):
        """Object path: the first a_feature_size channels of the frame features; motion path: the BiLSTM/self-attention
        encoding of all channels.  Both graph encoders run as one block (one region-projection GEMM)."""
        appearance = visual_feats[:, :, :self.a_feature_size]
        return tun_pair_forward(self.obj_encoder, appearance, self.motion_encoder, self.motion_pre_encoder(visual_feats), region_feats)


class CapBaselineModel(_Captioner):
    def __init__(self, args, vocab):
        super().__init__()
        self.use_visual_gan = args.use_visual_gan
        declare(self, [('encoder', lambda: CapGnnEncoder(args, baseline=True)),
                       ('linear_baseline', lambda: lin(args.visual_hidden_size * 2, args.visual_hidden_size)),
                       ('decoder', lambda: Decoder(args, vocab, multi_modal=False, baseline=True))])

    def forward(self, visual_feats, region_feats, caption, max_words=None, teacher_forcing_ratio=1.0):
        _, motion_nodes = self.encoder(visual_feats, region_feats)
        return self.decoder(motion_nodes, caption, max_words, teacher_forcing_ratio)[0], 0, 0, 0


class CapBaseline1(_Captioner):
    def __init__(self, args, vocab):
        super().__init__()
        self.use_visual_gan = args.use_visual_gan
        declare(self, [('encoder', lambda: EncoderVisual(args, baseline=True)),
                       ('decoder', lambda: Decoder(args, vocab, multi_modal=False, baseline=True))])

    def forward(self, visual_feats, region_feats, caption, max_words=None, teacher_forcing_ratio=1.0):
        return self.decoder._run(self.encoder(visual_feats), caption, max_words, teacher_forcing_ratio)[0], 0, 0, 0


class DiscV2(nn.Module):
    def __init__(self, opt, vocab_size):
        super().__init__()
        d = self.dim = 512
        self.num_top, self.seq_len, self.num_psl = opt.num_topk, opt.max_words, opt.num_proposals
        declare(self, [('block', lambda: nn.Sequential(ResBlock(d))),
                       ('conv1d', lambda: nn.Conv1d(vocab_size, d, 1)),
                       ('lstm', lambda: nn.LSTM(d, d, batch_first=True, bidirectional=False)),
                       ('layer_norm', lambda: nn.LayerNorm(d)),
                       ('lstm_drop', lambda: nn.Dropout(0.3)),
                       ('att', lambda: SelfAttention(d, d, d, 0.3)),
                       ('att_norm', lambda: tanh_norm(d)),
                       ('motion_psl_score', lambda: PSLScore2(opt.num_proposals, self.num_top)),
                       ('obj_psl_score', lambda: PSLScore2(opt.num_proposals, self.num_top)),
                       ('text_sum', lambda: LatentPSL(d, 1)),
                       ('fusion', lambda: xavier_uniform_param(2, d, 'tanh'))])

    @staticmethod
    def get_discriminator_block(input_dim, output_dim):
        """Unused helper of the reference (model.py:138-143), kept for API compatibility."""
        return nn.Sequential(lin(input_dim, output_dim), nn.LeakyReLU(0.2))

    @G.param_scope
    def forward(self, inputs, obj_proposals, motion_proposals, att_mask=None, alpha_all=None, _groups=1, _tokens=None):
        """inputs (B,L,V) one-hot / logits / mix -> score (B,)   (model.py:145-168).
        _groups > 1 (our own extension, used by dlsg.gan): the batch is `_groups` independent reference calls stacked along
        dim 0 - D(real), D(fake), D(mixed) of a critic step - so the per-call batch means of PSLScore2 (layer.py:713-714)
        are taken per group and the result equals the concatenation of the separate calls.
        _tokens (our own extension, dlsg.gan): the (B,L,512) output of the input projection computed by the caller (a column
        gather for one-hot captions, a linear mix for the WGAN-GP interpolate); `inputs` is ignored."""
        drop = 0.3 if self.training else 0.0
        res_conv = self.block[0].res_block[1]
        rnn, ln = self.lstm, self.layer_norm
        if _tokens is not None:
            tokens = _tokens
        else:
            tokens = G.linear(inputs, self.conv1d.weight[:, :, 0], self.conv1d.bias)  # conv1d with k=1 == per-token Linear (B,L,512)
        tokens = G.resblock_blc(tokens, res_conv.weight, res_conv.bias)               # relu(x) + 0.3 * conv3(relu(x))
        states = G.lstm(tokens, rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0)
        states = G.norm(states, ln.weight, ln.bias, p_drop=drop)
        words = G.norm(self.att(states, att_mask), self.att_norm[1].weight, self.att_norm[1].bias, pre_tanh=True)   # (B,L,512)
        valid = att_mask[:, 0, :].unsqueeze(2)                                        # (B,L,1): 1 for real words
        alpha = G.mul(alpha_all, valid.expand_as(alpha_all).contiguous())
        word_mask = valid.repeat(1, 1, self.num_top)
        P = self.num_psl
        s_obj = self.obj_psl_score(obj_proposals, alpha[:, :, :P], words, word_mask, _groups)
        s_mot = self.motion_psl_score(motion_proposals, alpha[:, :, -P:], words, word_mask, _groups)
        summary = self.text_sum(words).squeeze()                                      # (B,512) sentence summary
        mix = G.softmax(G.linear(summary, self.fusion), dim=-1)                       # (B,2) object / motion weights
        return G.fuse_scores(s_obj, s_mot, mix, _groups)
