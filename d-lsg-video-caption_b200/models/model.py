"""Drop-in mirror of the reference's models/model.py (CapGnnModel :25-53, CapGnnEncoder :56-73,
DiscV2 :110-168, CapBaseline1 :94-107, CapBaselineModel :76-91, CapModel :10-22): same ctor arguments,
forward signatures, return tuples and state_dict keys, so run_gun.py / run_graph.py / evaluate.py drive it
unchanged.  Arithmetic: libdlsg sm_100a kernels (dlsg.functional / dlsg.decoder / dlsg.generic).
"""
from models.layer import (EncoderVisual, EncoderVisualGraph, Decoder, EncoderVisualGAT, EncoderVisualGraphTUN,
                          PSLScore, PSLScore2, tun_pair_forward)
from models.sublayer import SelfAttention, JointEmbedVideoModel2, AttentionShare, ResBlock, LatentGNN, LatentPSL
import torch.nn as nn
import torch
import torch.nn.functional as F
import numpy as np
import random

from dlsg import generic as G


class CapModel(nn.Module):
    """Stale in the reference (run.py is broken): Decoder(args, vocab) expects a 2H global feature that
    this wiring never provides.  Constructible for import compatibility."""

    def __init__(self, args, vocab):
        super(CapModel, self).__init__()
        self.encoder = EncoderVisual(args)
        self.decoder = Decoder(args, vocab)

    def forward(self, visual_feats, caption, max_words=None, teacher_forcing_ratio=1.0):
        visual_feats_embed = self.encoder(visual_feats)
        outputs, _ = self.decoder(visual_feats_embed, caption, max_words, teacher_forcing_ratio)
        return outputs

    def update_beam_size(self, beam_size):
        self.decoder.update_beam_size(beam_size)


class CapGnnModel(nn.Module):
    def __init__(self, args, vocab):
        super(CapGnnModel, self).__init__()
        self.use_visual_gan = args.use_visual_gan
        self.encoder = CapGnnEncoder(args)
        self.decoder = Decoder(args, vocab, multi_modal=True)

    def forward(self, visual_feats, region_feats, caption, max_words=None, teacher_forcing_ratio=1.0):
        obj_proposals, motion_proposals = self.encoder(visual_feats, region_feats)
        outputs, alpha = self.decoder._run(obj_proposals, caption, max_words, teacher_forcing_ratio, motion_proposals)
        alpha_all = alpha if alpha is not None else []          # (B, T, 2P), as torch.cat(...).transpose(1,2) gives
        return outputs, obj_proposals, motion_proposals, alpha_all

    def update_beam_size(self, beam_size):
        self.decoder.update_beam_size(beam_size)

    def load_encoder(self, model, model_path):
        model.load_state_dict(torch.load(model_path, map_location='cuda:0'))
        self.encoder = model.encoder
        self.decoder.word_embed = model.decoder.word_embed
        for param in self.decoder.word_embed.parameters():
            param.requires_grad = False


class CapGnnEncoder(nn.Module):
    def __init__(self, args, baseline=False):
        super(CapGnnEncoder, self).__init__()
        self.a_feature_size = args.a_feature_size
        self.obj_encoder = EncoderVisualGraphTUN(args, input_type='object', baseline=baseline)
        self.motion_pre_encoder = EncoderVisual(args)
        self.motion_encoder = EncoderVisualGraphTUN(args, input_type='motion', use_embed=False, baseline=baseline)

    def forward(self, visual_feats, region_feats):
        motion_input = self.motion_pre_encoder(visual_feats)
        obj_proposals, motion_proposals = tun_pair_forward(
            self.obj_encoder, visual_feats[:, :, :self.a_feature_size], self.motion_encoder, motion_input, region_feats)
        return obj_proposals, motion_proposals


class CapBaselineModel(nn.Module):
    def __init__(self, args, vocab):
        super(CapBaselineModel, self).__init__()
        self.use_visual_gan = args.use_visual_gan
        self.encoder = CapGnnEncoder(args, baseline=True)
        self.linear_baseline = nn.Linear(args.visual_hidden_size * 2, args.visual_hidden_size)
        self.decoder = Decoder(args, vocab, multi_modal=False, baseline=True)

    def forward(self, visual_feats, region_feats, caption, max_words=None, teacher_forcing_ratio=1.0):
        obj_proposals, motion_proposals = self.encoder(visual_feats, region_feats)
        outputs, _ = self.decoder(motion_proposals, caption, max_words, teacher_forcing_ratio)
        return outputs, 0, 0, 0

    def update_beam_size(self, beam_size):
        self.decoder.update_beam_size(beam_size)


class CapBaseline1(nn.Module):
    def __init__(self, args, vocab):
        super(CapBaseline1, self).__init__()
        self.use_visual_gan = args.use_visual_gan
        self.encoder = EncoderVisual(args, baseline=True)
        self.decoder = Decoder(args, vocab, multi_modal=False, baseline=True)

    def forward(self, visual_feats, region_feats, caption, max_words=None, teacher_forcing_ratio=1.0):
        visual_feats_encode = self.encoder(visual_feats)
        outputs, _ = self.decoder._run(visual_feats_encode, caption, max_words, teacher_forcing_ratio)
        return outputs, 0, 0, 0

    def update_beam_size(self, beam_size):
        self.decoder.update_beam_size(beam_size)


class DiscV2(nn.Module):
    def __init__(self, opt, vocab_size):
        super(DiscV2, self).__init__()
        self.dim = 512
        self.num_top = opt.num_topk
        self.seq_len = opt.max_words
        self.num_psl = opt.num_proposals
        self.block = nn.Sequential(
            ResBlock(self.dim),
        )
        self.conv1d = nn.Conv1d(vocab_size, self.dim, 1)
        self.lstm = nn.LSTM(512, 512, batch_first=True, bidirectional=False)
        self.layer_norm = nn.LayerNorm(512)
        self.lstm_drop = nn.Dropout(0.3)
        self.att = SelfAttention(512, 512, 512, 0.3)
        self.att_norm = nn.Sequential(
            nn.Tanh(),
            nn.LayerNorm(512)
        )
        self.motion_psl_score = PSLScore2(opt.num_proposals, self.num_top)
        self.obj_psl_score = PSLScore2(opt.num_proposals, self.num_top)
        self.text_sum = LatentPSL(512, 1)
        self.fusion = nn.Parameter(torch.empty(size=(2, 512)))
        nn.init.xavier_uniform_(self.fusion, gain=nn.init.calculate_gain('tanh'))

    @staticmethod
    def get_discriminator_block(input_dim, output_dim):
        return nn.Sequential(
            nn.Linear(input_dim, output_dim),
            nn.LeakyReLU(0.2)
        )

    @G.param_scope
    def forward(self, inputs, obj_proposals, motion_proposals, att_mask=None, alpha_all=None, _groups=1):
        """inputs (B,L,V) one-hot / logits / mix -> score (B,)   (model.py:145-168).
        _groups > 1 (our own extension, used by dlsg.gan): the batch is `_groups` independent reference calls stacked along
        dim 0 - D(real), D(fake), D(mixed) of a critic step - so the per-call batch means of PSLScore2 (layer.py:713-714)
        are taken per group and the result equals the concatenation of the separate calls."""
        p = 0.3 if self.training else 0.0
        x = G.linear(inputs, self.conv1d.weight[:, :, 0], self.conv1d.bias)          # conv1d k=1 == per-token Linear
        conv = self.block[0].res_block[1]
        y = G.resblock_blc(x, conv.weight, conv.bias)                                # relu(x) + 0.3*conv3(relu(x)), (B,L,512)
        h = G.lstm(y, self.lstm.weight_ih_l0, self.lstm.weight_hh_l0, self.lstm.bias_ih_l0, self.lstm.bias_hh_l0)
        h = G.norm(h, self.layer_norm.weight, self.layer_norm.bias, p_drop=p)
        att_out = self.att(h, att_mask)
        att_out = G.norm(att_out, self.att_norm[1].weight, self.att_norm[1].bias, pre_tanh=True)
        seq = att_mask[:, 0, :].unsqueeze(dim=2)
        alpha_all = G.mul(alpha_all, seq.expand_as(alpha_all).contiguous())
        seq_mask_spl = seq.repeat(1, 1, self.num_top)
        obj_score_out = self.obj_psl_score(obj_proposals, alpha_all[:, :, :self.num_psl], att_out, seq_mask_spl, _groups)
        motion_score_out = self.motion_psl_score(motion_proposals, alpha_all[:, :, -self.num_psl:], att_out, seq_mask_spl, _groups)
        sent_sum = self.text_sum(att_out).squeeze()                                   # (B,512)
        fusion_score = G.softmax(G.linear(sent_sum, self.fusion), dim=-1)             # (B,2)
        return G.fuse_scores(obj_score_out, motion_score_out, fusion_score, _groups)
