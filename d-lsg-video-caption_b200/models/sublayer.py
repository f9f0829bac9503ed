"""Drop-in mirror of the reference's models/sublayer.py: same class names, constructor arguments, state_dict keys and
initialisation.  The torch.nn containers only HOLD parameters (declared through dlsg.modspec tables); the arithmetic of the
live classes runs in libdlsg kernels via dlsg.generic.

Live (reference line numbers): AttentionShare :10-43, SelfAttention :46-82, PositionalEncoding_old :85-104, ResBlock
:107-119, LatentPSL :176-198, JointEmbedVideoModel2 :292-306.  Dead in the reference (never constructed by a live model) but
kept importable with identical parameters: GNN :121-144, LatentGNN :147-173, GraphAttentionLayer :200-289.
"""
import math

import numpy as np                      # noqa: F401  (re-exported: models.layer star-imports this module, as the reference does)
import torch
import torch.nn as nn
import torch.nn.functional as F

from dlsg import generic as G
from dlsg.modspec import declare, lin, tanh_norm, xavier_uniform_param, sinusoid_table


class AttentionShare(nn.Module):
    def __init__(self, input_value_size, input_key_size, output_size, dropout=0.1):
        super().__init__()
        dv, dk, d = input_value_size, input_key_size, output_size
        self.input_value_size, self.input_key_size, self.attention_size, self.dropout = dv, dk, d, dropout
        declare(self, [('K', lambda: lin(dv, d, bias=False)),
                       ('Q', lambda: lin(dk, d, bias=False)),
                       ('V', lambda: lin(dv, d, bias=False)),
                       ('output_layer', lambda: nn.Sequential(lin(d, d, bias=False), nn.Tanh(), nn.LayerNorm(d), nn.Dropout(dropout)))])

    @G.param_scope
    def forward(self, meta_state, hidden_previous):
        """(B,P,Dv),(B,Dk) -> (attention (B,out), weight (B,P,1)); softmax over the node axis."""
        K = G.linear(meta_state, self.K.weight)
        V = G.linear(meta_state, self.V.weight)
        Q = G.linear(hidden_previous, self.Q.weight).unsqueeze(1)                 # (B,1,d)
        logits = G.bmm_nt(K, Q)                                                   # (B,P,1)
        weight = G.softmax(logits, dim=1, scale=1.0 / math.sqrt(self.attention_size))
        mid = G.bmm_nt(weight.transpose(1, 2), V.transpose(1, 2)).squeeze(1)      # (B,d)
        out = G.linear(mid, self.output_layer[0].weight)
        ln = self.output_layer[2]
        out = G.norm(out, ln.weight, ln.bias, pre_tanh=True, p_drop=self.dropout if self.training else 0.0)
        return out, weight


class SelfAttention(nn.Module):
    def __init__(self, input_size, attention_size, output_size, dropout=0.2, get_pe=False):
        super().__init__()
        d = attention_size
        self.attention_size, self.dropout, self.get_pe = d, dropout, get_pe
        declare(self, [('pe', lambda: PositionalEncoding_old(d)),
                       ('K', lambda: lin(input_size, d, bias=False)),
                       ('Q', lambda: lin(input_size, d, bias=False)),
                       ('V', lambda: lin(input_size, d, bias=False)),
                       ('output_layer', lambda: nn.Sequential(lin(d, output_size, bias=False), nn.Dropout(dropout)))])

    @G.param_scope
    def forward(self, x, att_mask=None):
        if self.get_pe:
            x = self.pe(x)
        K = G.linear(x, self.K.weight)
        Q = G.linear(x, self.Q.weight)
        V = G.linear(x, self.V.weight)
        logits = G.bmm_nt(K, Q)
        weight = G.softmax(logits, dim=-1, scale=1.0 / math.sqrt(self.attention_size), mask=att_mask, mask_mode=1)
        att = G.bmm_nt(weight, V.transpose(1, 2))
        out = G.linear(att, self.output_layer[0].weight)
        return G.dropout(out, self.dropout if self.training else 0.0)


class PositionalEncoding_old(nn.Module):
    """Sinusoidal position table added to the sequence, then dropout (buffer 'pe' of shape (1, max_len, d_model))."""

    def __init__(self, d_model, dropout=0.2, max_len=72):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.register_buffer('pe', sinusoid_table(max_len, d_model))

    @G.param_scope
    def forward(self, x):
        return G.add_pe(x, self.pe, self.dropout.p if self.training else 0.0)


class ResBlock(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.res_block = nn.Sequential(nn.ReLU(True), nn.Conv1d(dim, dim, 3, padding=1))     # key 'res_block.1.*'

    @G.param_scope
    def forward(self, input):
        """input (B,dim,L).  The reference's in-place ReLU makes this relu(x) + 0.3*conv3(relu(x))."""
        conv = self.res_block[1]
        return G.resblock(input, conv.weight, conv.bias)


class GNN(nn.Module):
    """Dead in the reference (never constructed); parameters kept for import/state_dict compatibility."""

    def __init__(self):
        super().__init__()
        declare(self, [('adj_Q', lambda: lin(2048, 2048)), ('adj_K', lambda: lin(2048, 2048)),
                       ('graph_update', lambda: lin(2048, 1024))])

    @G.param_scope
    def forward(self, region_feats):
        bs, win_len, num_obj, fs = region_feats.shape
        feats = region_feats.contiguous().view(bs, win_len * num_obj, fs)
        q = G.linear(feats, self.adj_Q.weight, self.adj_Q.bias)
        k = G.linear(feats, self.adj_K.weight, self.adj_K.bias)
        adj = G.softmax(G.bmm_nt(q, k), dim=-1)
        upd = G.linear(feats, self.graph_update.weight, self.graph_update.bias)
        return G.bmm_nt(adj, upd.transpose(1, 2)).view(bs, win_len, num_obj, -1)


class LatentGNN(nn.Module):
    """Dead in the reference (uses BatchNorm2d; never constructed by a live model)."""

    def __init__(self, input_size, num_latent, norm_func):
        super().__init__()
        self.norm_func = F.normalize
        self.v2l_adj_conv = nn.Sequential(nn.Conv2d(input_size, num_latent, kernel_size=1, padding=0, bias=False),
                                          nn.BatchNorm2d(num_latent), nn.ReLU(inplace=True))

    def forward(self, input_seq, mask=None):
        raise NotImplementedError('LatentGNN is dead code in the reference (sublayer.py:147-173) and outside the '
                                  'B200 hot path (SURVEY.md 8a); use LatentPSL')


class LatentPSL(nn.Module):
    def __init__(self, input_size, num_psl):
        super().__init__()
        declare(self, [('theta', lambda: xavier_uniform_param(num_psl, input_size, 'tanh')),
                       ('out_norm', lambda: tanh_norm(input_size, 0.3))])

    @G.param_scope
    def forward(self, input_seq, mask=None):
        adj = G.softmax(G.linear(input_seq, self.theta), dim=1)                   # (B,T,P) over the sequence axis
        out = G.bmm_nt(adj.transpose(1, 2), input_seq.transpose(1, 2))            # (B,P,d)
        ln = self.out_norm[1]
        return G.norm(out, ln.weight, ln.bias, pre_tanh=True, p_drop=0.3 if self.training else 0.0)


class GraphAttentionLayer(nn.Module):
    """Dead in the reference (only used by the dead EncoderVisualGAT)."""

    def __init__(self, in_features, out_features, dropout, alpha=0.2, concat=True):
        super().__init__()
        self.dropout, self.in_features, self.out_features, self.alpha, self.concat = dropout, in_features, out_features, alpha, concat
        declare(self, [('Ws', lambda: xavier_uniform_param(in_features, out_features, 'relu')),
                       ('We', lambda: xavier_uniform_param(in_features, out_features, 'relu')),
                       ('a', lambda: xavier_uniform_param(2 * out_features, 1, 'relu')),
                       ('leakyrelu', lambda: nn.LeakyReLU(alpha))])

    def forward(self, start_feature, end_feature):
        raise NotImplementedError('GraphAttentionLayer is dead code in the reference (sublayer.py:200-289) and '
                                  'outside the B200 hot path (SURVEY.md 8a)')


class JointEmbedVideoModel2(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        h = hidden_size
        declare(self, [('classify', lambda: lin(h, 1)),
                       ('visual_embed', lambda: nn.Sequential(lin(h, h), nn.Tanh())),
                       ('sent_embed', lambda: nn.Sequential(lin(h, h), nn.Tanh()))])

    @G.param_scope
    def forward(self, visual, sent):
        v = G.linear(visual, self.visual_embed[0].weight, self.visual_embed[0].bias, tanh=True)
        s = G.linear(sent, self.sent_embed[0].weight, self.sent_embed[0].bias, tanh=True)
        return G.linear(G.mul(v, s), self.classify.weight, self.classify.bias)
