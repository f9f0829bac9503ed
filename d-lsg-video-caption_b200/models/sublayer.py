"""Drop-in mirror of the reference's models/sublayer.py (same class names, ctor arguments,
state_dict keys, init conventions).  torch.nn containers only HOLD parameters here; the arithmetic
of the live classes runs in libdlsg kernels via dlsg.functional / dlsg.generic.

Live:  AttentionShare :10-43, SelfAttention :46-82, PositionalEncoding_old :85-104, ResBlock :107-119,
       LatentPSL :176-198, JointEmbedVideoModel2 :292-306   (reference line numbers)
Dead in the reference (never constructed by a live model): GNN :121-144, LatentGNN :147-173,
       GraphAttentionLayer :200-289 - kept importable with identical parameters.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Parameter
from torch.autograd import Variable

from dlsg import generic as G


class AttentionShare(nn.Module):
    def __init__(self, input_value_size, input_key_size, output_size, dropout=0.1):
        super(AttentionShare, self).__init__()
        self.input_value_size = input_value_size
        self.input_key_size = input_key_size
        self.attention_size = output_size
        self.dropout = dropout
        self.K = nn.Linear(in_features=input_value_size, out_features=output_size, bias=False)
        self.Q = nn.Linear(in_features=input_key_size, out_features=output_size, bias=False)
        self.V = nn.Linear(in_features=input_value_size, out_features=output_size, bias=False)
        self.output_layer = nn.Sequential(
            nn.Linear(in_features=self.attention_size, out_features=output_size, bias=False),
            nn.Tanh(),
            nn.LayerNorm(output_size),
            nn.Dropout(self.dropout)
        )

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
        super(SelfAttention, self).__init__()
        self.attention_size = attention_size
        self.dropout = dropout
        self.get_pe = get_pe
        self.pe = PositionalEncoding_old(attention_size)
        self.K = nn.Linear(in_features=input_size, out_features=self.attention_size, bias=False)
        self.Q = nn.Linear(in_features=input_size, out_features=self.attention_size, bias=False)
        self.V = nn.Linear(in_features=input_size, out_features=self.attention_size, bias=False)
        self.output_layer = nn.Sequential(
            nn.Linear(in_features=self.attention_size, out_features=output_size, bias=False),
            nn.Dropout(self.dropout)
        )

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
    "Implement the PE function."

    def __init__(self, d_model, dropout=0.2, max_len=72):
        super(PositionalEncoding_old, self).__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0., max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        pe = pe.unsqueeze(0)
        self.register_buffer('pe', pe)

    @G.param_scope
    def forward(self, x):
        return G.add_pe(x, self.pe, self.dropout.p if self.training else 0.0)


class ResBlock(nn.Module):
    def __init__(self, dim):
        super(ResBlock, self).__init__()
        self.res_block = nn.Sequential(
            nn.ReLU(True),
            nn.Conv1d(dim, dim, 3, padding=1),
        )

    @G.param_scope
    def forward(self, input):
        """input (B,dim,L).  The reference's in-place ReLU makes this relu(x) + 0.3*conv3(relu(x))."""
        conv = self.res_block[1]
        return G.resblock(input, conv.weight, conv.bias)


class GNN(nn.Module):
    """Dead in the reference (never constructed); parameters kept for import/state_dict compatibility."""

    def __init__(self):
        super(GNN, self).__init__()
        self.adj_Q = nn.Linear(2048, 2048)
        self.adj_K = nn.Linear(2048, 2048)
        self.graph_update = nn.Linear(2048, 1024)

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
        super(LatentGNN, self).__init__()
        self.norm_func = F.normalize
        self.v2l_adj_conv = nn.Sequential(
            nn.Conv2d(in_channels=input_size, out_channels=num_latent, kernel_size=1, padding=0, bias=False),
            nn.BatchNorm2d(num_latent),
            nn.ReLU(inplace=True)
        )

    def forward(self, input_seq, mask=None):
        raise NotImplementedError('LatentGNN is dead code in the reference (sublayer.py:147-173) and outside the '
                                  'B200 hot path (SURVEY.md 8a); use LatentPSL')


class LatentPSL(nn.Module):
    def __init__(self, input_size, num_psl):
        super(LatentPSL, self).__init__()
        self.theta = nn.Parameter(torch.empty(size=(num_psl, input_size)))
        nn.init.xavier_uniform_(self.theta, gain=nn.init.calculate_gain('tanh'))
        self.out_norm = nn.Sequential(
            nn.Tanh(),
            nn.LayerNorm(input_size),
            nn.Dropout(0.3)
        )

    @G.param_scope
    def forward(self, input_seq, mask=None):
        adj = G.softmax(G.linear(input_seq, self.theta), dim=1)                   # (B,T,P) over the sequence axis
        out = G.bmm_nt(adj.transpose(1, 2), input_seq.transpose(1, 2))            # (B,P,d)
        ln = self.out_norm[1]
        return G.norm(out, ln.weight, ln.bias, pre_tanh=True, p_drop=0.3 if self.training else 0.0)


class GraphAttentionLayer(nn.Module):
    """Dead in the reference (only used by the dead EncoderVisualGAT)."""

    def __init__(self, in_features, out_features, dropout, alpha=0.2, concat=True):
        super(GraphAttentionLayer, self).__init__()
        self.dropout = dropout
        self.in_features = in_features
        self.out_features = out_features
        self.alpha = alpha
        self.concat = concat
        self.Ws = nn.Parameter(torch.empty(size=(in_features, out_features)))
        nn.init.xavier_uniform_(self.Ws, gain=nn.init.calculate_gain('relu'))
        self.We = nn.Parameter(torch.empty(size=(in_features, out_features)))
        nn.init.xavier_uniform_(self.We, gain=nn.init.calculate_gain('relu'))
        self.a = nn.Parameter(torch.empty(size=(2 * out_features, 1)))
        nn.init.xavier_uniform_(self.a.data, gain=nn.init.calculate_gain('relu'))
        self.leakyrelu = nn.LeakyReLU(self.alpha)

    def forward(self, start_feature, end_feature):
        raise NotImplementedError('GraphAttentionLayer is dead code in the reference (sublayer.py:200-289) and '
                                  'outside the B200 hot path (SURVEY.md 8a)')


class JointEmbedVideoModel2(nn.Module):
    def __init__(self, hidden_size):
        super(JointEmbedVideoModel2, self).__init__()
        self.classify = nn.Linear(hidden_size, 1)
        self.visual_embed = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.Tanh())
        self.sent_embed = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.Tanh())

    @G.param_scope
    def forward(self, visual, sent):
        v = G.linear(visual, self.visual_embed[0].weight, self.visual_embed[0].bias, tanh=True)
        s = G.linear(sent, self.sent_embed[0].weight, self.sent_embed[0].bias, tanh=True)
        return G.linear(G.mul(v, s), self.classify.weight, self.classify.bias)
