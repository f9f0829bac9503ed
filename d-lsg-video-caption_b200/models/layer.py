"""Drop-in mirror of the reference's models/layer.py: same classes, ctor arguments, attribute names,
state_dict keys and forward signatures; the arithmetic runs in libdlsg (sm_100a) kernels.

Live classes (reference line numbers): EncoderVisual :7-61, EncoderVisualGraphTUN :139-201,
Decoder :276-602, PSLScore2 :661-715.  Dead alternates kept importable with identical parameters:
EncoderVisualGraph :64-136, EncoderVisualGAT :204-272, PSLScore :605-658.
"""
from models.sublayer import *              # noqa: F401,F403  (the reference's layer.py re-exports sublayer the same way)
from models.allennlp_beamsearch import BeamSearch
import random
import os
from collections import OrderedDict

from dlsg import functional as DF
from dlsg import decoder as DD
from dlsg import generic as G
from dlsg.modspec import declare, lin, tanh_norm, lin_tanh_norm


def _named(module, prefix=''):
    """Ordered dict of the module's parameters with a name prefix (detached from nothing: autograd sees them)."""
    return OrderedDict((prefix + k, v) for k, v in module.named_parameters())


class EncoderVisual(nn.Module):
    def __init__(self, args, input_type='frame+motion', embed=True, baseline=False):
        super().__init__()
        H = self.hidden_size = args.visual_hidden_size
        self.embed, self.baseline = embed, baseline
        if embed:
            self.input_size = {'object': args.a_feature_size, 'motion': args.m_feature_size}.get(
                input_type, args.a_feature_size + args.m_feature_size)
            print('batch size', args.train_batch_size)
        p = args.dropout
        declare(self, [('linear_embed', (lambda: lin(self.input_size, H, xavier_normal=True)) if embed else None),
                       ('lstm', lambda: nn.LSTM(H, H, batch_first=True, bidirectional=True)),
                       ('layernorm_lstm', lambda: nn.LayerNorm(2 * H)),
                       ('drop_lstm', lambda: nn.Dropout(p))])
        if baseline:
            self.out_try = lin(2 * H, H, xavier_normal=True)
        else:
            declare(self, [('self_attention', lambda: SelfAttention(2 * H, 2 * H, H, p, True)),
                           ('layernorm_sa', lambda: nn.LayerNorm(H)),
                           ('drop_sa', lambda: nn.Dropout(p))])

    def _init_lstm_state(self, d):
        """Zero (h, c) for the two directions (layer.py:40-44); the fused BiLSTM starts from zeros implicitly."""
        z = d.new_zeros(2, d.size(0), self.hidden_size)
        return z, z.clone()

    def forward(self, inputs):
        if not self.embed:
            raise NotImplementedError('EncoderVisual(embed=False) is never constructed by the reference models')
        t = OrderedDict(frames=inputs)
        t.update(_named(self))
        if not self.baseline:
            t['pe'] = self.self_attention.pe.pe
        blk = DF.EncoderVisualBlock('', self.baseline, self.drop_lstm.p, self.training)
        return DF.run_block(blk, t)[0]


def _graph_encoder_parts(self, args, input_type, use_embed, baseline):
    """Parameter containers shared by the three graph-encoder variants of the reference (layer.py:64-136, 139-201, 204-272),
    in the reference's registration order up to and including `obj_visual_norm`."""
    H = args.visual_hidden_size
    self.baseline, self.use_embed = baseline, use_embed
    with_regions = args.num_obj > 4
    vis_in = args.m_feature_size if input_type == 'motion' else args.a_feature_size
    declare(self, [('obj_embed', (lambda: lin(args.region_feature_size, args.region_projected_size)) if with_regions else None),
                   ('obj_norm', (lambda: tanh_norm(args.region_projected_size)) if with_regions else None),
                   ('visual_embed', (lambda: lin(vis_in, H)) if use_embed else None),
                   ('visual_norm', lambda: tanh_norm(H)),
                   ('obj_visual_norm', lambda: tanh_norm(H))])
    return H


class EncoderVisualGraph(nn.Module):
    """Dead alternate (reference layer.py:64-136, commented out in model.py:61). Parameters only."""

    def __init__(self, args, input_type='motion', use_embed=True, baseline=False):
        super().__init__()
        H = _graph_encoder_parts(self, args, input_type, use_embed, baseline)
        declare(self, [('v2l_layer', lambda: LatentPSL(H, args.num_proposals)),
                       ('att_l2l', lambda: SelfAttention(H, H, H, args.dropout)),
                       ('att_l2l_norm', lambda: nn.LayerNorm(H))])

    def forward(self, visual_feats, obj_feats):
        raise NotImplementedError('EncoderVisualGraph is a dead alternate in the reference (model.py:61 is commented '
                                  'out); the live encoder is EncoderVisualGraphTUN')


class EncoderVisualGraphTUN(nn.Module):
    def __init__(self, args, input_type='motion', use_embed=True, baseline=False):
        super().__init__()
        H = _graph_encoder_parts(self, args, input_type, use_embed, baseline)
        self.num_proposals = args.num_proposals
        self.norm_func = F.normalize
        declare(self, [('v2l_layer', lambda: LatentPSL(H, args.num_proposals)),
                       ('att_l2l_norm', lambda: nn.LayerNorm(H)),          # constructed, never used (reference quirk)
                       ('drop_o2v', lambda: nn.Dropout(args.dropout)),
                       ('drop_v2l', lambda: nn.Dropout(args.dropout))])

    def _used(self, prefix=''):
        """Parameters that take part in forward (att_l2l_norm never does - reference quirk)."""
        return OrderedDict((prefix + k, v) for k, v in self.named_parameters() if not k.startswith('att_l2l_norm'))

    def forward(self, visual_feats, obj_feats):
        t = OrderedDict(regions=obj_feats, visual0=visual_feats)
        t.update(self._used())
        blk = DF.TunBlock([{'prefix': '', 'use_embed': self.use_embed}], self.num_proposals, self.training, self.baseline)
        return DF.run_block(blk, t)[0]


def tun_pair_forward(enc_a, visual_a, enc_b, visual_b, obj_feats):
    """Both graph encoders of CapGnnEncoder in one block: the region features are converted once and
    projected by ONE GEMM over the concatenated obj_embed weights (layer.py:184)."""
    t = OrderedDict(regions=obj_feats, visual0=visual_a, visual1=visual_b)
    t.update(enc_a._used('a.'))
    t.update(enc_b._used('b.'))
    blk = DF.TunBlock([{'prefix': 'a.', 'use_embed': enc_a.use_embed}, {'prefix': 'b.', 'use_embed': enc_b.use_embed}],
                      enc_a.num_proposals, enc_a.training, enc_a.baseline)
    return DF.run_block(blk, t)


class EncoderVisualGAT(nn.Module):
    """Dead alternate (reference layer.py:204-272, commented out in model.py:62). Parameters only."""

    def __init__(self, args, input_type='motion', use_embed=True, baseline=False):
        super().__init__()
        H = _graph_encoder_parts(self, args, input_type, use_embed, baseline)
        declare(self, [('o2v_gat', lambda: GraphAttentionLayer(H, H, args.dropout)),
                       ('v2l_layer', lambda: LatentPSL(H, args.num_proposals)),
                       ('att_l2l', lambda: SelfAttention(H, H, H, args.dropout)),
                       ('att_l2l_norm', lambda: nn.LayerNorm(H))])

    def forward(self, visual_feats, obj_feats):
        raise NotImplementedError('EncoderVisualGAT is a dead alternate in the reference (model.py:62 is commented out)')


class Decoder(nn.Module):
    def __init__(self, args, vocab, multi_modal=False, baseline=False, use_fusion=False):
        super().__init__()
        H, W, Hq, Hd, p = args.visual_hidden_size, args.word_size, args.query_hidden_size, args.decode_hidden_size, args.dropout
        self.vocab, self.vocab_size, self.dataset = vocab, len(vocab), args.dataset
        self.word_size, self.max_words, self.beam_size, self.batch_size = W, args.max_words, args.beam_size, args.train_batch_size
        self.query_hidden_size, self.decode_hidden_size = Hq, Hd
        self.multi_modal, self.use_fusion = multi_modal, use_fusion
        two_heads = multi_modal and not use_fusion
        query_in = H + W + Hd + (0 if baseline else H)            # [lang_h | global feature (H or 2H) | word]
        lang_in = H + Hq + (H if two_heads else 0)                # [ctx (| ctx2) | q]
        att = lambda: AttentionShare(input_value_size=H, input_key_size=Hq, output_size=H)
        declare(self, [('beta_fusion', (lambda: nn.Sequential(lin(2 * H, 1), nn.Sigmoid())) if (multi_modal and use_fusion) else None),
                       ('word_embed', lambda: nn.Embedding(self.vocab_size, W))])
        if args.use_glove:
            self.get_glove_embedding()
        declare(self, [('word_drop', lambda: nn.Dropout(p=p)),
                       ('query_lstm', lambda: nn.LSTMCell(query_in, Hq)),
                       ('query_lstm_layernorm', lambda: nn.LayerNorm(Hq)),
                       ('query_lstm_drop', lambda: nn.Dropout(p=p)),
                       ('lang_lstm', lambda: nn.LSTMCell(lang_in, Hd)),
                       ('lang_lstm_layernorm', lambda: nn.LayerNorm(Hd)),
                       ('lang_lstm_drop', lambda: nn.Dropout(p=p)),
                       ('context_att', att),
                       ('context_layernorm', lambda: nn.LayerNorm(Hd)),      # constructed, never used (reference quirk)
                       ('context_att_2', att if multi_modal else None),
                       ('word_restore', lambda: lin(Hd, self.vocab_size, xavier_normal=True))])
        self.update_beam_size(self.beam_size)

    def update_beam_size(self, beam_size):
        self.beam_size = beam_size
        self.beam_search = BeamSearch(self.vocab('<end>'), self.max_words, beam_size, per_node_beam_size=beam_size)

    def get_glove_embedding(self):
        """Load the (V, word_size) GloVe table the reference prepares offline (layer.py:352-386) into word_embed."""
        path = './data/%s_glove.npy' % self.dataset
        if not os.path.exists(path):
            raise FileNotFoundError('%s not found (GloVe table is prepared offline by the reference, layer.py:352-386)' % path)
        self.word_embed.load_state_dict({'weight': torch.from_numpy(np.load(path))})

    def _init_lstm_state(self, d, hidden_size):
        """Zero (h, c) of one LSTMCell (layer.py:388-392)."""
        z = d.new_zeros(d.size(0), hidden_size)
        return z, z.clone()

    def _used(self):
        """Parameters that take part in decoding (context_layernorm never does - reference quirk)."""
        return OrderedDict((k, v) for k, v in self.named_parameters()
                           if not k.startswith('context_layernorm') and not k.startswith('beta_fusion'))

    def forward(self, cnn_feats, captions, max_words, teacher_forcing_ratio, cnn_feats_2=None, step_feats=None):
        outputs, alpha = self._run(cnn_feats, captions, max_words, teacher_forcing_ratio, cnn_feats_2, step_feats)
        if alpha is None:
            return outputs, []
        return outputs, [alpha[:, i, :].unsqueeze(-1) for i in range(alpha.shape[1])]     # list of (B, nh*P, 1)

    def _run(self, cnn_feats, captions, max_words, teacher_forcing_ratio, cnn_feats_2=None, step_feats=None):
        """Same as forward but returns the attention weights as one (B, T, nh*P) tensor (or None at inference)."""
        if self.use_fusion or step_feats is not None:
            raise NotImplementedError('use_fusion / step_feats are never used by the reference models')
        self.batch_size = cnn_feats.size(0)
        infer = captions is None
        if max_words is None:
            max_words = self.max_words
        if not infer:
            # one Python-RNG draw per step, exactly as layer.py:432 (consumed before any kernel runs)
            tf = [random.random() < teacher_forcing_ratio for _ in range(max_words)]
            t = OrderedDict(n1=cnn_feats)
            if cnn_feats_2 is not None:
                t['n2'] = cnn_feats_2
            t['captions'] = captions
            t.update(self._used())
            blk = DD.DecoderTrainBlock('', self.multi_modal, self.word_drop.p, self.training, max_words, tf)
            logits, alpha = DF.run_block(blk, t)
            return logits, alpha
        with torch.no_grad():
            t = OrderedDict((k, v.detach()) for k, v in self._used().items())
            if self.beam_size == 1:
                outputs = DD.decode_greedy(t, '', self.multi_modal, cnn_feats.detach(),
                                           None if cnn_feats_2 is None else cnn_feats_2.detach(), max_words)
            else:
                outputs, _, _ = DD.decode_beam(t, '', self.multi_modal, cnn_feats.detach(),
                                               None if cnn_feats_2 is None else cnn_feats_2.detach(), self.max_words,
                                               self.beam_size, self.vocab('<end>'), self.beam_search.per_node_beam_size)
        return outputs, None

    def decode_tokens(self, tokens):
        """Word ids -> caption string, up to (not including) the first <end> (layer.py:464-477)."""
        ids = tokens.tolist() if torch.is_tensor(tokens) else list(tokens)       # one D2H copy, not one sync per token
        end = self.vocab('<end>')
        n = ids.index(end) if end in ids else len(ids)
        return ' '.join(self.vocab.idx2word[i] for i in ids[:n])

    def caption2wordembedding(self, caption):
        with torch.no_grad():
            return G.embedding(caption, self.word_embed.weight)

    def output2wordembedding(self, ouput):
        return G.matmul_nn(ouput, self.word_embed.weight.detach())

    def beam_step(self, last_predictions, current_state):
        """AllenNLP step-function contract (layer.py:489-567): (group,) ids + state dict -> (log-probs (group,V), state).
        All beams are decoded in one batched pass; constant node tensors are indexed, never restacked."""
        with torch.no_grad():
            t = OrderedDict((k, v.detach()) for k, v in self._used().items())
            return DD.beam_step_api(t, self.multi_modal, self.batch_size, last_predictions, current_state)

    def decode(self, word, query_lstm_h, query_lstm_c, lang_lstm_h, lang_lstm_c, global_feat, cnn_feats, cnn_feats_2=None):
        """Single step with explicit state (layer.py:569-602); inference-only convenience over the same kernels."""
        with torch.no_grad():
            t = OrderedDict((k, v.detach()) for k, v in self._used().items())
            return DD.decode_api(t, self.multi_modal, word, query_lstm_h, query_lstm_c, lang_lstm_h, lang_lstm_c,
                                 global_feat, cnn_feats, cnn_feats_2)


class _PSLBase(nn.Module):
    def __init__(self, num_psl, num_top):
        super().__init__()
        declare(self, [('psl_scorer', lambda: JointEmbedVideoModel2(512)),
                       ('psl_embed', lambda: lin_tanh_norm(1024, 512)),
                       ('psl_norm', lambda: tanh_norm(512, 0.3)),
                       ('att_norm', lambda: lin_tanh_norm(512, 512))])
        self.num_top = num_top
        self.select = num_psl > num_top                       # keep the top-`num_top` nodes by attention mass (layer.py:684-686)

    def _common(self, psl, psl_alpha, att_out):
        bs = psl.size(0)
        p = G.linear(psl, self.psl_embed[0].weight, self.psl_embed[0].bias)
        p = G.norm(p, self.psl_embed[2].weight, self.psl_embed[2].bias, pre_tanh=True)
        if self.select:
            idx = G.topk_indices(G.sum_dim1(psl_alpha), self.num_top)
            p = G.gather_rows(p, idx)
        a = G.linear(att_out, self.att_norm[0].weight, self.att_norm[0].bias)
        a = G.norm(a, self.att_norm[2].weight, self.att_norm[2].bias, pre_tanh=True)
        adj = G.bmm_nt(a, p)                                   # (B, L, K) raw scores
        return p, a, adj


class PSLScore(_PSLBase):
    """Dead alternate (reference layer.py:605-658); same parameters as PSLScore2."""

    @G.param_scope
    def forward(self, psl, psl_alpha, att_out, seq_mask):
        p, a, adj = self._common(psl, psl_alpha, att_out)
        adj = G.softmax(adj, dim=1, scale=1.0 / math.sqrt(512), mask=seq_mask, mask_mode=1)
        g = G.bmm_nt(adj.transpose(1, 2), a.transpose(1, 2))
        g = G.norm(g, self.psl_norm[1].weight, self.psl_norm[1].bias, pre_tanh=True, p_drop=0.3 if self.training else 0.0)
        s = self.psl_scorer(p, g).squeeze()
        return s.mean(axis=-1)


class PSLScore2(_PSLBase):
    @G.param_scope
    def forward(self, psl, psl_alpha, att_out, seq_mask, _groups=1):
        """layer.py:688-715 -> 0-dim scalar (batch mean of alpha-weighted node scores); (_groups,) when the batch stacks
        several independent calls (dlsg.gan)."""
        p, a, adj = self._common(psl, psl_alpha, att_out)
        adj = G.softmax(adj, dim=1, scale=1.0 / math.sqrt(512), mask=seq_mask, mask_mode=2)   # over words, then zero pads
        adj_alpha = G.sum_dim1(adj)                                                          # (B,K)
        g = G.bmm_nt(adj.transpose(1, 2), a.transpose(1, 2))                                  # (B,K,512)
        g = G.norm(g, self.psl_norm[1].weight, self.psl_norm[1].bias, pre_tanh=True, p_drop=0.3 if self.training else 0.0)
        s = self.psl_scorer(p, g).squeeze()
        return G.weighted_mean_score(s, adj_alpha, _groups)
