"""Drop-in mirror of the reference's models/layer.py: same classes, ctor arguments, attribute names,
state_dict keys and forward signatures; the arithmetic runs in libdlsg (sm_100a) kernels.

Live classes (reference line numbers): EncoderVisual :7-61, EncoderVisualGraphTUN :139-201,
Decoder :276-602, PSLScore2 :661-715.  Dead alternates kept importable with identical parameters:
EncoderVisualGraph :64-136, EncoderVisualGAT :204-272, PSLScore :605-658.
"""
from models.sublayer import *
from models.allennlp_beamsearch import BeamSearch
import random
import os
from collections import OrderedDict

from dlsg import functional as DF
from dlsg import decoder as DD
from dlsg import generic as G


def _named(module, prefix=''):
    """Ordered dict of the module's parameters with a name prefix (detached from nothing: autograd sees them)."""
    return OrderedDict((prefix + k, v) for k, v in module.named_parameters())


class EncoderVisual(nn.Module):
    def __init__(self, args, input_type='frame+motion', embed=True, baseline=False):
        super(EncoderVisual, self).__init__()
        self.embed = embed
        hidden_size = args.visual_hidden_size
        self.hidden_size = hidden_size
        if embed:
            input_size = args.a_feature_size + args.m_feature_size
            if input_type == 'object':
                input_size = args.a_feature_size
            if input_type == 'motion':
                input_size = args.m_feature_size
            self.input_size = input_size
            print('batch size', args.train_batch_size)
            self.linear_embed = nn.Linear(input_size, hidden_size)
            nn.init.xavier_normal_(self.linear_embed.weight)
        self.lstm = nn.LSTM(hidden_size, hidden_size, batch_first=True, bidirectional=True)
        self.layernorm_lstm = nn.LayerNorm(hidden_size * 2)
        self.drop_lstm = nn.Dropout(args.dropout)
        self.baseline = baseline
        if not self.baseline:
            self.self_attention = SelfAttention(hidden_size * 2, hidden_size * 2, hidden_size, args.dropout, True)
            self.layernorm_sa = nn.LayerNorm(hidden_size)
            self.drop_sa = nn.Dropout(args.dropout)
        else:
            self.out_try = nn.Linear(hidden_size * 2, hidden_size)
            nn.init.xavier_normal_(self.out_try.weight)

    def _init_lstm_state(self, d):
        batch_size = d.size(0)
        lstm_state_h = d.data.new(2, batch_size, self.hidden_size).zero_()
        lstm_state_c = d.data.new(2, batch_size, self.hidden_size).zero_()
        return lstm_state_h, lstm_state_c

    def forward(self, inputs):
        if not self.embed:
            raise NotImplementedError('EncoderVisual(embed=False) is never constructed by the reference models')
        t = OrderedDict(frames=inputs)
        t.update(_named(self))
        if not self.baseline:
            t['pe'] = self.self_attention.pe.pe
        blk = DF.EncoderVisualBlock('', self.baseline, self.drop_lstm.p, self.training)
        return DF.run_block(blk, t)[0]


class EncoderVisualGraph(nn.Module):
    """Dead alternate (reference layer.py:64-136, commented out in model.py:61). Parameters only."""

    def __init__(self, args, input_type='motion', use_embed=True, baseline=False):
        super(EncoderVisualGraph, self).__init__()
        self.baseline = baseline
        H = args.visual_hidden_size
        if args.num_obj > 4:
            self.obj_embed = nn.Linear(args.region_feature_size, args.region_projected_size)
            self.obj_norm = nn.Sequential(nn.Tanh(), nn.LayerNorm(args.region_projected_size))
        visual_input_size = args.m_feature_size if input_type == 'motion' else args.a_feature_size
        self.use_embed = use_embed
        if self.use_embed:
            self.visual_embed = nn.Linear(visual_input_size, H)
        self.visual_norm = nn.Sequential(nn.Tanh(), nn.LayerNorm(H))
        self.obj_visual_norm = nn.Sequential(nn.Tanh(), nn.LayerNorm(H))
        self.v2l_layer = LatentPSL(H, args.num_proposals)
        self.att_l2l = SelfAttention(H, H, H, args.dropout)
        self.att_l2l_norm = nn.LayerNorm(H)

    def forward(self, visual_feats, obj_feats):
        raise NotImplementedError('EncoderVisualGraph is a dead alternate in the reference (model.py:61 is commented '
                                  'out); the live encoder is EncoderVisualGraphTUN')


class EncoderVisualGraphTUN(nn.Module):
    def __init__(self, args, input_type='motion', use_embed=True, baseline=False):
        super(EncoderVisualGraphTUN, self).__init__()
        self.baseline = baseline
        if args.num_obj > 4:
            self.obj_embed = nn.Linear(args.region_feature_size, args.region_projected_size)
            self.obj_norm = nn.Sequential(
                nn.Tanh(),
                nn.LayerNorm(args.region_projected_size)
            )
        visual_input_size = args.m_feature_size
        if input_type != 'motion':
            visual_input_size = args.a_feature_size
        self.use_embed = use_embed
        if self.use_embed:
            self.visual_embed = nn.Linear(visual_input_size, args.visual_hidden_size)
        self.visual_norm = nn.Sequential(
            nn.Tanh(),
            nn.LayerNorm(args.visual_hidden_size)
        )
        self.obj_visual_norm = nn.Sequential(
            nn.Tanh(),
            nn.LayerNorm(args.visual_hidden_size),
        )
        self.v2l_layer = LatentPSL(args.visual_hidden_size, args.num_proposals)
        self.att_l2l_norm = nn.LayerNorm(args.visual_hidden_size)
        self.norm_func = F.normalize
        self.drop_o2v = nn.Dropout(args.dropout)
        self.drop_v2l = nn.Dropout(args.dropout)
        self.num_proposals = args.num_proposals

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
        super(EncoderVisualGAT, self).__init__()
        self.baseline = baseline
        H = args.visual_hidden_size
        if args.num_obj > 4:
            self.obj_embed = nn.Linear(args.region_feature_size, args.region_projected_size)
            self.obj_norm = nn.Sequential(nn.Tanh(), nn.LayerNorm(args.region_projected_size))
        visual_input_size = args.m_feature_size if input_type == 'motion' else args.a_feature_size
        self.use_embed = use_embed
        if self.use_embed:
            self.visual_embed = nn.Linear(visual_input_size, H)
        self.visual_norm = nn.Sequential(nn.Tanh(), nn.LayerNorm(H))
        self.obj_visual_norm = nn.Sequential(nn.Tanh(), nn.LayerNorm(H))
        self.o2v_gat = GraphAttentionLayer(H, H, args.dropout)
        self.v2l_layer = LatentPSL(H, args.num_proposals)
        self.att_l2l = SelfAttention(H, H, H, args.dropout)
        self.att_l2l_norm = nn.LayerNorm(H)

    def forward(self, visual_feats, obj_feats):
        raise NotImplementedError('EncoderVisualGAT is a dead alternate in the reference (model.py:62 is commented out)')


class Decoder(nn.Module):
    def __init__(self, args, vocab, multi_modal=False, baseline=False, use_fusion=False):
        super(Decoder, self).__init__()
        self.word_size = args.word_size
        self.max_words = args.max_words
        self.vocab = vocab
        self.dataset = args.dataset
        self.vocab_size = len(vocab)
        self.beam_size = args.beam_size
        self.batch_size = args.train_batch_size
        self.query_hidden_size = args.query_hidden_size
        self.decode_hidden_size = args.decode_hidden_size
        self.multi_modal = multi_modal
        self.use_fusion = use_fusion
        if multi_modal and use_fusion:
            self.beta_fusion = nn.Sequential(nn.Linear(2 * args.visual_hidden_size, 1), nn.Sigmoid())
        self.word_embed = nn.Embedding(self.vocab_size, self.word_size)
        if args.use_glove:
            self.get_glove_embedding()
        self.word_drop = nn.Dropout(p=args.dropout)
        query_input_size = args.visual_hidden_size + args.word_size + args.decode_hidden_size
        if baseline is False:
            query_input_size += args.visual_hidden_size
        self.query_lstm = nn.LSTMCell(query_input_size, args.query_hidden_size)
        self.query_lstm_layernorm = nn.LayerNorm(args.query_hidden_size)
        self.query_lstm_drop = nn.Dropout(p=args.dropout)
        lang_decode_hidden_size = args.visual_hidden_size + args.query_hidden_size
        if self.multi_modal and self.use_fusion is False:
            lang_decode_hidden_size += args.visual_hidden_size
        self.lang_lstm = nn.LSTMCell(lang_decode_hidden_size, args.decode_hidden_size)
        self.lang_lstm_layernorm = nn.LayerNorm(args.decode_hidden_size)
        self.lang_lstm_drop = nn.Dropout(p=args.dropout)
        self.context_att = AttentionShare(input_value_size=args.visual_hidden_size,
                                          input_key_size=args.query_hidden_size,
                                          output_size=args.visual_hidden_size)
        self.context_layernorm = nn.LayerNorm(args.decode_hidden_size)
        if self.multi_modal:
            self.context_att_2 = AttentionShare(input_value_size=args.visual_hidden_size,
                                                input_key_size=args.query_hidden_size,
                                                output_size=args.visual_hidden_size)
        self.word_restore = nn.Linear(args.decode_hidden_size, self.vocab_size)
        nn.init.xavier_normal_(self.word_restore.weight)
        self.beam_search = BeamSearch(vocab('<end>'), self.max_words, self.beam_size, per_node_beam_size=self.beam_size)

    def update_beam_size(self, beam_size):
        self.beam_size = beam_size
        self.beam_search = BeamSearch(self.vocab('<end>'), self.max_words, beam_size, per_node_beam_size=beam_size)

    def get_glove_embedding(self):
        glove_np_path = f'./data/{self.dataset}_glove.npy'
        if not os.path.exists(glove_np_path):
            raise FileNotFoundError('%s not found (GloVe table is prepared offline by the reference, layer.py:352-386)'
                                    % glove_np_path)
        weight_matrix = torch.from_numpy(np.load(glove_np_path))
        self.word_embed.load_state_dict({'weight': weight_matrix})

    def _init_lstm_state(self, d, hidden_size):
        batch_size = d.size(0)
        lstm_state_h = d.data.new(batch_size, hidden_size).zero_()
        lstm_state_c = d.data.new(batch_size, hidden_size).zero_()
        return lstm_state_h, lstm_state_c

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
        '''convert word index to caption'''
        if torch.is_tensor(tokens):
            tokens = tokens.tolist()                     # one D2H copy instead of one sync per token
        words = []
        end = self.vocab('<end>')
        for token in tokens:
            if token == end:
                break
            words.append(self.vocab.idx2word[token])
        return ' '.join(words)

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
        super(_PSLBase, self).__init__()
        self.psl_scorer = JointEmbedVideoModel2(512)
        self.psl_embed = nn.Sequential(nn.Linear(1024, 512), nn.Tanh(), nn.LayerNorm(512))
        self.psl_norm = nn.Sequential(nn.Tanh(), nn.LayerNorm(512), nn.Dropout(0.3))
        self.att_norm = nn.Sequential(nn.Linear(512, 512), nn.Tanh(), nn.LayerNorm(512))
        self.num_top = num_top
        self.select = True
        if num_psl <= self.num_top:
            self.select = False

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
