"""Deterministic synthetic weights / inputs shaped like the RMN feature files.

Everything is a closed-form function of (name, shape) so that the golden-vector
generator (run once against the reference, tests/golden/make_golden.py), the CPU
oracle tests and the GPU parity tests all see bit-identical tensors without
shipping weights.  Input layout follows the reference loader
(utils/data.py:55-63: frames (T, Da+Dm) f32, regions (T, 36, Dr) f32) and the
special token ids of utils/utils.py:17-20 (<pad>=0 <start>=1 <end>=2 <unk>=3).
"""
import types
import zlib

import numpy as np
import torch

PAD, START, END, UNK = 0, 1, 2, 3


class Vocab:
    """Minimal stand-in for utils/utils.py:12-43 (callable, len, idx2word)."""

    def __init__(self, size):
        self.idx2word = ['<pad>', '<start>', '<end>', '<unk>'] + ['w%d' % i for i in range(4, size)]
        self.word2idx = {w: i for i, w in enumerate(self.idx2word)}
        self.nwords = size

    def __call__(self, w):
        return self.word2idx.get(w, UNK)

    def __len__(self):
        return self.nwords


def make_args(**kw):
    """argparse-like namespace with the fields the model ctors read."""
    d = dict(use_visual_gan=True, a_feature_size=1536, m_feature_size=2048, visual_hidden_size=1024,
             train_batch_size=64, dropout=0.3, num_obj=36, region_feature_size=2048,
             region_projected_size=1024, num_proposals=5, word_size=300, max_words=26,
             max_frames=26, dataset='msr-vtt', beam_size=5, query_hidden_size=1024,
             decode_hidden_size=1536, use_glove=False, num_topk=5)
    d.update(kw)
    return types.SimpleNamespace(**d)


def msr_args(**kw):
    """run_gun.py:36-40 overrides for msr-vtt, BASELINE Dm=2048."""
    return make_args(**kw)


def msvd_args(**kw):
    """run_gun.py:31-35 overrides for msvd."""
    d = dict(dataset='msvd', decode_hidden_size=1024, num_proposals=8, num_obj=16, num_topk=3)
    d.update(kw)
    return make_args(**d)


def small_args(**kw):
    """Reduced widths for golden vectors; structure identical (R>=5 so the region path runs)."""
    d = dict(a_feature_size=48, m_feature_size=32, visual_hidden_size=64, region_feature_size=40,
             region_projected_size=64, num_obj=6, num_proposals=5, word_size=20, max_words=7,
             max_frames=6, query_hidden_size=64, decode_hidden_size=96, num_topk=5,
             train_batch_size=3, beam_size=5)
    d.update(kw)
    return make_args(**d)


def _rs(name):
    return np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)


def fill_tensor(name, shape, kind=None):
    """Deterministic fp32 tensor for a state_dict entry."""
    shape = tuple(shape)
    n = int(np.prod(shape)) if len(shape) else 1
    r = _rs(name).standard_normal(n).astype(np.float32).reshape(shape)
    leaf = name.split('.')[-1]
    if kind is None:
        if leaf == 'pe':
            kind = 'keep'
        elif len(shape) >= 2:
            kind = 'matrix'
        elif leaf.startswith('weight') and len(shape) == 1:
            kind = 'ln_weight'
        else:
            kind = 'bias'
    if kind == 'matrix':
        fan_in = int(np.prod(shape[1:]))
        if 'word_embed' in name:
            r *= 0.5
        elif leaf == 'theta' or leaf == 'fusion':
            r *= 1.5 / np.sqrt(fan_in)
        else:
            r *= 1.0 / np.sqrt(fan_in)
    elif kind == 'ln_weight':
        r = 1.0 + 0.1 * r
    elif kind == 'bias':
        r *= 0.1
    return torch.from_numpy(np.ascontiguousarray(r))


def fill_state_dict(module_or_sd, prefix=''):
    """Overwrite every parameter (not buffers named 'pe') of a module / state_dict in place."""
    sd = module_or_sd.state_dict() if hasattr(module_or_sd, 'state_dict') else module_or_sd
    out = {}
    for k, v in sd.items():
        if k.endswith('pe.pe'):
            out[k] = v.detach().clone().float()
            continue
        out[k] = fill_tensor(prefix + k, v.shape).to(v.dtype)
    if hasattr(module_or_sd, 'load_state_dict'):
        module_or_sd.load_state_dict(out)
    return out


def make_inputs(B, args, V, seed=12, dist='relu', full_len=False):
    """frames (B,T,Da+Dm), regions (B,T,R,Dr), captions (B,L) i64, cap_lens list.

    dist='relu' -> relu(randn)*0.5 (pooled CNN features are non-negative); 'randn' = stress.
    Captions: len ~ U{4..L}, tokens U{4..V-1}, <end> at len-1, <pad> after.
    """
    rs = np.random.RandomState(seed)
    T, R, L = args.max_frames, args.num_obj, args.max_words
    fr = rs.standard_normal((B, T, args.a_feature_size + args.m_feature_size)).astype(np.float32)
    rg = rs.standard_normal((B, T, R, args.region_feature_size)).astype(np.float32)
    if dist == 'relu':
        fr = np.maximum(fr, 0) * 0.5
        rg = np.maximum(rg, 0) * 0.5
    caps = np.zeros((B, L), dtype=np.int64)
    lens = []
    for b in range(B):
        n = L if full_len else int(rs.randint(min(4, L), L + 1))
        caps[b, :n - 1] = rs.randint(4, V, size=n - 1)
        caps[b, n - 1] = END
        lens.append(n)
    return (torch.from_numpy(fr), torch.from_numpy(rg), torch.from_numpy(caps), lens)


def att_mask_from_captions(captions):
    """run_gun.py:164-166: outer product of the (captions>0) mask."""
    m = (captions > 0).to(torch.float32)
    return m.unsqueeze(2) * m.unsqueeze(1)
