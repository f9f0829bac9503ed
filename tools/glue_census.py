"""Census of the torch (non-libdlsg) operations a critic step still launches: runs dlsg.gan.GanIteration's critic steps on the
CPU emulation backend (tests/cpu_emul.py) under a TorchDispatchMode that counts every aten op executed OUTSIDE a backend
call, grouped by the innermost dlsg / models source line that issued it.  Development tool (no GPU needed): every op counted
here is an eager torch launch on the GPU path.      python tools/glue_census.py [--top 40] [--bf16]
"""
import argparse
import collections
import contextlib
import io
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'd-lsg-video-caption_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch
from torch.utils._python_dispatch import TorchDispatchMode

from cpu_emul import CpuEmulBackend
from dlsg import synth, ops, linalg as la, functional as DF

DEPTH = int(os.environ.get('CENSUS_DEPTH', '1'))
SKIP = {'aten.view.default', 'aten.detach.default', 'aten.t.default', 'aten.transpose.int', 'aten.unsqueeze.default',
        'aten.expand.default', 'aten.slice.Tensor', 'aten.select.int', 'aten._unsafe_view.default', 'aten.as_strided.default',
        'aten.alias.default', 'aten.squeeze.dim', 'aten.squeeze.default', 'aten.permute.default', 'aten.unbind.int',
        'aten.reshape.default', 'aten.empty.memory_format', 'aten.empty_like.default', 'aten.split.Tensor',
        'aten.empty_strided.default', 'aten.new_empty.default', 'aten.is_same_size.default', 'aten._local_scalar_dense.default',
        'aten.view_as.default', 'aten.result_type.Tensor', 'aten.lift_fresh.default', 'aten.split_with_sizes.default'}


class Census(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.inside = 0
        self.by_site = collections.Counter()
        self.by_op = collections.Counter()
        self.by_kernel = collections.Counter()
        self.kernel_site = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not self.inside and name not in SKIP:
            site = []
            for fr in reversed(traceback.extract_stack(limit=40)):
                fn = fr.filename
                if ('/dlsg/' in fn or '/models/' in fn) and 'glue_census' not in fn:
                    site.append('%s:%d' % (os.path.basename(fn), fr.lineno))
                    if len(site) == DEPTH:
                        break
            site = ' < '.join(site) or '?'
            ts = [x for x in list(args) + list((kwargs or {}).values()) if torch.is_tensor(x)]
            if any(not x.is_contiguous() for x in ts) or len({tuple(x.shape) for x in ts if x.dim() > 0}) > 1:
                name += '  [strided/broadcast %s]' % ','.join('x'.join(map(str, x.shape)) for x in ts[:2])
            self.by_site[(site, name)] += 1
            self.by_op[name] += 1
        return func(*args, **(kwargs or {}))


def wrap_backend(be, census):
    for k in dir(be):
        if k.startswith('_') or k in ('launches',):
            continue
        f = getattr(be, k)
        if not callable(f):
            continue

        def mk(f, k=k):
            def g(*a, **kw):
                if not census.inside and k not in ('make_convert_plan', 'make_adam_plan', 'fused_step_supported', 'attn2_supported',
                                                   'lstm_step_supported', 'region_aggregate_supported', 'latent_psl_supported'):
                    census.by_kernel[k] += 1
                    site = []
                    for fr in reversed(traceback.extract_stack(limit=40)):
                        fn = fr.filename
                        if ('/dlsg/' in fn or '/models/' in fn) and not fn.endswith(('ops.py', 'linalg.py')):
                            site.append('%s:%d' % (os.path.basename(fn), fr.lineno))
                            if len(site) == DEPTH:
                                break
                    site = ' < '.join(site)
                    census.kernel_site[(k, site)] += 1
                census.inside += 1
                try:
                    return f(*a, **kw)
                finally:
                    census.inside -= 1
            return g
        setattr(be, k, mk(f))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--top', type=int, default=60)
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--T", type=int, default=26)
    ap.add_argument("--old-lstm", action="store_true")
    a = ap.parse_args()
    import models.model as M
    from dlsg.gan import GanIteration
    from dlsg import generic as GN
    GN.FUSED_LSTM_BPTT2 = not a.old_lstm
    be = CpuEmulBackend()
    ops.set_backend(be)
    la.set_precision('bf16' if a.bf16 else 'fp32')
    args = synth.small_args(visual_hidden_size=1024, region_projected_size=1024, query_hidden_size=1024, max_words=a.T, max_frames=4)
    V, B = 37, 3
    frames, regions, caps, lens = synth.make_inputs(B, args, V, seed=21)
    with contextlib.redirect_stdout(io.StringIO()):
        G_ = M.CapGnnModel(args, synth.Vocab(V))
    D_ = M.DiscV2(args, V)
    synth.fill_state_dict(G_)
    synth.fill_state_dict(D_, prefix='D.')
    G_.eval()
    D_.eval()
    og = torch.optim.Adam(G_.parameters(), lr=1.6e-4, betas=(0.5, 0.9))
    od = torch.optim.Adam(D_.parameters(), lr=1.6e-4, betas=(0.5, 0.9))
    it = GanIteration(G_, D_, og, od, frames, regions, caps, lens, args.max_words, 1.0, num_d=1, gan_lambda=0.05, graph=False)
    L = args.max_words
    seq = (caps[:, :L] > 0).to(torch.float32)
    att_mask = seq.unsqueeze(2) * seq.unsqueeze(1)
    with torch.no_grad():
        f_cap, obj, mot, alpha = G_(frames, regions, caps, L, 1.0)
    census = Census()
    wrap_backend(be, census)
    l0 = be.launches
    la.set_manual_param_epochs(True)
    la.new_param_epoch()
    with census:
        it._disc_steps(caps[:, :L].contiguous(), f_cap, obj, mot, att_mask, alpha)
    la.set_manual_param_epochs(False)
    print('one critic step (T=%d): %d libdlsg launches, %d torch ops outside them' % (L, be.launches - l0, sum(census.by_op.values())))
    print('--- libdlsg entries')
    for k, v in census.by_kernel.most_common():
        print('%5d  %s' % (v, k))
    for (k, site), v in census.kernel_site.most_common(a.top):
        print('%5d  %-18s %s' % (v, k, site))
    print('--- by op')
    for k, v in census.by_op.most_common(30):
        print('%5d  %s' % (v, k))
    print('--- by site')
    for (site, name), v in census.by_site.most_common(a.top):
        print('%5d  %-34s %s' % (v, site, name))


if __name__ == '__main__':
    main()
