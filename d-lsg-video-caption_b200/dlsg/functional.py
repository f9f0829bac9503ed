"""Host orchestration of the D-LSG blocks over the libdlsg kernels: explicit forward AND backward
for each block (no autograd inside), wrapped by `run_block` into one torch.autograd.Function per
block so that the reference trainers' autograd / DDP / Adam drive them unchanged.

Blocks (reference file:line they replace):
  TunBlock            models/layer.py:172-201  EncoderVisualGraphTUN (+ sublayer.py:189-198 LatentPSL),
                      E encoders share one region-projection GEMM (regions are read once).
  EncoderVisualBlock  models/layer.py:46-61    Linear -> BiLSTM -> LN -> SelfAttention(PE) -> LN
  DecoderBlock        models/layer.py:394-447,569-602  26-step two-LSTM decoder with node attention
"""
import itertools
import math
import os

import torch

from . import linalg as la
from . import ops
from .linalg import empty, zeros, small_zeros, op, op_empty, op_zeros, mm32, flat2

# DLSG_FUSED_REGION_AGG=0 keeps the unfused LayerNorm / GEMM / softmax / GEMM composition (measurement switch)
FUSED_REGION_AGG = os.environ.get('DLSG_FUSED_REGION_AGG', '1') != '0'
# DLSG_FUSED_LSTM_STEP=0 keeps the recurrent GEMM + cell kernel pair per BiLSTM step (measurement switch)
FUSED_LSTM_STEP = os.environ.get('DLSG_FUSED_LSTM_STEP', '1') != '0'

_seed_counter = itertools.count(1)


def next_seed():
    """Per-call dropout seed derived from torch's seed (deterministic under torch.manual_seed)."""
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + next(_seed_counter) * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF


def site(p, seed, sid):
    return (float(p), seed, sid << 32) if p > 0 else None


WC = la.WeightCache()


class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, block, names, *tensors):
        t = dict(zip(names, tensors))
        # weight-copy cache scope: a training forward always re-converts its weights (see linalg.WeightCache)
        ctx.wc_scope = WC.begin_train_block() if getattr(block, 'need_grad', False) else WC.eval_scope()
        outs, saved = block.forward(t)
        ctx.block, ctx.saved, ctx.names = block, saved, names
        ctx.pids = {n: id(x) for n, x in zip(names, tensors) if isinstance(x, torch.nn.Parameter)}
        if isinstance(block, TunBlock):
            # an encoder of the pair whose output nobody uses (CapBaselineModel drops the object nodes, model.py:86-88) must see
            # None, not a zero tensor, so that its parameters end with grad None exactly as under the reference's autograd
            ctx.set_materialize_grads(False)
        for o in outs:
            if o.dtype not in (torch.float32,):
                ctx.mark_non_differentiable(o)
        return tuple(outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gouts):
        if BLOCK_BWD_HOOK is not None:
            BLOCK_BWD_HOOK(ctx.block)         # (every AccumulateGrad of the blocks that ran before this one has fired)
        WC.set_scope(ctx.wc_scope)
        if GRAD_SYNC is not None:
            GRAD_SYNC.pids = ctx.pids
        la.begin_pool(gouts[0].device if gouts[0] is not None else next(g for g in gouts if g is not None).device)
        try:
            grads = ctx.block.backward(ctx.saved, gouts)
        finally:
            la.end_pool()
        ctx.saved = None
        if GRAD_SYNC is not None:
            GRAD_SYNC.reduce(grads)
            GRAD_SYNC.skip.clear()
        return (None, None) + tuple(grads.get(n) for n in ctx.names)


BLOCK_INPUTS = ('regions', 'visual0', 'visual1', 'frames', 'pe', 'n1', 'n2', 'captions')
def _rows2(x):
    """2-D view of a gradient / bucket slice for the multi-segment kernels (segments are spread over CTAs by rows)."""
    if x.dim() == 2:
        return x
    return x.reshape(x.shape[0], -1) if x.dim() > 2 else x.reshape(1, -1)


GRAD_SYNC = None
BLOCK_BWD_HOOK = None       # callable(block) invoked at the start of every block backward (dlsg.graphs: overlapped Adam)


class GradSync:
    """Data-parallel gradient averaging for a captured (or eager) step; the only collective of the path (run_gun.py:63-72).

    Every reduce() call (one per block backward, plus the early ones in the middle of a block) owns a PERSISTENT flat
    bucket: ONE multi-segment launch (dlsg_multi_convert) packs that call's parameter gradients into it - cast to bf16 by
    default, so the bytes on NVLink and the bytes Adam reads are halved - and the all-reduce (AVG = SUM / world, the same
    arithmetic as DistributedDataParallel) runs on ONE side stream, in the same order on every rank, the moment the
    gradients exist: the decoder bucket travels while the encoder backward computes; only the last bucket is exposed.
    No torch.cat, no per-step allocation, no unpacking: `reduced[id(param)]` is a view of the bucket in the parameter's
    shape and dlsg_adam_multi reads it in place (AdamDriver.step(grad_of=...)).  p.grad keeps the LOCAL gradient.

    dtype: torch.bfloat16 (default; DLSG_GRAD_REDUCE=fp32 selects torch.float32 buckets)."""

    def __init__(self, process_group=None, dtype=None):
        import os
        import torch.distributed as dist
        self.dist, self.pg = dist, process_group
        if dtype is None:
            dtype = torch.float32 if os.environ.get('DLSG_GRAD_REDUCE', 'bf16') == 'fp32' else torch.bfloat16
        self.dtype = dtype
        self.side = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.skip = set()           # ids of gradient tensors that early_sync() already reduced in the running block backward
        self.pids = {}              # name -> id(parameter) of the block whose backward is running (set by _BlockFn.backward)
        self.reduced = {}           # id(parameter) -> bucket view (parameter shape, self.dtype) holding the averaged gradient
        self._buckets = {}          # call signature -> flat bucket
        self._plans = []            # pack plans of the current step (kept alive for graph replays)
        self.bytes = 0              # bytes all-reduced in the last step
        self.debug = set(x for x in os.environ.get('DLSG_SYNC_DEBUG', '').split(',') if x)   # measurement aids: nonccl | nopack | nowait | pgrad | noearly | norecord

    def begin_step(self):
        self.reduced.clear()
        self.skip.clear()
        self._plans = []
        self.bytes = 0

    def reduce(self, grads):
        items = [(k, v) for k, v in grads.items() if v is not None and k not in BLOCK_INPUTS and id(v) not in self.skip]
        if not items:
            return
        uniq = {}
        for k, v in items:
            uniq.setdefault(id(v), v)
        # a 2-D row-pitched view (the recurrent-weight slices of the packed gate gradients) is packed in place through its pitch
        tensors = [x if (x.is_contiguous() or (x.dim() == 2 and x.stride(1) == 1)) else x.contiguous() for x in uniq.values()]
        sig = tuple((k, tuple(v.shape)) for k, v in items)
        offs, n = [], 0
        for x in tensors:
            offs.append(n)
            n += (x.numel() + 7) // 8 * 8                      # 16-byte aligned slices (vector loads in pack and Adam)
        flat = self._buckets.get(sig)
        if flat is None or flat.device != tensors[0].device:
            flat = self._buckets[sig] = torch.empty(n, dtype=self.dtype, device=tensors[0].device)
        views = {}
        pairs = []
        for key, x, off in zip(uniq.keys(), tensors, offs):
            v = flat[off:off + x.numel()].view(x.shape)
            views[key] = v
            pairs.append((_rows2(x), None, _rows2(v)))
        be = ops.backend()
        plan = be.make_convert_plan(pairs, host=True)
        self._plans.append(plan)
        if self.side is not None:
            cur = torch.cuda.current_stream()
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                if 'norecord' not in self.debug:
                    for x in tensors:
                        x.record_stream(self.side)
                if 'nopack' not in self.debug:
                    be.multi_convert(plan)
                if not (self.debug & {'nopack', 'nonccl'}):
                    self.dist.all_reduce(flat, op=self.dist.ReduceOp.AVG, group=self.pg)
        else:                                                   # gloo on CPU (tests): no AVG, no streams
            be.multi_convert(plan)
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.pg)
            flat.div_(self.dist.get_world_size(self.pg))
        self.bytes += flat.numel() * flat.element_size()
        for k, v in items:
            pid = self.pids.get(k)
            if pid is not None:
                self.reduced[pid] = views[id(v)]
            self.skip.add(id(v))

    def wait(self):
        if self.side is not None and 'nowait' not in self.debug:
            torch.cuda.current_stream().wait_stream(self.side)
        self.skip.clear()

    def grad_of(self, p):
        """Averaged gradient of parameter p (bucket view) - what the optimizer consumes."""
        if 'pgrad' in self.debug:
            return p.grad
        return self.reduced.get(id(p))

    def reduce_params(self, params):
        """Average p.grad of `params` across the ranks in place, on the CURRENT stream (the critic's gradients,
        run_gun.py:71-72: pack -> all-reduce -> unpack, three launches, one persistent bucket)."""
        ps = [p for p in params if p.grad is not None]
        if not ps:
            return
        sig = ('params',) + tuple((id(p), tuple(p.shape)) for p in ps)
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += (p.numel() + 7) // 8 * 8
        flat = self._buckets.get(sig)
        if flat is None:
            flat = self._buckets[sig] = torch.zeros(n, dtype=self.dtype, device=ps[0].device)
        # 2-D views for the multi-segment kernel, which spreads a segment over CTAs by ROWS: a tensor of more than two dims keeps
        # its leading dim as rows (the critic's conv1d weight (512, V, 1) as ONE row of 5.4 M elements was converted by a single
        # CTA: 4.1 ms per critic step, tools/diag_critic_sync.py)
        two = _rows2
        fwd, back = [], []
        for p, off in zip(ps, offs):
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            if g is not p.grad:
                p.grad = g
            v = flat[off:off + p.numel()].view(p.shape)
            fwd.append((two(g), None, two(v)))
            back.append((two(v), None, two(g)))
        be = ops.backend()
        plan_f, plan_b = be.make_convert_plan(fwd, host=True), be.make_convert_plan(back, host=True)
        self._plans += [plan_f, plan_b]
        be.multi_convert(plan_f)
        if flat.is_cuda:
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.AVG, group=self.pg)
        else:
            self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.pg)
            flat.div_(self.dist.get_world_size(self.pg))
        be.multi_convert(plan_b)
        self.bytes += flat.numel() * flat.element_size()

    def write_back(self, params):
        """p.grad <- averaged gradient (fp32), for callers that hand the gradients to a foreign optimizer or inspect them
        (tests): one multi-segment launch on the current stream.  Call after wait()."""
        pairs = []
        for p in params:
            r = self.reduced.get(id(p))
            if r is not None and p.grad is not None:
                pairs.append((_rows2(r), None, _rows2(p.grad)))
        if pairs:
            be = ops.backend()
            plan = be.make_convert_plan(pairs, host=True)
            self._plans.append(plan)
            be.multi_convert(plan)


def early_sync(grads, keys=None):
    """Called by a block in the middle of its backward with the parameter gradients that are already final (everything
    in `grads` so far): their all-reduce starts now and overlaps the rest of the block (the decoder's vocabulary-projection
    gradients travel during its 26-step BPTT, the EncoderVisual attention gradients during the BiLSTM BPTT)."""
    if GRAD_SYNC is None or 'noearly' in GRAD_SYNC.debug:
        return
    GRAD_SYNC.reduce(grads if keys is None else {k: grads[k] for k in keys if k in grads})


def run_block(block, tensors):
    """tensors: ordered dict name -> tensor (parameters and inputs). Returns tuple of outputs."""
    names = tuple(tensors.keys())
    block.need_grad = torch.is_grad_enabled() and any(v.requires_grad for v in tensors.values())
    return _BlockFn.apply(block, names, *[tensors[n] for n in names])


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


_SIDE = {}


def two_streams(like, f0, f1):
    """Run two independent launch sequences concurrently: f0 on the current stream, f1 on a side stream (fork/join by
    events, so it is CUDA-graph capturable).  Used for the two directions of the BiLSTM, whose 26-step chains are
    latency-bound and each use a fraction of the SMs."""
    if not like.is_cuda:
        f0()
        f1()
        return
    cur = torch.cuda.current_stream()
    side = _SIDE.get(like.device)
    if side is None:
        side = _SIDE[like.device] = torch.cuda.Stream(device=like.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        f1()
    f0()
    cur.wait_stream(side)


# =============================================================================================== TUN
class TunBlock:
    """encs: list of dicts {'prefix': str, 'use_embed': bool}.  Tensor names: 'regions', 'visual<e>',
    '<prefix>obj_embed.weight' ... exactly the module's parameter names (prefix-qualified)."""

    def __init__(self, encs, P, training, baseline=False):
        self.encs, self.P, self.training, self.baseline = encs, P, training, baseline

    def forward(self, t):
        be = ops.backend()
        regions = t['regions']
        B, T, R, Dr = regions.shape
        E = len(self.encs)
        TR, M = T * R, B * T * R
        sv = {'dims': (B, T, R, Dr), 'enc': []}
        use_regions = R >= 5
        need_grad = getattr(self, 'need_grad', True)
        outs = []
        if use_regions:
            R2 = _c(regions).view(M, Dr)
            Rb = op(R2)                 # one cast; the weight-gradient GEMM reads it transposed in place
            RbT = Rb.t()
            H = t[self.encs[0]['prefix'] + 'obj_embed.weight'].shape[0]
            ws = [t[e['prefix'] + 'obj_embed.weight'] for e in self.encs]
            bs = [t[e['prefix'] + 'obj_embed.bias'] for e in self.encs]

            def build():
                Wc = op_empty((E * H,), Dr, R2)
                bc = empty((E * H,), R2)
                for i, (w, b) in enumerate(zip(ws, bs)):
                    WC.cv(w.detach(), Wc[i * H:(i + 1) * H])
                    WC.cv(b.detach().view(1, H), bc[i * H:(i + 1) * H].view(1, H))
                return Wc, bc
            Wc, bc = WC.packed(('tunW',) + tuple(id(w) for w in ws), ws, la.pver(*ws, *bs), build)
            Ot = empty((M, E * H), R2, la.opdtype())
            be.gemm(Rb, Wc, Ot, bias=bc, tanh=True)
            sv.update(RbT=RbT, Ot=Ot)
        seed = next_seed()
        fused = use_regions and be.region_aggregate_supported(T, TR, H, Ot.dtype) and FUSED_REGION_AGG
        sv['fused'] = fused
        scale = 1.0 / math.sqrt(Dr)
        for i, e in enumerate(self.encs):          # frame vectors F = LN(tanh(visual_embed(v)))  (layer.py:176-181)
            pf = e['prefix']
            v = _c(t['visual%d' % i])
            H = t[pf + 'visual_norm.1.weight'].shape[0]
            v2 = v.view(B * T, v.shape[-1])
            s = {}
            if e['use_embed']:
                w = t[pf + 'visual_embed.weight']
                Fv = la.mm(v2, WC.get(w), bias=t[pf + 'visual_embed.bias'])
                s['v2'] = v2
            else:
                Fv = v2
            F = empty((B * T, H), v2)
            Fop = None if (fused or not use_regions) else op_empty((B * T,), H, v2)
            stF = empty((B * T, 2), v2)
            be.norm_fwd(Fv, t[pf + 'visual_norm.1.weight'], t[pf + 'visual_norm.1.bias'], y=F, y2=Fop, stats=stF,
                        pre_tanh=True)
            s.update(Fv=Fv, F=F, Fop=Fop, stF=stF, scale=scale)
            sv['enc'].append(s)
        if fused:
            # ONE launch for both encoders: obj_norm + scores + softmax over the T*R regions + weighted sum, Y read once
            # (csrc/region_agg.cu); the backward recomputes the normalised rows from Y and the saved row statistics
            S_ = sv['enc']
            for s in S_:
                s['agg'] = empty((B * T, H), S_[0]['F'])
                if need_grad:
                    s.update(U=empty((B * T, H), s['F']), stO=empty((M, 2), s['F']), St=empty((B, T, TR), s['F']),
                             tcF=empty((B * T, 4), s['F']))
            opt = (lambda k: [s[k] for s in S_]) if need_grad else (lambda k: None)
            be.region_aggregate_fwd([Ot[:, i * H:(i + 1) * H] for i in range(E)], [s['F'] for s in S_],
                                    [t[e['prefix'] + 'obj_norm.1.weight'].detach() for e in self.encs],
                                    [t[e['prefix'] + 'obj_norm.1.bias'].detach() for e in self.encs], scale, T,
                                    agg=[s['agg'] for s in S_], U=opt('U'), stats=opt('stO'), St=opt('St'), tconst=opt('tcF'))
        for i, e in enumerate(self.encs):
            pf, s = e['prefix'], sv['enc'][i]
            F = s['F']
            v2 = F
            H = F.shape[1]
            if use_regions:
                if not fused:
                    O = empty((M, H), v2, la.opdtype())
                    stO = empty((M, 2), v2)
                    be.norm_fwd(Ot[:, i * H:(i + 1) * H], t[pf + 'obj_norm.1.weight'], t[pf + 'obj_norm.1.bias'], y=O, stats=stO)
                    O3 = O.view(B, TR, H)
                    St = empty((B, T, TR), v2)                      # raw scores, transposed: (frame, object)
                    be.gemm(s['Fop'].view(B, T, H), O3, St)
                    Sm = empty((B, T, TR), v2)
                    be.softmax_fwd(St, Sm, dim=2, scale=scale)       # softmax over the T*R objects (layer.py:188 dim=1)
                    OT = op(O3.transpose(1, 2))                      # (B,H,TR) K-major over objects
                    agg = empty((B, T, H), v2)
                    be.gemm(op(Sm), OT, agg)
                    s.update(O=O, stO=stO, St=St, Sm=Sm, OT=OT, agg=agg)
                agg = s['agg']
                X = empty((B * T, H), v2)
                stX = empty((B * T, 2), v2)
                be.norm_fwd(agg.view(B * T, H), t[pf + 'obj_visual_norm.1.weight'], t[pf + 'obj_visual_norm.1.bias'],
                            y=X, res=F, stats=stX, pre_tanh=True)
                s.update(X=X, stX=stX)
            else:
                X = F
                s['X'] = X
            if self.baseline:
                outs.append(X.view(B, T, H))
        if not self.baseline:
            # LatentPSL pooling (sublayer.py:189-198) of BOTH encoders in one launch, then the output LayerNorms
            P = self.P
            S_ = sv['enc']
            H = S_[0]['X'].shape[-1]
            like = S_[0]['X']
            thetas = [t[e['prefix'] + 'v2l_layer.theta'] for e in self.encs]
            for s in S_:
                s.update(Gs=empty((B, T, P), like), N=empty((B, P, H), like), G=None)
            if be.latent_psl_supported(T, P, H):
                be.latent_psl_fwd_multi([s['X'].view(B, T, H) for s in S_], [th.detach() for th in thetas], [s['Gs'] for s in S_],
                                        [s['N'] for s in S_])
            else:
                for s, theta in zip(S_, thetas):
                    X = s['X']
                    s['G'] = mm32(X, theta.detach())                                   # (B*T,P)
                    be.softmax_fwd(s['G'].view(B, T, P), s['Gs'], dim=1)              # over the T frames (sublayer.py:192)
                    be.gemm(s['Gs'].transpose(1, 2), X.view(B, T, H).transpose(1, 2), s['N'])
            for i, e in enumerate(self.encs):
                pf, s = e['prefix'], S_[i]
                nodes = empty((B, P, H), like)
                stN = empty((B * P, 2), like)
                dn = site(0.3 if self.training else 0.0, seed, i)
                be.norm_fwd(s['N'].view(B * P, H), t[pf + 'v2l_layer.out_norm.1.weight'], t[pf + 'v2l_layer.out_norm.1.bias'],
                            y=nodes.view(B * P, H), stats=stN, pre_tanh=True, drop=dn)
                s.update(stN=stN, dn=dn)
                outs.append(nodes)
        sv['t'] = t
        return outs, sv

    def backward(self, sv, gouts):
        be = ops.backend()
        t = sv['t']
        B, T, R, Dr = sv['dims']
        TR, M = T * R, B * T * R
        E = len(self.encs)
        use_regions = R >= 5
        fused = sv.get('fused', False)
        grads = {}
        dOpre = dbc = None
        active = [i for i in range(E) if gouts[i] is not None]

        def lnp(pf, name):
            w, b = t[pf + name + '.weight'], t[pf + name + '.bias']
            dw, db = small_zeros(w.shape, w), small_zeros(b.shape, b)
            grads[pf + name + '.weight'], grads[pf + name + '.bias'] = dw, db
            return w, b, dw, db
        dXs, dAs, dFs = {}, {}, {}
        if self.baseline:
            for i in active:
                X = sv['enc'][i]['X']
                dXs[i] = _c(gouts[i]).view(B, T, X.shape[-1])
        elif active:
            # output LayerNorm backward per encoder, then the pooling backward of all active encoders in ONE launch
            P = self.P
            dN3s, ths, dths = {}, {}, {}
            for i in active:
                pf, s = self.encs[i]['prefix'], sv['enc'][i]
                X = s['X']
                H = X.shape[-1]
                w, b, dw, db = lnp(pf, 'v2l_layer.out_norm.1')
                dN = empty((B * P, H), X)
                be.norm_bwd(_c(gouts[i]).view(B * P, H), s['N'].view(B * P, H), w, b, s['stN'], dx=dN, dgamma=dw, dbeta=db,
                            pre_tanh=True, drop=s['dn'])
                dN3s[i] = dN.view(B, P, H)
                dXs[i] = empty((B, T, H), X)
                ths[i] = t[pf + 'v2l_layer.theta'].detach()
            fusedp = [i for i in active if sv['enc'][i]['G'] is None]
            if fusedp:
                for i in fusedp:
                    dths[i] = small_zeros(ths[i].shape, sv['enc'][i]['X'])
                be.latent_psl_bwd_multi([sv['enc'][i]['X'].view(B, T, -1) for i in fusedp], [ths[i] for i in fusedp],
                                        [sv['enc'][i]['Gs'] for i in fusedp], [dN3s[i] for i in fusedp], [dXs[i] for i in fusedp],
                                        [dths[i] for i in fusedp])
            for i in active:
                if i in fusedp:
                    continue
                s = sv['enc'][i]
                X, dN3, dX, theta = s['X'], dN3s[i], dXs[i], ths[i]
                H = X.shape[-1]
                dGs = empty((B, T, P), X)
                be.gemm(X.view(B, T, H), dN3, dGs)
                be.gemm(s['Gs'], dN3.transpose(1, 2), dX)
                dG = empty((B, T, P), X)
                be.softmax_bwd(s['G'].view(B, T, P), dGs, dG, dim=1)
                be.gemm(dG.view(B * T, P), theta.t(), dX.view(B * T, H), accum=True)
                dths[i] = empty(theta.shape, X)
                be.gemm(dG.view(B * T, P).t(), X.t(), dths[i])
            for i in active:
                grads[self.encs[i]['prefix'] + 'v2l_layer.theta'] = dths[i]
        for i in active:                             # obj_visual_norm backward: dA = d(agg + F)
            pf, s = self.encs[i]['prefix'], sv['enc'][i]
            X = s['X']
            H = X.shape[-1]
            dX = dXs[i]
            if use_regions:
                w, b, dw, db = lnp(pf, 'obj_visual_norm.1')
                dA = empty((B * T, H), X)
                be.norm_bwd(dX.view(B * T, H), s['agg'].view(B * T, H), w, b, s['stX'], dx=dA, res=s['F'], dgamma=dw,
                            dbeta=db, pre_tanh=True)
                dAs[i] = dA
        if use_regions and active:
            X = sv['enc'][active[0]]['X']
            H = X.shape[-1]
            dOpre = empty((M, E * H), X, la.opdtype())
            if len(active) < E:
                dOpre.zero_()
            dbc = small_zeros((E * H,), dOpre)
        if use_regions and active and fused:
            # pass 1: dSm = dA . LN(Y) (the forward kernel, scores only); pass 2: everything else, Y read once more and
            # d(pre-activation of the region projection) written once (csrc/region_agg.cu)
            S_ = [sv['enc'][i] for i in active]
            Ys = [sv['Ot'][:, i * H:(i + 1) * H] for i in active]
            lns = [lnp(self.encs[i]['prefix'], 'obj_norm.1') for i in active]
            gam, bet = [l[0].detach() for l in lns], [l[1].detach() for l in lns]
            dA_l = [dAs[i] for i in active]
            dSm = [empty((B, T, TR), X) for _ in active]
            tcA = [empty((B * T, 4), X) for _ in active]
            be.region_aggregate_fwd(Ys, dA_l, gam, bet, S_[0]['scale'], T, St=dSm, tconst=tcA, scores_only=True)
            nwork = be.region_aggregate_bwd_workspace(B, T, TR)
            for i in active:
                dFs[i] = empty((B * T, H), X)
            be.region_aggregate_bwd(Ys, [s['stO'] for s in S_], [s['St'] for s in S_], dSm,
                                    [s['F'] for s in S_], dA_l, [s['U'] for s in S_], [s['tcF'] for s in S_], tcA, gam, bet,
                                    S_[0]['scale'], T, dpre=[dOpre[:, i * H:(i + 1) * H] for i in active],
                                    dF=[dFs[i] for i in active], dgamma=[l[2] for l in lns], dbeta=[l[3] for l in lns],
                                    dbias=[dbc[i * H:(i + 1) * H] for i in active],
                                    work=[empty((nwork,), X, torch.uint8) for _ in active])
        for i in active:
            e = self.encs[i]
            pf, s = e['prefix'], sv['enc'][i]
            X = s['X']
            H = X.shape[-1]
            if use_regions and not fused:
                dA = dAs[i]
                dF = empty((B * T, H), X)
                be.axpby(dA, 1.0, dF, 0.0)
                O3 = s['O'].view(B, TR, H)
                dAop = op(dA.view(B, T, H))
                dSm = empty((B, T, TR), X)
                be.gemm(dAop, O3, dSm)
                dSt = empty((B, T, TR), X)
                be.softmax_bwd(s['St'], dSm, dSt, dim=2, scale=s['scale'])
                # dO = Sm^T dA + dSt^T F  as one GEMM with K = 2T
                Ac = op_zeros((B, TR), 2 * T, X)
                Bc = op_zeros((B, H), 2 * T, X)
                be.convert(s['Sm'], dstT=Ac[:, :, :T])
                be.convert(dSt, dstT=Ac[:, :, T:2 * T])
                be.convert(dA.view(B, T, H), dstT=Bc[:, :, :T])
                be.convert(s['F'].view(B, T, H), dstT=Bc[:, :, T:2 * T])
                dO = empty((M, H), X, la.opdtype())
                be.gemm(Ac, Bc, dO.view(B, TR, H))
                be.gemm(op(dSt), s['OT'], dF.view(B, T, H), accum=True)
                w, b, dw, db = lnp(pf, 'obj_norm.1')
                # the region-projection bias gradient (column sums of dOpre) is accumulated by the same kernel
                be.norm_bwd(dO, sv['Ot'][:, i * H:(i + 1) * H], w, b, s['stO'], dx=dOpre[:, i * H:(i + 1) * H], dgamma=dw,
                            dbeta=db, in_is_tanh=True, dxsum=dbc[i * H:(i + 1) * H])
            elif use_regions:
                dF = dFs[i]
            else:
                dF = dXs[i].view(B * T, H)
            w, b, dw, db = lnp(pf, 'visual_norm.1')
            dFv = empty((B * T, H), X)
            be.norm_bwd(dF, s['Fv'], w, b, s['stF'], dx=dFv, dgamma=dw, dbeta=db, pre_tanh=True)
            if e['use_embed']:
                grads[pf + 'visual_embed.weight'] = la.mm(dFv.t(), s['v2'].t())
                dbv = small_zeros((H,), X)
                be.colsum(dFv, dbv)
                grads[pf + 'visual_embed.bias'] = dbv
            else:
                grads['visual%d' % i] = dFv.view(B, T, H)
        if use_regions and dOpre is not None:
            H = dOpre.shape[1] // E
            dWc = empty((E * H, Dr), dOpre)
            be.gemm(op(dOpre.t()), sv['RbT'], dWc)
            for i, e in enumerate(self.encs):
                if gouts[i] is None:
                    continue        # an encoder whose output is unused gets grad None, as autograd gives the reference (CapBaselineModel)
                grads[e['prefix'] + 'obj_embed.weight'] = dWc[i * H:(i + 1) * H]
                grads[e['prefix'] + 'obj_embed.bias'] = dbc[i * H:(i + 1) * H]
        return grads


# =============================================================================================== EncoderVisual
class EncoderVisualBlock:
    """layer.py:46-61.  Tensor names: 'frames' + the module's parameter names (prefix-qualified) + 'pe'."""

    def __init__(self, prefix, baseline, p_drop, training):
        self.pf, self.baseline, self.p, self.training = prefix, baseline, p_drop, training

    @staticmethod
    def _packs(t, pf):
        wf, wr = t[pf + 'lstm.weight_ih_l0'], t[pf + 'lstm.weight_ih_l0_reverse']
        bs = [t[pf + 'lstm.bias_ih_l0'], t[pf + 'lstm.bias_hh_l0'], t[pf + 'lstm.bias_ih_l0_reverse'],
              t[pf + 'lstm.bias_hh_l0_reverse']]
        be = ops.backend()
        H4, H = wf.shape

        def build():
            Wih = op_empty((2 * H4,), H, wf)
            WC.cv(wf.detach(), Wih[:H4])
            WC.cv(wr.detach(), Wih[H4:])
            bsum = empty((2 * H4,), wf)
            WC.sum2(bs[0].detach(), bs[1].detach(), bsum[:H4])
            WC.sum2(bs[2].detach(), bs[3].detach(), bsum[H4:])
            return Wih, bsum
        return WC.packed(('evW', id(wf)), None, la.pver(wf, wr, *bs), build)

    def forward(self, t):
        be = ops.backend()
        pf = self.pf
        frames = _c(t['frames'])
        B, T, Din = frames.shape
        H = t[pf + 'linear_embed.weight'].shape[0]
        H4 = 4 * H
        training = self.training
        seed = next_seed()
        f2 = op(frames.view(B * T, Din))
        Xe = empty((B * T, H), frames, la.opdtype())
        be.gemm(f2, WC.get(t[pf + 'linear_embed.weight']), Xe, bias=t[pf + 'linear_embed.bias'])
        Wih, bsum = self._packs(t, pf)
        Gin = empty((B, T, 2 * H4), frames)
        be.gemm(Xe, Wih, Gin.view(B * T, 2 * H4), bias=bsum)
        lstm_out = empty((B, T, 2 * H), frames)
        # the two directions run concurrently: each recurrent GEMM is split over K to fill HALF of the SMs and the cell
        # kernel sums the partials itself (no reduce launch)
        whh = [WC.get(t[pf + 'lstm.weight_hh_l0']), WC.get(t[pf + 'lstm.weight_hh_l0_reverse'])]
        fused_step = FUSED_LSTM_STEP and whh[0].dtype == torch.bfloat16 and be.lstm_step_supported(B, H)
        Sg = 1 if fused_step else la.splitk_for(B, H4, H, sms=74)
        gates = zeros((Sg, 2, T, B, H4), frames)     # [0] ends up holding the activated gates (saved for BPTT)
        cs = zeros((2, T + 1, B, H), frames)
        hprev = op_zeros((2, B, T), H, frames)       # h fed INTO step t (operand dtype), clip-major like Gin / dGin

        def run_dir(d):
            order = list(range(T)) if d == 0 else list(range(T - 1, -1, -1))
            for k, tt in enumerate(order):
                if k > 0:
                    be.gemm(hprev[d, :, tt], whh[d], gates[:, d, tt] if Sg > 1 else gates[0, d, tt], splitk=Sg, b_static=True)
                nxt = order[k + 1] if k + 1 < T else None
                be.lstm_cell_fwd(gates[:, d, tt], cs[d, k], cs[d, k + 1], row_bias=Gin[:, tt, d * H4:(d + 1) * H4],
                                 h2=lstm_out[:, tt, d * H:(d + 1) * H], h3=(hprev[d, :, nxt] if nxt is not None else None))
        if fused_step:
            # ONE launch per time step for BOTH directions: recurrent product + cell per CTA owning 16 hidden units
            # (csrc/lstm_step.cu) instead of split-K tcgen05 GEMM + cell kernel per direction on two streams
            for k in range(T):
                tt = (k, T - 1 - k)
                nx = (k + 1, T - 2 - k) if k + 1 < T else None
                be.lstm_step_fwd(whh, [hprev[d, :, tt[d]] for d in (0, 1)] if k > 0 else None,
                                 [Gin[:, tt[d], d * H4:(d + 1) * H4] for d in (0, 1)], [cs[d, k] for d in (0, 1)],
                                 [cs[d, k + 1] for d in (0, 1)], [gates[0, d, tt[d]] for d in (0, 1)],
                                 [lstm_out[:, tt[d], d * H:(d + 1) * H] for d in (0, 1)],
                                 [hprev[d, :, nx[d]] for d in (0, 1)] if nx is not None else None)
        else:
            two_streams(frames, lambda: run_dir(0), lambda: run_dir(1))     # the two directions are independent chains
        Y = empty((B * T, 2 * H), frames)
        stY = empty((B * T, 2), frames)
        dY = site(self.p if training else 0.0, seed, 1)
        be.norm_fwd(lstm_out.view(B * T, 2 * H), t[pf + 'layernorm_lstm.weight'], t[pf + 'layernorm_lstm.bias'], y=Y,
                    stats=stY, drop=dY)
        sv = dict(t=t, dims=(B, T, Din, H), f2=f2, Xe=Xe, gates=gates[0], cs=cs, hprev=hprev, lstm_out=lstm_out, stY=stY, dY=dY)
        if self.baseline:
            out = la.mm(Y, WC.get(t[pf + 'out_try.weight']), bias=t[pf + 'out_try.bias'])
            sv.update(Y=Y)
            return [out.view(B, T, H)], sv
        D2 = 2 * H
        dpe = site(0.2 if training else 0.0, seed, 2)
        Ype = empty((B * T, D2), frames)
        pe = _c(t['pe'].detach()[0, :T])
        be.add_rowbcast(Y, pe, Ype, drop=dpe)
        wk, wq, wv = t[pf + 'self_attention.K.weight'], t[pf + 'self_attention.Q.weight'], t[pf + 'self_attention.V.weight']

        def build():
            W = op_empty((3 * D2,), D2, wk)
            for j, w in enumerate((wk, wq, wv)):
                WC.cv(w.detach(), W[j * D2:(j + 1) * D2])
            return W
        Wkqv = WC.packed(('evKQV', id(wk)), None, la.pver(wk, wq, wv), build)
        Ypo = op(Ype)
        KQV = empty((B, T, 3 * D2), frames, la.opdtype())          # K | Q | V projections as GEMM operands
        be.gemm(Ypo, Wkqv, KQV.view(B * T, 3 * D2))
        Kt, Qt, Vt = KQV[:, :, :D2], KQV[:, :, D2:2 * D2], KQV[:, :, 2 * D2:]
        lg = empty((B, T, T), frames)
        be.gemm(Kt, Qt, lg)
        Wt = empty((B, T, T), frames)
        sc = 1.0 / math.sqrt(D2)
        be.softmax_fwd(lg, Wt, dim=2, scale=sc)
        att_op = op_empty((B, T), D2, frames)
        be.gemm(op(Wt), Vt.transpose(1, 2), att_op)                  # V read transposed in place
        att_op = flat2(att_op)
        Z = la.mm(att_op, WC.get(t[pf + 'self_attention.output_layer.0.weight']))
        dz = site(self.p if training else 0.0, seed, 3)
        if dz is not None:
            be.dropout(Z, Z, dz)
        out = empty((B, T, H), frames)
        stZ = empty((B * T, 2), frames)
        be.norm_fwd(Z, t[pf + 'layernorm_sa.weight'], t[pf + 'layernorm_sa.bias'], y=out.view(B * T, H), stats=stZ)
        sv.update(dpe=dpe, Ypo=Ypo, Wkqv=Wkqv, KQV=KQV, lg=lg, Wt=Wt, sc=sc, att_op=att_op, Z=Z, dz=dz, stZ=stZ)
        return [out], sv

    def backward(self, sv, gouts):
        be = ops.backend()
        t, pf = sv['t'], self.pf
        B, T, Din, H = sv['dims']
        H4 = 4 * H
        grads = {}
        g = _c(gouts[0]).view(B * T, H)
        ref = g

        def lnp(name):
            w, b = t[pf + name + '.weight'], t[pf + name + '.bias']
            dw, db = small_zeros(w.shape, w), small_zeros(b.shape, b)
            grads[pf + name + '.weight'], grads[pf + name + '.bias'] = dw, db
            return w, b, dw, db
        dYv = empty((B * T, 2 * H), ref)
        if self.baseline:
            wo = t[pf + 'out_try.weight']
            gop = op(g)
            be.gemm(gop, WC.get(wo, transpose=True), dYv)
            grads[pf + 'out_try.weight'] = la.mm(g.t(), sv['Y'].t())
            dbo = small_zeros((H,), ref)
            be.colsum(g, dbo)
            grads[pf + 'out_try.bias'] = dbo
        else:
            D2 = 2 * H
            w, b, dw, db = lnp('layernorm_sa')
            dZ = empty((B * T, H), ref)
            be.norm_bwd(g, sv['Z'], w, b, sv['stZ'], dx=dZ, dgamma=dw, dbeta=db)
            if sv['dz'] is not None:
                be.dropout(dZ, dZ, sv['dz'])
            wo = t[pf + 'self_attention.output_layer.0.weight']
            dZop = op(dZ)
            datt = empty((B, T, D2), ref, la.opdtype())
            be.gemm(dZop, WC.get(wo, transpose=True), datt.view(B * T, D2))
            grads[pf + 'self_attention.output_layer.0.weight'] = la.mm(dZop.t(), sv['att_op'].t())
            KQV = sv['KQV']
            Kt, Qt, Vt = KQV[:, :, :D2], KQV[:, :, D2:2 * D2], KQV[:, :, 2 * D2:]
            dKQV = empty((B, T, 3 * D2), ref, la.opdtype())
            dK, dQ, dV = dKQV[:, :, :D2], dKQV[:, :, D2:2 * D2], dKQV[:, :, 2 * D2:]
            dW = empty((B, T, T), ref)
            datt_op, Wt_op = datt, op(sv['Wt'])
            be.gemm(datt_op, Vt, dW)                                          # dW[i,j] = datt_i . V_j
            be.gemm(Wt_op.transpose(1, 2), datt_op.transpose(1, 2), dV)       # dV_j = sum_i W[i,j] datt_i
            dlg = empty((B, T, T), ref)
            be.softmax_bwd(sv['lg'], dW, dlg, dim=2, scale=sv['sc'])
            dlg_op = op(dlg)
            be.gemm(dlg_op, Qt.transpose(1, 2), dK)                           # dK_i = sum_j dlg[i,j] Q_j
            be.gemm(dlg_op.transpose(1, 2), Kt.transpose(1, 2), dQ)           # dQ_j = sum_i dlg[i,j] K_i
            dKQVop = dKQV.view(B * T, 3 * D2)
            wk, wq, wv = t[pf + 'self_attention.K.weight'], t[pf + 'self_attention.Q.weight'], t[pf + 'self_attention.V.weight']
            dYpe = empty((B * T, D2), ref)
            # one GEMM with K = 3*D2 over the packed [K;Q;V] weight read transposed in place (no accumulate passes)
            be.gemm(dKQVop, sv['Wkqv'].t(), dYpe)
            dWkqv = la.mm(dKQVop.t(), sv['Ypo'].t())                # (3*D2, D2)
            for j, n in enumerate(('K', 'Q', 'V')):
                grads[pf + 'self_attention.%s.weight' % n] = dWkqv[j * D2:(j + 1) * D2]
            if sv['dpe'] is not None:
                be.dropout(dYpe, dYpe, sv['dpe'])
            dYv = dYpe
        w, b, dw, db = lnp('layernorm_lstm')
        dL = empty((B, T, 2 * H), ref)
        be.norm_bwd(dYv, sv['lstm_out'].view(B * T, 2 * H), w, b, sv['stY'], dx=dL.view(B * T, 2 * H), dgamma=dw, dbeta=db,
                    drop=sv['dY'])
        early_sync(grads)            # (data parallel) the attention / LayerNorm gradients travel during the BPTT below
        # BiLSTM BPTT
        dGin = empty((B, T, 2 * H4), ref, la.opdtype())
        gates, cs, hprev = sv['gates'], sv['cs'], sv['hprev']
        whhT = [WC.get(t[pf + 'lstm.weight_hh_l0'], transpose=True), WC.get(t[pf + 'lstm.weight_hh_l0_reverse'], transpose=True)]
        names_hh = ['lstm.weight_hh_l0', 'lstm.weight_hh_l0_reverse']
        # everything that crosses the stream fork/join is allocated here, on the main stream
        Sd = la.splitk_for(B, H, H4, sms=74)             # recurrent data-gradient GEMM: partials summed by the cell kernel
        bufs = [(zeros((Sd, B, H), ref), zeros((B, H), ref), empty((B, H), ref)) for _ in range(2)]

        def run_dir_bwd(d):
            order = list(range(T)) if d == 0 else list(range(T - 1, -1, -1))
            dhrec, dc, dc2 = bufs[d]
            for k in range(T - 1, -1, -1):
                tt = order[k]
                dg2 = dGin[:, tt, d * H4:(d + 1) * H4]
                be.lstm_cell_bwd(gates[d, tt], cs[d, k], cs[d, k + 1], dL[:, tt, d * H:(d + 1) * H], dc, dc2,
                                 dgates2=dg2, dh2=dhrec)
                dc, dc2 = dc2, dc
                if k > 0:
                    be.gemm(op(dg2), whhT[d], dhrec if Sd > 1 else dhrec[0], splitk=Sd, b_static=True)
        two_streams(ref, lambda: run_dir_bwd(0), lambda: run_dir_bwd(1))
        dGin2 = dGin.view(B * T, 2 * H4)
        for d in range(2):
            # time-batched recurrent weight gradient: dW_hh = dgates^T . h_prev, both operands read transposed in place
            hp = hprev[d]
            hp2 = hp.as_strided((B * T, H), (hp.stride(1), 1))
            grads[pf + names_hh[d]] = la.mm(dGin2[:, d * H4:(d + 1) * H4].t(), hp2.t())
        Wih, _ = self._packs(t, pf)
        dbg = small_zeros((2 * H4,), ref)
        be.colsum(dGin2, dbg)
        grads[pf + 'lstm.bias_ih_l0'] = dbg[:H4]
        grads[pf + 'lstm.bias_hh_l0'] = dbg[:H4]
        grads[pf + 'lstm.bias_ih_l0_reverse'] = dbg[H4:]
        grads[pf + 'lstm.bias_hh_l0_reverse'] = dbg[H4:]
        dWih = la.mm(dGin2.t(), sv['Xe'].t())
        grads[pf + 'lstm.weight_ih_l0'] = dWih[:H4]
        grads[pf + 'lstm.weight_ih_l0_reverse'] = dWih[H4:]
        dXe = empty((B * T, H), ref, la.opdtype())
        be.gemm(op(dGin2), op(Wih.t()), dXe)
        grads[pf + 'linear_embed.weight'] = la.mm(dXe.t(), sv['f2'].t())
        dbl = small_zeros((H,), ref)
        be.colsum(dXe, dbl)
        grads[pf + 'linear_embed.bias'] = dbl
        return grads
