"""Adam on our own multi-tensor kernel (dlsg_adam_multi), driving the state of the caller's torch.optim.Adam in place
(`exp_avg`, `exp_avg_sq`, `step` stay where torch keeps them, so checkpoints - run_gun.py:302-310 - and a later
`optimizer.step()` by torch remain valid).  Same arithmetic as torch.optim.Adam without weight decay / amsgrad / maximize
(run_gun.py:91,100: lr 1.6e-4, betas (0.5, 0.9)).

Why: SURVEY 8f-3.  The update is one launch per parameter block instead of six multi_tensor_apply launches, the bf16
GEMM-operand copy of every weight (dlsg.linalg.PinnedWeights) is written in the same pass instead of a separate
conversion, and a block can be updated on a side stream as soon as its gradients exist (the decoder's 76 M parameters
while the encoder backward is still running).
"""
import torch

from . import ops


def supported(opt):
    if type(opt) is not torch.optim.Adam:
        return False
    for g in opt.param_groups:
        if g.get('weight_decay', 0) != 0 or g.get('amsgrad', False) or g.get('maximize', False):
            return False
        if torch.is_tensor(g['lr']):
            return False
    return True


def _ensure_state(opt, p):
    st = opt.state[p]
    if len(st) == 0:
        st['step'] = torch.zeros((), dtype=torch.float32, device=p.device)          # what torch creates for fused / capturable Adam
        st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
    assert torch.is_tensor(st['step']) and st['step'].is_cuda == p.is_cuda, 'Adam state must be device-resident (fused / capturable)'
    return st


def _as2d(t):
    return t.view(1, -1) if t.dim() != 2 else t


class AdamDriver:
    def __init__(self, opt, pinned=None):
        assert supported(opt), 'AdamDriver: plain torch.optim.Adam (no weight decay / amsgrad / maximize, float lr) only'
        self.opt = opt
        self.group_of = {}
        for g in opt.param_groups:
            for p in g['params']:
                self.group_of[id(p)] = g
        # bf16 shadows: parameter -> list of (row0, col0, rows, cols, dst view) taken from the pinned conversion list when
        # the (single-source) conversions of a parameter tile it exactly; everything else stays with PinnedWeights.refresh
        self.shadow = {}
        self.residual = None
        if pinned is not None:
            self._map_shadows(pinned)
        # learning rates live in device scalars created HERE (outside any capture) so that a captured step follows them
        self.lr_dev = {id(g): torch.full((), float(g['lr']), dtype=torch.float32, device=g['params'][0].device)
                       for g in opt.param_groups}
        self._lr_seen = {id(g): float(g['lr']) for g in opt.param_groups}
        self._plans = []

    def _map_shadows(self, pinned):
        params = [p for g in self.opt.param_groups for p in g['params']]
        spans = sorted(((p.data_ptr(), p.data_ptr() + p.numel() * 4, p) for p in params), key=lambda x: x[0])
        per, rest = {}, []
        for src, src2, dst in pinned.pairs:
            owner = None
            if src2 is None and dst.dtype == torch.bfloat16:
                a = src.data_ptr()
                for lo, hi, p in spans:
                    if lo <= a < hi:
                        owner = p
                        break
            if owner is None or owner.dim() != 2:
                rest.append((src, src2, dst))
                continue
            cols = owner.shape[1]
            off = (src.data_ptr() - owner.data_ptr()) // 4
            r0, c0 = off // cols, off % cols
            if src.shape[0] > 1 and src.stride(0) != cols:
                rest.append((src, src2, dst))
                continue
            per.setdefault(id(owner), (owner, []))[1].append((r0, c0, src.shape[0], src.shape[1], dst, (src, src2, dst)))
        for pid, (owner, segs) in per.items():
            if sum(s_[2] * s_[3] for s_ in segs) == owner.numel():
                self.shadow[pid] = [s_[:5] for s_ in segs]
            else:                                       # partial / duplicated coverage: leave to the residual refresh
                rest.extend(s_[5] for s_ in segs)
        self.residual = ops.backend().make_convert_plan(rest) if rest else None

    def step(self, params, grad_of=None):
        """Update `params` (those with a gradient) with ONE kernel launch per (param group) + one step-counter increment.
        Must run after their gradients are final; capturable (the segment table rides in the kernel parameters).
        grad_of: callable(param) -> gradient tensor in the parameter's shape (fp32 or bf16, e.g. GradSync.grad_of: the
        all-reduced bucket view is read in place); default p.grad."""
        be = ops.backend()
        by_group = {}
        for p in params:
            if p.grad is None:
                continue
            by_group.setdefault(id(self.group_of[id(p)]), (self.group_of[id(p)], []))[1].append(p)
        for _, (g, ps) in by_group.items():
            segs, steps = [], []
            for p in ps:
                st = _ensure_state(self.opt, p)
                steps.append(st['step'])
                grad = p.grad if grad_of is None else grad_of(p)
                if grad is None:
                    raise RuntimeError('AdamDriver: no reduced gradient for a parameter of shape %s' % (tuple(p.shape),))
                grad = grad if grad.is_contiguous() else grad.contiguous()
                m, v = st['exp_avg'], st['exp_avg_sq']
                sh = self.shadow.get(id(p))
                if sh is None:
                    segs.append(dict(p=_as2d(p.data), g=_as2d(grad), m=_as2d(m), v=_as2d(v), dst=None, step=st['step']))
                else:
                    for r0, c0, nr, nc, dst in sh:
                        sl = (slice(r0, r0 + nr), slice(c0, c0 + nc))
                        segs.append(dict(p=p.data[sl], g=grad[sl], m=m[sl], v=v[sl], dst=dst, step=st['step']))
            torch._foreach_add_(steps, 1)
            plan = be.make_adam_plan(segs)
            b1, b2 = g['betas']
            be.adam_multi(plan, steps[0], float(g['lr']), b1, b2, g['eps'], lr_dev=self.lr_dev[id(g)])
            self._plans = self._plans[-7:] + [plan]      # keep the tables of the latest launches alive (graph replays)

    def refresh_residual(self):
        if self.residual is not None:
            ops.backend().multi_convert(self.residual)

    def sync_lr(self):
        """Copy the groups' current learning rates (a scheduler may have changed them) into the device scalars."""
        for g in self.opt.param_groups:
            if float(g['lr']) != self._lr_seen.get(id(g)):
                self.lr_dev[id(g)].fill_(float(g['lr']))
                self._lr_seen[id(g)] = float(g['lr'])
