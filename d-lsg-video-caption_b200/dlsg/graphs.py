"""CUDA-graph capture of a whole D-LSG training step (forward + fused masked CE + backward [+ gradient all-reduce]
+ Adam).  A step launches ~1000 small kernels; replaying them as one graph removes the host launch path.

Every libdlsg entry is capturable (no allocation / sync inside); the bf16 weight-copy cache is put in `force` mode
during capture so that the refresh kernels are part of the graph (weights change every replay).  Static input
buffers are filled with (non-blocking) copies before each replay.  Limitations of a captured step, by construction:
the teacher-forcing coin flips (layer.py:432) and the dropout seeds are frozen at capture time - use the eager path
(model(...), loss.backward(), opt.step()) when scheduled sampling with ratio < 1 must be re-drawn every step.
"""
import os

import torch

from . import functional as DF
from . import losses
from . import optim


def _inv_count(cap_lens, max_words):
    """1 / number of packed tokens: the reference slices outputs[j][:cap_lens[j]] of a (max_words)-long row
    (run_gun.py:189-197), so a caption longer than max_words contributes max_words tokens."""
    return 1.0 / max(1, sum(min(int(c), max_words) for c in cap_lens))


def check_capturable(opt):
    """A torch optimizer stepped inside a capture must keep its step counters on the device."""
    if isinstance(opt, torch.optim.Adam) and any(p.is_cuda for g in opt.param_groups for p in g['params']):
        for g in opt.param_groups:
            if not (g.get('capturable', False) or optim.supported(opt)):
                raise ValueError('optimizer must be torch.optim.Adam(..., capturable=True) to be captured in a CUDA graph')
        for st in opt.state.values():
            if 'step' in st and not (torch.is_tensor(st['step']) and st['step'].is_cuda):
                raise ValueError('optimizer state holds host-side step counters (created with capturable=False): build the optimizer '
                                 'with capturable=True (or fused=True) before its first step')


def ensure_adam_state(opt, params):
    if isinstance(opt, torch.optim.Adam):
        for p in params:
            if p.requires_grad:
                optim._ensure_state(opt, p)


def snapshot(model_or_models, opts):
    """Parameters, optimizer state, host / device RNG state and the dropout seed counter, to undo capture warm-up steps."""
    import copy
    import random
    models = model_or_models if isinstance(model_or_models, (list, tuple)) else [model_or_models]
    return {'params': [(p, p.detach().clone()) for m in models for p in m.parameters()],
            'opts': [(o, copy.deepcopy(o.state_dict())) for o in opts],
            'py': random.getstate(), 'torch': torch.get_rng_state(),
            'cuda': torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
            'seed_counter': copy.copy(DF._seed_counter)}


def restore(snap):
    import random
    with torch.no_grad():
        for p, v in snap['params']:
            p.copy_(v)
    for o, sd in snap['opts']:
        o.load_state_dict(sd)
    random.setstate(snap['py'])
    torch.set_rng_state(snap['torch'])
    if snap['cuda'] is not None:
        torch.cuda.set_rng_state(snap['cuda'])
    DF._seed_counter = snap['seed_counter']
    DF.WC.gen += 1


class GraphedTrainStep:
    def __init__(self, model, optimizer, frames, regions, captions, cap_lens, max_words=26, tf_ratio=1.0,
                 process_group=None, warmup=3, pin_weights=True, own_adam=True):
        dev = frames.device
        self.model, self.opt, self.pg = model, optimizer, process_group
        self.frames, self.regions, self.captions = frames.clone(), regions.clone(), captions.clone()
        self.lens = torch.as_tensor(list(cap_lens), dtype=torch.int32, device=dev)
        self.max_words, self.tf = max_words, tf_ratio
        self.inv = torch.tensor([_inv_count(cap_lens, max_words)], dtype=torch.float32, device=dev)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = 1
        self.pinned = None
        self.adam = None
        if process_group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.world = dist.get_world_size(process_group)
            self.sync = DF.GradSync(process_group)
        elif os.environ.get('DLSG_FORCE_SYNC') == '1':
            # measurement aid (one GPU): the data-parallel stream / bucket structure without a communicator
            # (use with DLSG_SYNC_DEBUG=nonccl | nopack)
            self.world = 2
            self.sync = DF.GradSync(None)
        check_capturable(optimizer)
        snap = snapshot(model, [optimizer]) if warmup > 0 else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if snap is not None:
            restore(snap)                         # the warm-up steps (allocator / NCCL warm-up) leave no trace in weights or Adam state
        ensure_adam_state(optimizer, self.params) # state tensors must exist BEFORE the capture (a zero-fill captured with them would reset them on every replay)
        self.pinned = self._record_weight_copies() if pin_weights else None
        self.adam = None
        if own_adam and self.pinned is not None and optim.supported(optimizer):
            # our multi-tensor Adam drives the optimizer's state in place and emits the bf16 operand copies itself
            self.adam = optim.AdamDriver(optimizer, self.pinned)
            self.pinned.refresh()                # copies are current before the first replay; Adam keeps them current
            self.dec_params = [p for n, p in model.named_parameters() if n.startswith('decoder.') and p.requires_grad]
            self.side = torch.cuda.Stream()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        DF.WC.force = True
        from . import ops
        l0 = ops.backend().launches
        try:
            with torch.cuda.graph(self.graph):
                self.loss = self._body()
        finally:
            DF.WC.force = False
        self.launches = ops.backend().launches - l0          # libdlsg kernels recorded in one replay
        self._adam_plans = list(self.adam._plans) if self.adam is not None else None     # (keeps the tensors the captured launches point at alive)
        torch.cuda.synchronize()

    def _side(self):
        """Stream of the overlapped decoder update: our own side stream, or - with several ranks - the gradient-sync stream
        (the update is queued right behind the decoder bucket's all-reduce)."""
        if self.world > 1 and self.sync.side is not None:
            return self.sync.side
        return self.side

    def _record_weight_copies(self):
        """One dry forward (nothing is updated; the host RNG / dropout seed counter are restored) that logs every weight
        conversion of a training forward, so the captured step refreshes ALL bf16 operand copies with ONE multi-segment
        launch (dlsg_multi_convert) instead of ~90 small conversion kernels."""
        import copy
        import random
        rng = random.getstate()
        counter = copy.copy(DF._seed_counter)
        try:
            with torch.enable_grad():
                pinned = DF.WC.record(lambda: self.model(self.frames, self.regions, self.captions, self.max_words, self.tf))
        finally:
            random.setstate(rng)
            DF._seed_counter = counter
        return pinned

    def _body(self):
        self.opt.zero_grad(set_to_none=True)
        grad_of = None
        if self.world > 1:
            self.sync.begin_step()
            grad_of = self.sync.grad_of           # Adam reads the averaged gradients in place from the reduced buckets
        if self.pinned is not None:
            if self.adam is None:
                self.pinned.refresh()             # current fp32 masters -> every operand copy, one launch
            DF.WC.pin(self.pinned)
        done = []
        if self.adam is not None:
            # Every block's parameters are updated on a side stream the moment the NEXT block starts its backward: all their
            # gradients are final by then (the decoder's 76 M parameters - 65 % of the model - while the encoder backward is
            # still computing), and the bandwidth-bound update hides behind the latency-bound backward.  With several ranks the
            # side stream is the gradient-sync stream: the update is queued right behind that block's bucket all-reduce and
            # reads the averaged gradients in place.
            side = self._side()
            seen = set()

            def hook(block):
                ps = [p for p in self.params if p.grad is not None and id(p) not in seen]
                if ps:
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        self.adam.step(ps, grad_of)
                    done.extend(ps)
                    seen.update(id(p) for p in ps)
            DF.BLOCK_BWD_HOOK = hook
        try:
            out = self.model(self.frames, self.regions, self.captions, self.max_words, self.tf)[0]
            loss = losses.packed_cross_entropy(out, self.captions, self.lens, self.inv, unit_grad=True)
            if self.world > 1:
                # per-block flat buckets, all-reduced on a side stream as soon as each block's backward is done
                DF.GRAD_SYNC = self.sync
            try:
                loss.backward()
            finally:
                DF.GRAD_SYNC = None
                DF.BLOCK_BWD_HOOK = None
        finally:
            DF.WC.unpin()
        if self.world > 1:
            self.sync.wait()
        if self.adam is None:
            if self.world > 1:
                self.sync.write_back(self.params)     # a foreign optimizer reads p.grad
            self.opt.step()
        else:
            seen = set(id(p) for p in done)
            self.adam.step([p for p in self.params if p.grad is not None and id(p) not in seen], grad_of)
            if done:
                torch.cuda.current_stream().wait_stream(self._side())
            self.adam.refresh_residual()          # the few copies Adam cannot emit itself (summed bias pairs)
        return loss.detach()

    def load(self, frames, regions, captions, cap_lens=None):
        """Copy a new batch (host pinned or device tensors) into the static buffers."""
        self.frames.copy_(frames, non_blocking=True)
        self.regions.copy_(regions, non_blocking=True)
        self.captions.copy_(captions, non_blocking=True)
        if cap_lens is not None:
            self.lens.copy_(torch.as_tensor(list(cap_lens), dtype=torch.int32), non_blocking=False)
            self.inv.fill_(_inv_count(cap_lens, self.max_words))

    def __call__(self):
        if self.adam is not None:
            self.adam.sync_lr()                   # a scheduler may have changed the learning rate since the last replay
        self.graph.replay()
        DF.WC.gen += 1        # the replay changed the parameters through raw pointers: no eval-scope bf16 copy survives it
        return self.loss

    def refresh_weights(self):
        """Call after changing the parameters from outside (load_state_dict, manual edits): re-derives every bf16 operand
        copy from the fp32 masters (otherwise only the in-graph Adam keeps them current)."""
        if self.pinned is not None:
            self.pinned.refresh()


class GraphedDecode:
    """CUDA-graph capture of CapGnnModel inference (evaluate.py:68: net(frames, regions, None)) for a fixed batch shape:
    encoder + all decode steps (greedy: 26 steps; beam: 26 batched beam steps) replay as ONE graph, so the ~450-570
    kernel launches per batch no longer pay the host launch path.  The beam search's only host interaction (early-stop
    length, one D2H read) happens after the replay.  bf16 weight copies are taken from the cache at capture time:
    re-capture after the weights change."""

    def __init__(self, net, frames, regions, beam_size, warmup=2):
        from . import decoder as DD
        self.DD, self.net, self.beam = DD, net, beam_size
        self.frames, self.regions = frames.clone(), regions.clone()
        dec = net.decoder
        self.T = dec.max_words
        self.end = dec.vocab('<end>')
        net.update_beam_size(beam_size)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(warmup):
                self._core()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.trace = self._core()
        torch.cuda.synchronize()

    def _core(self):
        net, DD = self.net, self.DD
        obj, mot = net.encoder(self.frames, self.regions)
        t = {k: v.detach() for k, v in net.decoder._used().items()}
        if self.beam == 1:
            return DD.decode_greedy(t, '', True, obj, mot, self.T)
        return DD.decode_beam_core(t, '', True, obj, mot, self.T, self.beam, self.end, net.decoder.beam_search.per_node_beam_size)

    def __call__(self, frames=None, regions=None):
        if frames is not None:
            self.frames.copy_(frames, non_blocking=True)
            self.regions.copy_(regions, non_blocking=True)
        self.graph.replay()
        if self.beam == 1:
            return self.trace
        return self.DD.decode_beam_finish(*self.trace, self.end)[0]
