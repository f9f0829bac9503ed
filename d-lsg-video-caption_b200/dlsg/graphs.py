"""CUDA-graph capture of a whole D-LSG training step (forward + fused masked CE + backward [+ gradient all-reduce]
+ Adam).  A step launches ~1000 small kernels; replaying them as one graph removes the host launch path.

Every libdlsg entry is capturable (no allocation / sync inside); the bf16 weight-copy cache is put in `force` mode
during capture so that the refresh kernels are part of the graph (weights change every replay).  Static input
buffers are filled with (non-blocking) copies before each replay.  Limitations of a captured step, by construction:
the teacher-forcing coin flips (layer.py:432) and the dropout seeds are frozen at capture time - use the eager path
(model(...), loss.backward(), opt.step()) when scheduled sampling with ratio < 1 must be re-drawn every step.
"""
import torch

from . import functional as DF
from . import losses


class GraphedTrainStep:
    def __init__(self, model, optimizer, frames, regions, captions, cap_lens, max_words=26, tf_ratio=1.0,
                 process_group=None, warmup=3):
        dev = frames.device
        self.model, self.opt, self.pg = model, optimizer, process_group
        self.frames, self.regions, self.captions = frames.clone(), regions.clone(), captions.clone()
        self.lens = torch.as_tensor(list(cap_lens), dtype=torch.int32, device=dev)
        self.inv = torch.tensor([1.0 / max(1, int(sum(cap_lens)))], dtype=torch.float32, device=dev)
        self.max_words, self.tf = max_words, tf_ratio
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.world = dist.get_world_size(process_group)
            self.sync = DF.GradSync(process_group)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        DF.WC.force = True
        try:
            with torch.cuda.graph(self.graph):
                self.loss = self._body()
        finally:
            DF.WC.force = False
        torch.cuda.synchronize()

    def _body(self):
        self.opt.zero_grad(set_to_none=True)
        out = self.model(self.frames, self.regions, self.captions, self.max_words, self.tf)[0]
        loss = losses.packed_cross_entropy(out, self.captions, self.lens, self.inv)
        if self.world > 1:
            # per-block flat buckets, all-reduced on a side stream as soon as each block's backward is done
            DF.GRAD_SYNC = self.sync
        try:
            loss.backward()
        finally:
            DF.GRAD_SYNC = None
        if self.world > 1:
            self.sync.wait()
        self.opt.step()
        return loss.detach()

    def load(self, frames, regions, captions, cap_lens=None):
        """Copy a new batch (host pinned or device tensors) into the static buffers."""
        self.frames.copy_(frames, non_blocking=True)
        self.regions.copy_(regions, non_blocking=True)
        self.captions.copy_(captions, non_blocking=True)
        if cap_lens is not None:
            self.lens.copy_(torch.as_tensor(list(cap_lens), dtype=torch.int32), non_blocking=False)
            self.inv.fill_(1.0 / max(1, int(sum(cap_lens))))

    def __call__(self):
        self.graph.replay()
        return self.loss
