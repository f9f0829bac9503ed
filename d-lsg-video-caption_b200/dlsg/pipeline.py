"""Input pipeline of the captured training step (reference: utils/data.py:55-63 yields fp32 features, the
trainer copies each batch to the device inside the iteration, run_gun.py:156-158).

At B=64 a batch of MSR-VTT-shaped features is 515 MB in fp32: over PCIe Gen5 (~55 GB/s) that is 9.3 ms, more than the
6.5 ms the step computes - the step would be transfer-bound.  `FeaturePipe` removes that bound without changing the
arithmetic of the bf16 step:

* host format: the features are kept (or staged once by `stage()`) in bf16 in PINNED host memory - the same rounding the
  bf16 step applies to them first thing on the device (`cast_f32_bf16_flat`), so results are bit-identical to feeding
  fp32; half the PCIe bytes (257 MB/step) and no cast kernel;
* the host->device copy of batch k+1 runs on a copy stream while step k computes (events, no host sync): `put()`
  returns immediately, `run()` hands the staged batch to the step (device-to-device into the captured graph's static
  inputs) and releases the staging buffers for the next `put()`.

Works with any step object that has `.load(frames, regions, captions)` and `__call__() -> loss tensor`
(`graphs.GraphedTrainStep`, `gan.GanIteration`).  On CPU tensors (tests) the copies are synchronous.
"""
import torch


def stage(frames, regions, captions, dtype=torch.bfloat16, pin=True):
    """Loader batch (fp32 host tensors) -> the pipeline's host format: bf16 features, int64 captions, pinned."""
    def conv(t, dt):
        t = t.to(dt) if t.dtype != dt else t
        t = t.contiguous()
        return t.pin_memory() if (pin and torch.cuda.is_available() and not t.is_pinned()) else t
    return conv(frames, dtype), conv(regions, dtype), conv(captions, torch.int64)


class FeaturePipe:
    def __init__(self, step, frames_like, regions_like, captions_like, device):
        """`*_like`: one host batch in the pipeline's host format (shapes / dtypes of every later batch)."""
        self.step = step
        self.device = torch.device(device)
        self.cuda = self.device.type == 'cuda'
        self.bufs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in (frames_like, regions_like, captions_like)]
        self.bytes_per_batch = sum(t.numel() * t.element_size() for t in (frames_like, regions_like, captions_like))
        self.pending = False
        self._lens = None
        if self.cuda:
            self.copy_stream = torch.cuda.Stream(device=self.device)
            self.ready, self.consumed = torch.cuda.Event(), torch.cuda.Event()
            self.consumed.record(torch.cuda.current_stream(self.device))

    def put(self, frames, regions, captions, cap_lens=None):
        """Start the host->device copy of the next batch (pinned host tensors); returns immediately.  `cap_lens` (the
        loader's per-clip caption lengths, run_gun.py:155) is handed to the step's `load` with the batch."""
        if self.pending:
            raise RuntimeError('FeaturePipe.put: the previous batch has not been consumed by run() yet')
        src = (frames, regions, captions)
        for s, d in zip(src, self.bufs):
            if s.shape != d.shape or s.dtype != d.dtype:
                raise ValueError('FeaturePipe.put: batch %s %s does not match the staged format %s %s'
                                 % (tuple(s.shape), s.dtype, tuple(d.shape), d.dtype))
        if self.cuda:
            self.copy_stream.wait_event(self.consumed)          # the step has taken the previous batch out of the buffers
            with torch.cuda.stream(self.copy_stream):
                for s, d in zip(src, self.bufs):
                    d.copy_(s, non_blocking=True)
                self.ready.record(self.copy_stream)
        else:
            for s, d in zip(src, self.bufs):
                d.copy_(s)
        self._lens = cap_lens
        self.pending = True

    def _load(self):
        if self._lens is not None:
            self.step.load(*self.bufs, cap_lens=self._lens)
        else:
            self.step.load(*self.bufs)

    def run(self):
        """Run the step on the staged batch; returns the step's result (a device tensor: no host sync here)."""
        if not self.pending:
            raise RuntimeError('FeaturePipe.run: no staged batch (call put() first)')
        if self.cuda:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self.ready)
            self._load()                                        # device-to-device into the graph's static inputs
            self.consumed.record(cur)
        else:
            self._load()
        self.pending = False
        return self.step()
