"""Stand-ins for the torch.cuda stream / event / graph objects - TEST INFRASTRUCTURE ONLY.

`install()` replaces the handful of torch.cuda entry points that bench.py, dlsg.graphs and dlsg.gan call with no-op
objects, so that their CONTROL FLOW (capture set-up, replays, prefetch pipeline, result line) can be driven on a machine
without a GPU together with the CPU emulation of the kernels (tests/cpu_emul.py).  A "captured" graph simply keeps the
tensors its body produced at capture time; replay() does nothing.  Nothing here is imported by the product: bench.py only
reaches it through DLSG_BENCH_EMUL=1, which exists for tests/test_bench_flow_cpu.py.
"""
import contextlib
import time

import torch


class Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


class Event:
    def __init__(self, enable_timing=False, **k):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)

    def synchronize(self):
        pass


class CUDAGraph:
    def replay(self):
        pass


@contextlib.contextmanager
def _ctx(*a, **k):
    yield


class _Profiler:
    @staticmethod
    def start():
        pass

    @staticmethod
    def stop():
        pass


def install():
    _cur = Stream()
    torch.cuda.Stream = Stream
    torch.cuda.Event = Event
    torch.cuda.CUDAGraph = CUDAGraph
    torch.cuda.graph = _ctx
    torch.cuda.stream = _ctx
    torch.cuda.current_stream = lambda *a, **k: _cur
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.profiler = _Profiler
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.Tensor.record_stream = lambda self, s: None
