"""TEST-ONLY stand-in for the few torch calls bench.py's GPU arm makes (device buffers, streams, events), so that the
arm's Python can be exercised end to end on the CPU emulator of the C-ABI library, where device memory is host
memory (tests/test_bench_contract.py::test_gpu_arm_runs_on_the_emulator).  Never on the product's import path."""
import time

import numpy as np

uint8, int64, float64 = np.uint8, np.int64, np.float64


class _T:
    def __init__(self, a):
        self.a = a

    def data_ptr(self):
        return self.a.ctypes.data

    def numel(self):
        return int(self.a.size)

    def to(self, _dev):
        return self

    def cpu(self):
        return self

    def numpy(self):
        return self.a

    def __getitem__(self, k):
        return _T(self.a[k])

    def copy_(self, other):
        np.copyto(self.a, other.a)
        return self


def from_numpy(a):
    return _T(np.ascontiguousarray(a))


def empty(n, dtype=uint8, device=None, pin_memory=False):
    return _T(np.zeros(n, dtype=dtype))


def device(kind, index=0):
    return (kind, index)


class _Stream:
    cuda_stream = 0


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return 1e3 * (other.t - self.t)


class cuda:                      # noqa: N801
    Event = _Event

    @staticmethod
    def is_available():
        return True

    @staticmethod
    def set_device(_i):
        pass

    @staticmethod
    def synchronize():
        pass

    @staticmethod
    def current_stream():
        return _Stream()
