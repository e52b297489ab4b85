"""owl -- Python 3 re-binding of Minerva's owl API over the B200 kernel library.

Mirrors the reference module surface (owl/owl/__init__.py:33-189): device management,
zeros / ones / randn / randb / from_numpy / concat / slice, and the NArray type with its
operators.  Differences, all behind the same names:
  * arrays live in HBM and every operator is one enqueue-only call into the C ABI of
    include/mnv.h on the current device's stream (the reference builds a lazy DAG and blocks the
    worker thread on cudaStreamSynchronize after every op, minerva/device/device.cpp:214-222);
    `wait_for_all()` / `to_numpy()` are still the only blocking points;
  * there is no CPU execution path: `create_cpu_device()` is accepted for script compatibility but
    computing on it raises (north_star: CPU fallback removed);
  * torch is used only for device memory (caching allocator == the reference's PooledDataStore
    role), streams and torch.distributed.
"""
import numpy as np

from . import _runtime as _rt
from ._runtime import (create_cpu_device, create_gpu_device, get_gpu_device_count, has_cuda,  # noqa: F401
                       set_device, wait_for_all, current_device)
from .narray import NArray


def zeros(shape):
    return NArray.zeros(shape)


def ones(shape):
    return NArray.ones(shape)


def _uninit(shape):
    """Uninitialised array for outputs an op overwrites completely (not part of the reference surface)."""
    return NArray._new(list(shape), current_device())


def randn(shape, mu, var):
    """N(mu, var^2): `var` is used as the standard deviation, as on both reference paths
    (minerva/op/impl/basic.cpp:275, cuda_perform.cu:621)."""
    return NArray.randn(shape, mu, var)


def randb(shape, prob):
    return NArray.randb(shape, prob)


def from_numpy(nparr):
    """numpy (C order) -> NArray with the shape REVERSED (owl/owl/libowl.pyx:425-431)."""
    return NArray.from_numpy(np.require(nparr, dtype=np.float32, requirements=["C"]))


def concat(narrays, concat_dim):
    return NArray.concat(narrays, concat_dim)


def slice(src, slice_dim, st_off, slice_count):  # noqa: A001 (reference name)
    return NArray.slice(src, slice_dim, st_off, slice_count)


def set_seed(seed):
    """Generators are keyed by (seed, call counter) instead of the wall clock
    (minerva/op/impl/cuda.cpp:601,606) so runs are reproducible."""
    _rt.set_seed(seed)


def set_rank_salt(rank):
    """Mix the data-parallel rank into the dropout (randb) seeds; weights (randn) stay identical across ranks."""
    _rt.set_rank_salt(rank)


def set_exact_math(flag):
    """True: exp / ln / sigmoid / tanh use the exact-mode kernels (mnv_*_exact), bit-identical to the reference's CPU ops
    (glibc's expf / logf / tanhf restated in the kernel); False (default): CUDA's fast functions, <= 4 ulp away."""
    NArray.exact_math = bool(flag)
