"""Helpers for the -m gpu parity tests: torch owns device memory and the stream (plumbing); every
computation goes through the C ABI of include/mnv.h via minerva_b200._lib (ctypes)."""
import numpy as np
import torch

from minerva_b200 import _lib


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def empty(n):
    return torch.empty(int(n), dtype=torch.float32, device="cuda")


def host(t):
    return t.detach().cpu().numpy()


def stream():
    return torch.cuda.current_stream().cuda_stream


# Alternate kernel paths are forced through the TUNING build (include/mnv_debug.h); while any option is off its
# default, run() routes calls there, otherwise through the product library.
_OPT_DEFAULTS = {"tall_min_stages": 32, "tma_tf32": 1, "wait_hint": 100, "sm_budget": 148, "pair_remote": 1, "wgrad_wide": 1}
_nondefault = {}


def set_option(key, val):
    prev = _lib.load_tuning().mnv_debug_set_option(key.encode(), int(val))
    assert prev != -1 or val == -1, "unknown tuning option %r" % key
    if int(val) == _OPT_DEFAULTS.get(key, 0):
        _nondefault.pop(key, None)
    else:
        _nondefault[key] = int(val)
    return prev


def run(name, *args):
    conv = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]
    _lib.call(name, *conv, stream(), lib=_lib.load_tuning() if _nondefault else None)
    torch.cuda.synchronize()


_ws = None


def workspace(nbytes=768 << 20):
    global _ws
    if _ws is None or _ws.numel() < nbytes:
        _ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    return _ws


def ulp_diff(a, b):
    """Distance in units in the last place between two float32 arrays (same-sign finite values)."""
    ai = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def assert_bits_equal(got, want, msg=""):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    assert got.shape == want.shape, (got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), msg + " NaN pattern differs"
    ok = (got.view(np.uint32) == want.view(np.uint32)) | nan_g
    assert ok.all(), "%s %d/%d elements differ, first at %d: %r vs %r" % (
        msg, (~ok).sum(), ok.size, np.argmin(ok), got.ravel()[np.argmin(ok)], want.ravel()[np.argmin(ok)])


def norm_rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))
