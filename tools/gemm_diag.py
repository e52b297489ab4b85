#!/usr/bin/env python
"""GPU triage for the tcgen05 GEMM/conv kernel: prints the error of the SIMT checker and of the
tensor-core path per shape, and never raises -- run under `timeout` on the GPU box."""
import ctypes
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from minerva_b200 import _lib
from oracle import pyoracle as orc
from tests import gpu_util as g

lib = _lib.use_tuning()   # include/mnv_debug.h: runtime-settable options exist in the tuning build only
lib.mnv_debug_set_option.restype = ctypes.c_int
lib.mnv_debug_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
rng = np.random.default_rng(0)


def mm(a, b, m, n, k):
    ws = g.workspace()
    c = g.empty(m * n)
    c.fill_(float("nan"))
    g.run("mnv_matmult", g.dev(a), g.dev(b), c, m, n, k, ws, ws.numel())
    return g.host(c)


for (m, n, k) in [(128, 16, 32), (128, 16, 8), (128, 256, 32), (128, 16, 64), (256, 32, 128), (9, 7, 11), (300, 200, 100), (4096, 256, 1024)]:
    a = rng.normal(0, 1, m * k).astype(np.float32)
    b = rng.normal(0, 1, k * n).astype(np.float32)
    want = (b.reshape(n, k).astype(np.float64) @ a.reshape(k, m).astype(np.float64)).ravel()
    lib.mnv_debug_set_option(b"simt", 1)
    e_simt = g.norm_rel(mm(a, b, m, n, k), want)
    lib.mnv_debug_set_option(b"simt", 0)
    t0 = time.time()
    got = mm(a, b, m, n, k)
    e_tc = g.norm_rel(got, want)
    print("matmult m=%d n=%d k=%d  simt_err=%.2e  tcgen05_err=%.2e  nan=%d  (%.1f ms)" % (
        m, n, k, e_simt, e_tc, int(np.isnan(got).sum()), 1e3 * (time.time() - t0)), flush=True)
    if e_tc > 5e-3 and m * n <= 128 * 16:
        G, Wt = got.reshape(n, m), want.reshape(n, m)
        print("  first rows got :", G[0, :8])
        print("  first rows want:", Wt[0, :8])
        bad = np.argwhere(np.abs(G - Wt) > 1e-2 * np.abs(Wt).max())
        print("  #bad", len(bad), "first bad (n,m):", bad[:8].tolist())
print("DIAG DONE")
