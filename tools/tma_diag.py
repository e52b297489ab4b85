#!/usr/bin/env python
"""GPU triage for the all-TMA operand paths of the tcgen05 kernel (im2col tensor maps for conv forward /
backward-data / backward-filter, MN-major A for MatMult).  Each case runs the TMA path and the gather path
(mnv_debug_set_option("no_tma_a", 7)) on the same inputs: both round operands to TF32 the same way, so they
differ only by summation order.  Prints errors and timings, never raises -- run under `timeout`."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from minerva_b200 import _lib
from tests import gpu_util as g

lib = _lib.use_tuning()   # include/mnv_debug.h: runtime-settable options exist in the tuning build only
lib.mnv_debug_set_option.restype = ctypes.c_int
lib.mnv_debug_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
gen = torch.Generator(device="cuda").manual_seed(0)
quick = "--quick" in sys.argv
alex_only = "--alex-only" in sys.argv   # AlexNet shapes once each (for ncu launch lists)


def rnd(n):
    return torch.randn(int(n), device="cuda", generator=gen)


def opt(k, v):
    lib.mnv_debug_set_option(k.encode(), v)


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


VARIANTS = [("gather", {"no_tma_a": None}), ("tma", {"force_tma_a": 1, "no_tall": 1}), ("tall", {"force_tma_a": 1})]
ALL_KEYS = ("no_tma_a", "force_tma_a", "no_tall")


def compare(name, fn, out, flops, iters, bit):
    """fn() fills `out`; run it on the gather path, the all-TMA path and the all-TMA path with the 256-row tile."""
    res = {}
    for label, opts in VARIANTS:
        for k in ALL_KEYS:
            opt(k, 0)
        for k, v in opts.items():
            opt(k, bit if v is None else v)
        out.fill_(float("nan"))
        try:
            fn()
            torch.cuda.synchronize()
            res[label] = out.clone()
            res[label + "_ms"] = timed(fn, iters) if iters else float("nan")
        except Exception as e:  # noqa: BLE001
            print("  %s %s FAILED: %s" % (name, label, e), flush=True)
            break
    for k in ALL_KEYS:
        opt(k, 0)
    if "gather" not in res:
        return
    a = res["gather"].double()
    line = "%-44s" % name
    for label, _ in VARIANTS[1:]:
        if label not in res:
            continue
        b = res[label].double()
        err = float((a - b).norm() / max(float(a.norm()), 1e-30))
        line += " %s: err %.1e nan %d" % (label, err, int(torch.isnan(res[label]).sum()))
    if iters:
        line += " |" + "".join("  %s %.3f ms (%.0f TF/s)" % (l, res[l + "_ms"], flops / res[l + "_ms"] / 1e9) for l, _ in VARIANTS if l in res)
    print(line, flush=True)


ws = g.workspace()
S = g.stream


def conv_cases():
    # N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, iters
    small = [
        (2, 32, 16, 8, 8, 1, 1, 1, 1, 3, 3, 0),
        (3, 48, 40, 9, 7, 1, 1, 1, 1, 3, 3, 0),
        (2, 64, 24, 12, 10, 0, 0, 1, 1, 1, 1, 0),
        (2, 96, 64, 27, 27, 2, 2, 1, 1, 5, 5, 0),
        (4, 64, 32, 13, 13, 0, 0, 2, 2, 3, 3, 0),
        (2, 32, 32, 11, 14, 1, 2, 1, 1, 3, 5, 0),
        (2, 40, 36, 15, 15, 1, 1, 2, 3, 3, 3, 0),
        (5, 160, 300, 7, 7, 1, 1, 1, 1, 3, 3, 0),
    ]
    alex = [
        (256, 96, 256, 27, 27, 2, 2, 1, 1, 5, 5, 10),
        (256, 256, 384, 13, 13, 1, 1, 1, 1, 3, 3, 10),
        (256, 384, 384, 13, 13, 1, 1, 1, 1, 3, 3, 10),
        (256, 384, 256, 13, 13, 1, 1, 1, 1, 3, 3, 10),
    ]
    if alex_only:
        return [c[:-1] + (0,) for c in alex]
    return small + ([] if quick else alex)


for (N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, iters) in conv_cases():
    Ho, Wo = (H + 2 * ph - fh) // sv + 1, (W + 2 * pw - fw) // sh + 1
    x, w, b = rnd(N * Ci * H * W), rnd(Co * Ci * fh * fw) * 0.05, rnd(Co)
    dy = rnd(N * Co * Ho * Wo)
    flops = 2.0 * N * Ho * Wo * Co * Ci * fh * fw
    tag = "N%d C%d->%d %dx%d p%d,%d s%d,%d f%dx%d" % (N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw)
    y = g.empty(N * Co * Ho * Wo)
    compare("fwd   " + tag, lambda: _lib.call("mnv_conv_forward", x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, Ci, Co,
                                             H, W, ph, pw, sv, sh, fh, fw, ws.data_ptr(), ws.numel(), S()), y, flops, iters, 1)
    dx = g.empty(N * Ci * H * W)
    compare("dgrad " + tag, lambda: _lib.call("mnv_conv_backward_data", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), N, Ci, Co,
                                             H, W, ph, pw, sv, sh, fh, fw, ws.data_ptr(), ws.numel(), S()), dx, flops, iters, 1)
    dw = g.empty(Co * Ci * fh * fw)
    compare("wgrad " + tag, lambda: _lib.call("mnv_conv_backward_filter", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, Ci, Co,
                                             H, W, ph, pw, sv, sh, fh, fw, ws.data_ptr(), ws.numel(), S()), dw, flops, iters, 4)

mm = [(128, 64, 64, 0), (256, 96, 160, 0), (1000, 200, 300, 0), (132, 17, 36, 0), (4096, 256, 9216, 10), (256, 4096, 9216, 10),
      (9216, 4096, 256, 10)]
if not quick:
    mm.append((8192, 8192, 8192, 3))
if alex_only:
    mm = [(4096, 256, 9216, 0), (8192, 8192, 8192, 0)]
for (m, n, k, iters) in mm:
    a, b = rnd(m * k), rnd(k * n)
    c = g.empty(m * n)
    compare("matmult m=%d n=%d k=%d" % (m, n, k), lambda: _lib.call("mnv_matmult", a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k,
                                                                    ws.data_ptr(), ws.numel(), S()), c, 2.0 * m * n * k, iters, 2)

if alex_only:
    print("DIAG DONE")
    sys.exit(0)
# does a TFLOAT32-typed tensor map round (instead of truncate) what it delivers?  compare against fp64
m, n, k = 512, 256, 2048
a, b = rnd(m * k), rnd(k * n)
want = (b.view(n, k).double() @ a.view(k, m).double()).ravel()
c = g.empty(m * n)
for tf in (0, 1):
    opt("tma_tf32", tf)
    _lib.call("mnv_matmult", a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, ws.data_ptr(), ws.numel(), S())
    torch.cuda.synchronize()
    d = c.double() - want
    print("matmult tmap dtype %s: norm-rel err vs fp64 %.3e  mean signed rel %.3e" % (
        "TFLOAT32" if tf else "FLOAT32", float(d.norm() / want.norm()), float((d * want.sign()).mean() / want.abs().mean())), flush=True)
opt("tma_tf32", 0)
print("DIAG DONE")
