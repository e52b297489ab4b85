#!/usr/bin/env python
"""Per-op device timing on the AlexNet (b256) shapes: achieved HBM GB/s for the memory-bound ops and
TFLOP/s for MatMult / conv, each with its roofline fraction against MEASURED_PEAKS.json.
CUDA events on the launching stream, warm-up, L2 flushed between timed launches.
Usage: python tools/opbench.py [--out gpurun_out/opbench.json] [--filter substr] [--batch 256]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from minerva_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/opbench.json")
ap.add_argument("--filter", default="")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--mnv-opt", action="append", default=[], metavar="KEY=INT", help="mnv_debug_set_option before timing (tuning)")
args = ap.parse_args()

lib = _lib.use_tuning() if args.mnv_opt else _lib.load()   # options exist in the tuning build only
if args.mnv_opt:
    import ctypes
    lib.mnv_debug_set_option.restype = ctypes.c_int
    lib.mnv_debug_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    for kv in args.mnv_opt:
        k, v = kv.split("=")
        lib.mnv_debug_set_option(k.encode(), int(v))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}
try:
    peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
    peaks["source"] = "measured"
except Exception:
    pass
HBM = peaks["hbm_gbs"]
TF32 = peaks["bf16_tflops"] / 2.0   # dense TF32 tensor peak = half the (measured) bf16 rate

st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ws = torch.empty(lib.mnv_workspace_bytes_hint(), dtype=torch.uint8, device="cuda")
B = args.batch


def rnd(n):
    return torch.randn(int(n), device="cuda")


def timeit(fn, iters=args.iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()   # > L2 (126 MB): evicts the operands
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


results = []


def report(name, seconds, bytes_=None, flops=None):
    r = {"op": name, "ms": seconds * 1e3}
    if bytes_ is not None:
        r.update(bound="hbm", gbs=bytes_ / seconds / 1e9, frac=bytes_ / seconds / 1e9 / HBM)
        print("%-46s %9.3f ms  %8.1f GB/s  %5.1f%% of HBM peak" % (name, r["ms"], r["gbs"], 100 * r["frac"]), flush=True)
    else:
        r.update(bound="tensor", tflops=flops / seconds / 1e12, frac=flops / seconds / 1e12 / TF32)
        print("%-46s %9.3f ms  %8.1f TFLOP/s %5.1f%% of TF32 peak" % (name, r["ms"], r["tflops"], 100 * r["frac"]), flush=True)
    results.append(r)


def want(name):
    return args.filter in name or (args.filter != "" and name in args.filter)


def call(name, *a):
    rc = getattr(lib, name)(*[x.data_ptr() if isinstance(x, torch.Tensor) else x for x in a], st)
    assert rc == 0, (name, rc)


# ---- memory-bound ops ---------------------------------------------------------------------------
E = 290400 * B  # conv1 activations
if want("add"):
    a, b, c = rnd(E), rnd(E), rnd(E)
    report("add 74.3M", timeit(lambda: call("mnv_add", a, b, c, E)), bytes_=12 * E)
    report("relu_forward 74.3M", timeit(lambda: call("mnv_relu_forward", a, c, 1, 1, 1, E)), bytes_=8 * E)
    report("relu_backward 74.3M", timeit(lambda: call("mnv_relu_backward", a, a, b, c, 1, 1, 1, E)), bytes_=12 * E)
    report("scale 74.3M", timeit(lambda: call("mnv_scale", a, c, E, 0.5)), bytes_=8 * E)
    report("fill 74.3M", timeit(lambda: call("mnv_fill", c, E, 0.0)), bytes_=4 * E)
    del a, b, c
if want("sgd"):
    n = 62378344
    w, d, g = rnd(n), rnd(n), rnd(n)
    report("sgd_momentum_update 62.4M params", timeit(lambda: call("mnv_sgd_momentum_update", w, d, g, n, 0.9, 1e-4, 5e-6)), bytes_=20 * n)
    del w, d, g
if want("pool"):
    for (C, H) in ((96, 55), (256, 27), (256, 13)):
        Ho = lib.mnv_pooled_size(H, 0, 3, 2)
        x, y = torch.relu(rnd(B * C * H * H)), rnd(B * C * Ho * Ho)
        dy, dx = rnd(B * C * Ho * Ho), rnd(B * C * H * H)
        ein, eout = x.numel(), y.numel()
        report("max_pool_fwd 3x3/2 C%d %d->%d" % (C, H, Ho), timeit(lambda: call("mnv_max_pooling_forward", x, y, B, C, H, H, 2, 2, 3, 3, 0, 0)), bytes_=4 * (ein + eout))
        report("max_pool_bwd 3x3/2 C%d %d->%d" % (C, H, Ho), timeit(lambda: call("mnv_max_pooling_backward", x, y, dy, dx, B, C, H, H, 2, 2, 3, 3, 0, 0)), bytes_=4 * (2 * ein + 2 * eout))
        # the arg-max remembering pair owl.net uses (5 B per pooled element on both sides instead of re-reading the bottom)
        idx = torch.empty(eout, dtype=torch.uint8, device="cuda")
        report("max_pool_fwd_idx 3x3/2 C%d %d->%d" % (C, H, Ho), timeit(lambda: call("mnv_max_pooling_forward_idx", x, y, idx, B, C, H, H, 2, 2, 3, 3, 0, 0)), bytes_=4 * ein + 5 * eout)
        report("max_pool_bwd_idx+relu 3x3/2 C%d %d->%d" % (C, H, Ho), timeit(lambda: call("mnv_max_pooling_backward_idx", dy, idx, y, dx, B, C, H, H, 2, 2, 3, 3, 0, 0)), bytes_=9 * eout + 4 * ein)
        del x, y, dy, dx, idx
if want("lrn"):
    for (C, H) in ((96, 55), (256, 27)):
        n = B * C * H * H
        x, sc, y, dy, dx = rnd(n), rnd(n), rnd(n), rnd(n), rnd(n)
        report("lrn_fwd C%d %dx%d" % (C, H, H), timeit(lambda: call("mnv_lrn_forward", x, sc, y, 5, 1e-4, 0.75, B, C, H, H)), bytes_=12 * n)
        report("lrn_bwd C%d %dx%d" % (C, H, H), timeit(lambda: call("mnv_lrn_backward", x, y, sc, dy, dx, 5, 1e-4, 0.75, B, C, H, H)), bytes_=20 * n)
        # the scale-less pair owl.net uses: algorithmic bytes 8 and 12 per element
        report("lrn_fwd_lite C%d %dx%d" % (C, H, H), timeit(lambda: call("mnv_lrn_forward_lite", x, y, 5, 1e-4, 0.75, B, C, H, H)), bytes_=8 * n)
        report("lrn_bwd_lite+relu C%d %dx%d" % (C, H, H), timeit(lambda: call("mnv_lrn_backward_lite", x, dy, dx, 5, 1e-4, 0.75, B, C, H, H, 1)), bytes_=12 * n)
        del x, sc, y, dy, dx
if want("bias"):
    for (C, H) in ((96, 55), (256, 27), (384, 13)):
        n = B * C * H * H
        dy, db = rnd(n), rnd(C)
        report("conv_backward_bias C%d %dx%d" % (C, H, H), timeit(lambda: call("mnv_conv_backward_bias", dy, db, B, C, H, H, ws, ws.numel())), bytes_=4 * (n + C))
        del dy
if want("misc"):
    m, n = 4096, B
    a, v, c = rnd(m * n), rnd(m), rnd(m * n)
    report("norm_add_on_row 4096x%d" % n, timeit(lambda: call("mnv_norm_add_on_row", a, v, c, m, n)), bytes_=8 * m * n)
    report("reduction_sum_on_row 4096x%d" % n, timeit(lambda: call("mnv_reduction_sum_on_row", a, v, m, n)), bytes_=4 * (m * n + m))
    x, y = rnd(1000 * n), rnd(1000 * n)
    report("instance_softmax_fwd 1000x%d" % n, timeit(lambda: call("mnv_instance_softmax_forward", x, y, n, 1, 1, 1000)), bytes_=8 * 1000 * n)
    w, wt = rnd(4096 * 9216), rnd(4096 * 9216)
    report("transpose 4096x9216", timeit(lambda: call("mnv_transpose", w, wt, 4096, 9216)), bytes_=8 * 4096 * 9216)
    mask = rnd(4096 * n)
    report("rand_bernoulli 4096x%d" % n, timeit(lambda: call("mnv_rand_bernoulli", mask, 4096 * n, 1, 0.5)), bytes_=4 * 4096 * n)

# ---- tensor-core ops ----------------------------------------------------------------------------
if want("matmult"):
    for (m, n, k) in ((4096, B, 9216), (4096, B, 4096), (1000, B, 4096), (9216, B, 4096), (4096, 9216, B), (4096, 4096, B), (8192, 8192, 8192)):
        a, b, c = rnd(m * k), rnd(k * n), rnd(m * n)
        report("matmult %dx%dx%d" % (m, n, k), timeit(lambda: call("mnv_matmult", a, b, c, m, n, k, ws, ws.numel())), flops=2.0 * m * n * k)
        del a, b, c
if want("conv"):
    layers = [("conv1", 3, 96, 227, 0, 4, 11), ("conv2", 96, 256, 27, 2, 1, 5), ("conv3", 256, 384, 13, 1, 1, 3),
              ("conv4", 384, 384, 13, 1, 1, 3), ("conv5", 384, 256, 13, 1, 1, 3)]
    for (name, Ci, Co, H, p, s, f) in layers:
        Ho = (H + 2 * p - f) // s + 1
        x, w, bias = rnd(B * Ci * H * H), rnd(Co * Ci * f * f), rnd(Co)
        y, dy = rnd(B * Co * Ho * Ho), rnd(B * Co * Ho * Ho)
        dx, dw = rnd(x.numel()), rnd(w.numel())
        geo = (B, Ci, Co, H, H, p, p, s, s, f, f)
        fl = 2.0 * B * Ho * Ho * Co * Ci * f * f
        report(name + " forward", timeit(lambda: call("mnv_conv_forward", x, w, bias, y, *geo, ws, ws.numel())), flops=fl)
        if name != "conv1":
            report(name + " backward_data", timeit(lambda: call("mnv_conv_backward_data", dy, w, dx, *geo, ws, ws.numel())), flops=fl)
        report(name + " backward_filter", timeit(lambda: call("mnv_conv_backward_filter", x, dy, dw, *geo, ws, ws.numel())), flops=fl)
        del x, w, y, dy, dx, dw

os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump({"peaks": peaks, "tf32_peak_tflops": TF32, "batch": B, "results": results}, open(args.out, "w"), indent=1)
print("wrote", args.out)
