#!/usr/bin/env python
"""Time one op under different mnv_debug_set_option values.  Usage: tune_opt.py <key> <v1,v2,...> <op> [op...]"""
import ctypes
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from minerva_b200 import _lib

lib = _lib.use_tuning()   # include/mnv_debug.h: runtime-settable options exist in the tuning build only
lib.mnv_debug_set_option.restype = ctypes.c_int
lib.mnv_debug_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
key, vals, ops = sys.argv[1], [int(v, 0) for v in sys.argv[2].split(",")], sys.argv[3:]
B = 256
st = torch.cuda.current_stream().cuda_stream
ws = torch.empty(lib.mnv_workspace_bytes_hint(), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
LAYERS = {"conv1": (3, 96, 227, 0, 4, 11), "conv2": (96, 256, 27, 2, 1, 5), "conv3": (256, 384, 13, 1, 1, 3),
          "conv4": (384, 384, 13, 1, 1, 3), "conv5": (384, 256, 13, 1, 1, 3)}


def call(fn, *a):
    rc = getattr(lib, fn)(*[x.data_ptr() if isinstance(x, torch.Tensor) else x for x in a], st)
    assert rc == 0, (fn, rc)


def make(op):
    if op.startswith("conv"):
        layer, kind = op.split("_")
        Ci, Co, H, p, s, f = LAYERS[layer]
        Ho = (H + 2 * p - f) // s + 1
        x, w, bias = torch.randn(B * Ci * H * H, device="cuda"), torch.randn(Co * Ci * f * f, device="cuda"), torch.randn(Co, device="cuda")
        y, dy = torch.randn(B * Co * Ho * Ho, device="cuda"), torch.randn(B * Co * Ho * Ho, device="cuda")
        dx, dw = torch.empty_like(x), torch.empty_like(w)
        geo = (B, Ci, Co, H, H, p, p, s, s, f, f)
        fl = 2.0 * B * Ho * Ho * Co * Ci * f * f
        return {"fwd": lambda: call("mnv_conv_forward", x, w, bias, y, *geo, ws, ws.numel()),
                "bwd": lambda: call("mnv_conv_backward_data", dy, w, dx, *geo, ws, ws.numel()),
                "wgrad": lambda: call("mnv_conv_backward_filter", x, dy, dw, *geo, ws, ws.numel())}[kind], fl
    m, n, k = {"gemm8k": (8192, 8192, 8192), "fc6": (4096, B, 9216), "fc6dw": (4096, 9216, B)}[op]
    a, b, c = torch.randn(m * k, device="cuda"), torch.randn(k * n, device="cuda"), torch.empty(m * n, device="cuda")
    return (lambda: call("mnv_matmult", a, b, c, m, n, k, ws, ws.numel())), 2.0 * m * n * k


for op in ops:
    fn, fl = make(op)
    for v in vals:
        lib.mnv_debug_set_option(key.encode(), v)
        for _ in range(3):
            fn()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print("%-12s %s=%-10d %8.3f ms %8.1f TFLOP/s" % (op, key, v, ts[3], fl / ts[3] / 1e9), flush=True)
