#!/usr/bin/env python
"""Launch one op a few times (for `ncu -k regex:... -s N -c M`).  Usage: one_op.py <name> [iters [KEY=INT ...]]
names: conv{1..5}_{fwd,bwd,wgrad}, fc6_fwd, gemm8k, pool1_fwd, pool1_bwd, pool1_bwd_idx, lrn1_fwd, lrn1_bwd, lrn1_fwd_lite, lrn1_bwd_lite, bias1"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from minerva_b200 import _lib

lib = _lib.use_tuning()   # include/mnv_debug.h: runtime-settable options exist in the tuning build only
name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for kv in sys.argv[3:]:      # KEY=INT tuning / debug options (mnv_debug_set_option)
    import ctypes
    lib.mnv_debug_set_option.restype = ctypes.c_int
    lib.mnv_debug_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.mnv_debug_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1]))
B = 256
st = torch.cuda.current_stream().cuda_stream
ws = torch.empty(lib.mnv_workspace_bytes_hint(), dtype=torch.uint8, device="cuda")


def rnd(n):
    return torch.randn(int(n), device="cuda")


def call(fn, *a):
    rc = getattr(lib, fn)(*[x.data_ptr() if isinstance(x, torch.Tensor) else x for x in a], st)
    assert rc == 0, (fn, rc)


LAYERS = {"conv1": (3, 96, 227, 0, 4, 11), "conv2": (96, 256, 27, 2, 1, 5), "conv3": (256, 384, 13, 1, 1, 3),
          "conv4": (384, 384, 13, 1, 1, 3), "conv5": (384, 256, 13, 1, 1, 3)}
if name.startswith("conv"):
    layer, kind = name.split("_")
    Ci, Co, H, p, s, f = LAYERS[layer]
    Ho = (H + 2 * p - f) // s + 1
    x, w, bias = rnd(B * Ci * H * H), rnd(Co * Ci * f * f), rnd(Co)
    y, dy = rnd(B * Co * Ho * Ho), rnd(B * Co * Ho * Ho)
    dx, dw = rnd(x.numel()), rnd(w.numel())
    geo = (B, Ci, Co, H, H, p, p, s, s, f, f)
    fns = {"fwd": lambda: call("mnv_conv_forward", x, w, bias, y, *geo, ws, ws.numel()),
           "bwd": lambda: call("mnv_conv_backward_data", dy, w, dx, *geo, ws, ws.numel()),
           "wgrad": lambda: call("mnv_conv_backward_filter", x, dy, dw, *geo, ws, ws.numel())}
    fn = fns[kind]
elif name in ("fc6_fwd", "gemm8k"):
    m, n, k = (4096, B, 9216) if name == "fc6_fwd" else (8192, 8192, 8192)
    a, b, c = rnd(m * k), rnd(k * n), rnd(m * n)
    fn = lambda: call("mnv_matmult", a, b, c, m, n, k, ws, ws.numel())
elif name.startswith("pool1"):
    C, H = 96, 55
    Ho = lib.mnv_pooled_size(H, 0, 3, 2)
    x, y = torch.relu(rnd(B * C * H * H)), rnd(B * C * Ho * Ho)
    dy, dx = rnd(B * C * Ho * Ho), rnd(B * C * H * H)
    if name.endswith("bwd_idx"):      # the arg-max remembering pair owl.net uses
        idx = torch.empty(y.numel(), dtype=torch.uint8, device="cuda")
        call("mnv_max_pooling_forward_idx", x, y, idx, B, C, H, H, 2, 2, 3, 3, 0, 0)
        fn = lambda: call("mnv_max_pooling_backward_idx", dy, idx, y, dx, B, C, H, H, 2, 2, 3, 3, 0, 0)
    elif name.endswith("fwd"):
        fn = lambda: call("mnv_max_pooling_forward", x, y, B, C, H, H, 2, 2, 3, 3, 0, 0)
    else:
        call("mnv_max_pooling_forward", x, y, B, C, H, H, 2, 2, 3, 3, 0, 0)
        fn = lambda: call("mnv_max_pooling_backward", x, y, dy, dx, B, C, H, H, 2, 2, 3, 3, 0, 0)
elif name.startswith("lrn1"):
    C, H = 96, 55
    n = B * C * H * H
    x, sc, y, dy, dx = rnd(n), rnd(n).abs() + 1, rnd(n), rnd(n), rnd(n)
    if name.endswith("fwd_lite"):
        fn = lambda: call("mnv_lrn_forward_lite", x, y, 5, 1e-4, 0.75, B, C, H, H)
    elif name.endswith("bwd_lite"):
        fn = lambda: call("mnv_lrn_backward_lite", x, dy, dx, 5, 1e-4, 0.75, B, C, H, H, 1)
    elif name.endswith("fwd"):
        fn = lambda: call("mnv_lrn_forward", x, sc, y, 5, 1e-4, 0.75, B, C, H, H)
    else:
        fn = lambda: call("mnv_lrn_backward", x, y, sc, dy, dx, 5, 1e-4, 0.75, B, C, H, H)
elif name == "bias1":
    C, H = 96, 55
    dy, db = rnd(B * C * H * H), rnd(C)
    fn = lambda: call("mnv_conv_backward_bias", dy, db, B, C, H, H, ws, ws.numel())
else:
    raise SystemExit("unknown op " + name)

for _ in range(iters):
    fn()
torch.cuda.synchronize()
print("done", name)
