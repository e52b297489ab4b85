#!/usr/bin/env python
"""Which operand stream bounds the all-TMA convolution kernels?  AlexNet conv2-5, all three directions, twins current (no
pre-pass in the timed call), with the A copies, the B copies or both skipped (tuning build, pf_dist = -1 / -2 / -3: results are
garbage, the time is what is read).  'none' with both skipped = the tensor pipe + epilogue alone."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minerva_b200 import _lib
lib = _lib.use_tuning()
for kv in sys.argv[1:]:      # extra KEY=INT tuning options
    lib.mnv_debug_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1]))
st = torch.cuda.current_stream().cuda_stream
ws = torch.empty(lib.mnv_workspace_bytes_hint(), dtype=torch.uint8, device="cuda")
B = 256


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


LAYERS = {"conv2": (96, 256, 27, 2, 1, 5), "conv3": (256, 384, 13, 1, 1, 3), "conv4": (384, 384, 13, 1, 1, 3), "conv5": (384, 256, 13, 1, 1, 3)}
print("%-22s %9s %9s %9s %9s   (ms; TF/s of the full run)" % ("call", "full", "no A", "no B", "neither"))
for name, (Ci, Co, H, p, s, f) in LAYERS.items():
    Ho = (H + 2 * p - f) // s + 1
    x, w, b = torch.randn(B * Ci * H * H, device="cuda"), torch.randn(Co * Ci * f * f, device="cuda"), torch.randn(Co, device="cuda")
    y, dy = torch.empty(B * Co * Ho * Ho, device="cuda"), torch.randn(B * Co * Ho * Ho, device="cuda")
    dx, dw = torch.empty_like(x), torch.empty_like(w)
    geo = (B, Ci, Co, H, H, p, p, s, s, f, f)
    xt = torch.empty(lib.mnv_conv_twin_bytes(B, Ci, H, H) // 4, device="cuda")
    dt = torch.empty(lib.mnv_conv_twin_bytes(B, Co, Ho, Ho) // 4, device="cuda")
    xs, ds = ctypes.c_int(0), ctypes.c_int(0)
    calls = {
        "forward": lambda: lib.mnv_conv_forward_tw(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *geo, 1, xt.data_ptr(), ctypes.byref(xs), ws.data_ptr(), ws.numel(), st),
        "backward_data": lambda: lib.mnv_conv_backward_data_tw(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), *geo, dt.data_ptr(), ctypes.byref(ds), ws.data_ptr(), ws.numel(), st),
        "backward_filter": lambda: lib.mnv_conv_backward_filter_tw(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), 0, *geo, xt.data_ptr(), ctypes.byref(xs), dt.data_ptr(), ctypes.byref(ds), ws.data_ptr(), ws.numel(), st),
    }
    fl = 2.0 * B * Ho * Ho * Co * Ci * f * f
    for cname, fn in calls.items():
        row = []
        for skip in (0, -1, -2, -3):
            lib.mnv_debug_set_option(b"pf_dist", skip)
            row.append(timeit(fn))
        lib.mnv_debug_set_option(b"pf_dist", 0)
        print("%-22s %9.3f %9.3f %9.3f %9.3f   %.0f TF/s" % (name + " " + cname, *row, fl / row[0] / 1e9), flush=True)
