#!/usr/bin/env python
"""Is a TMA im2col box slower than a tiled box of the same bytes?  A 1x1 / stride 1 / pad 0 convolution forward with a
current channels-last twin (im2col map over [n][h][w][C]) against the same product as mnv_matmult_ex with the twin as a
K-major A operand (tiled map over [pixels][C]): same kernel, same tiles, same bytes -- only the copy mode differs."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minerva_b200 import _lib
lib = _lib.use_tuning()
for kv in sys.argv[1:]:
    lib.mnv_debug_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1]))
st = torch.cuda.current_stream().cuda_stream
ws = torch.empty(lib.mnv_workspace_bytes_hint(), dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (N, C, Co, H, f, p) in ((256, 384, 384, 13, 1, 0), (256, 256, 384, 13, 1, 0), (256, 384, 384, 13, 3, 1), (120, 480, 192, 14, 1, 0), (120, 192, 64, 28, 1, 0)):
    x = torch.randn(N * C * H * H, device="cuda")
    w = torch.randn(Co * C * f * f, device="cuda")
    b = torch.randn(Co, device="cuda")
    y = torch.empty(N * Co * H * H, device="cuda")
    geo = (N, C, Co, H, H, p, p, 1, 1, f, f)
    tw = torch.empty(lib.mnv_conv_twin_bytes(N, C, H, H) // 4, device="cuda")
    state = ctypes.c_int(0)

    def conv():
        rc = lib.mnv_conv_forward_tw(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *geo, 0, tw.data_ptr(), ctypes.byref(state), ws.data_ptr(), ws.numel(), st)
        assert rc == 0, rc
    t_conv = timeit(conv)
    fl = 2.0 * N * H * H * Co * C * f * f
    line = "N%d C%d Co%d %dx%d f%d: conv (im2col A, twin current) %.3f ms %.0f TF/s" % (N, C, Co, H, H, f, t_conv, fl / t_conv / 1e9)
    if f == 1:
        M, K = N * H * H, C
        c = torch.empty(M * Co, device="cuda")

        def gemm():     # c{M x Co} = op(a){M x K} op(b){K x Co}: a stored {K, M} (= the twin's [pixel][C] rows), b = filter [co][c] = {K, Co} column-major
            rc = lib.mnv_matmult_ex(tw.data_ptr(), w.data_ptr(), c.data_ptr(), M, Co, K, 1, 0, ws.data_ptr(), ws.numel(), st)
            assert rc == 0, rc
        t_gemm = timeit(gemm)
        line += " | matmult_ex (tiled A) %.3f ms %.0f TF/s" % (t_gemm, fl / t_gemm / 1e9)
    print(line, flush=True)
