#!/usr/bin/env python
"""Does plane alignment bound the channel-marching LRN kernels?  Same kernel, planes of 55x55 (odd: every warp store
is misaligned to its 32-byte sectors) vs 56x56 (every warp access is a whole 128-byte line)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minerva_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
for (N, C, H, W) in [(256, 96, 55, 55), (256, 96, 56, 56), (256, 256, 27, 27), (256, 256, 28, 28), (256, 256, 32, 32)]:
    n = N * C * H * W
    x = torch.randn(n, device="cuda").abs() + 0.5
    sc, y, dy, dx = (torch.empty(n, device="cuda") for _ in range(4))
    dy.normal_()
    ms = timeit(lambda: _lib.call("mnv_lrn_forward", x.data_ptr(), sc.data_ptr(), y.data_ptr(), 5, 1e-4, 0.75, N, C, W, H, st))
    print("lrn_fwd N%d C%d %dx%d  %.3f ms  %.0f GB/s" % (N, C, H, W, ms, 3 * n * 4 / ms / 1e6), flush=True)
    ms = timeit(lambda: _lib.call("mnv_lrn_backward", x.data_ptr(), y.data_ptr(), sc.data_ptr(), dy.data_ptr(), dx.data_ptr(), 5, 1e-4, 0.75, N, C, W, H, st))
    print("lrn_bwd N%d C%d %dx%d  %.3f ms  %.0f GB/s" % (N, C, H, W, ms, 5 * n * 4 / ms / 1e6), flush=True)
