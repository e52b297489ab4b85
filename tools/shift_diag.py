#!/usr/bin/env python
"""Diagnostic for the shift-GEMM convolution kernel: forward error against the oracle per geometry and descriptor mode."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from minerva_b200 import _lib
from oracle import pyoracle as orc

import ctypes
lib = _lib.use_tuning()   # include/mnv_debug.h: runtime-settable options exist in the tuning build only
st = torch.cuda.current_stream().cuda_stream
_so = lib.mnv_debug_set_option
_so.restype = ctypes.c_int
_so.argtypes = [ctypes.c_char_p, ctypes.c_int]
ws = torch.empty(lib.mnv_workspace_bytes_hint(), dtype=torch.uint8, device="cuda")
rng = np.random.default_rng(0)
CASES = [
    ("taps 1x1", (2, 3, 16, 32, 32, 0, 0, 4, 4, 4, 4)),
    ("taps 1x2", (2, 3, 16, 32, 32, 0, 0, 4, 4, 4, 8)),
    ("taps 2x1 Wv=8", (2, 3, 16, 36, 32, 0, 0, 4, 4, 8, 4)),
    ("taps 2x1 Wv=7", (2, 3, 16, 32, 28, 0, 0, 4, 4, 8, 4)),
    ("taps 3x3", (2, 3, 8, 23, 23, 0, 0, 4, 4, 11, 11)),
    ("conv1 small", (2, 3, 96, 67, 67, 0, 0, 4, 4, 11, 11)),
]
for name, case in CASES:
    N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw = case
    Ho, Wo = orc.conv_out(H, ph, fh, sv), orc.conv_out(W, pw, fw, sh)
    x = rng.normal(0, 1, N * Ci * H * W).astype(np.float32)
    w = rng.normal(0, 1, Co * Ci * fh * fw).astype(np.float32)
    b = rng.normal(0, 1, Co).astype(np.float32)
    want = orc.conv_forward(x, w, b, *case)
    for bo in (0,):
        y = torch.full((N * Co * Ho * Wo,), float("nan"), device="cuda")
        xd, wd, bd = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda()
        rc = 0
        _lib.call("mnv_conv_forward", xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), *case, ws.data_ptr(), ws.numel(), st)
        torch.cuda.synchronize()
        got = y.cpu().numpy()
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        bad = np.abs(got - want) > 1e-2 * np.abs(want).max()
        idx = np.nonzero(bad.reshape(N, Co, Ho * Wo).any(axis=1))
        print("%-16s bo=%d rc=%d err=%.3e bad pixels %d / %d  first bad (n,pix): %s" % (
            name, bo, rc, err, len(idx[0]), N * Ho * Wo, list(zip(idx[0][:6], idx[1][:6]))))
