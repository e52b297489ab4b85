#!/usr/bin/env python
"""Per-geometry device time of the convolution calls of one training step (call-by-call, CUDA events around every call).
usage: conv_breakdown.py [googlenet|alexnet]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import minerva_b200.owl as owl
import minerva_b200.owl.net as onet
from minerva_b200.owl import _runtime as rt
name = sys.argv[1] if len(sys.argv) > 1 else "googlenet"
if len(sys.argv) > 2:      # KEY=INT tuning options: the whole process runs on the tuning build of the library
    from minerva_b200 import _lib
    lib = _lib.use_tuning()
    for kv in sys.argv[2:]:
        lib.mnv_debug_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1]))
wl = bench.WORKLOADS[name]
owl.set_device(owl.create_gpu_device(0))
owl.set_seed(1)
net = getattr(onet, wl["builder"])()
net.batch_size = wl["batch"]
x, onehot = bench.host_batch(wl, net.input_shape, wl["batch"], 100)
du = net.get_data_unit()
du.data, du.label = owl.from_numpy(x), owl.from_numpy(onehot)
tr = onet.NetTrainer(net, None)
for _ in range(4):
    tr.step()
torch.cuda.synchronize()
rt.profiler = rt.EventProfiler()
steps = 3
for _ in range(steps):
    tr.step()
torch.cuda.synchronize()
tab = rt.profiler.table()
rt.profiler = None
agg = {}
for n, a, ms in tab:
    if "conv" not in n:
        continue
    off = 4 if n.startswith("mnv_conv_forward") or n == "mnv_conv_backward_filter_tw" else 3
    geo = tuple(a[off:off + 11])
    k = (n.replace("mnv_conv_", "").replace("_tw", ""), geo)
    e = agg.setdefault(k, [0, 0.0])
    e[0] += 1
    e[1] += ms
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
tot = sum(v[1] for v in agg.values()) / steps
print("conv calls: %.3f ms per step" % tot)
for (n, geo), (cnt, ms) in rows[:40]:
    N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw = geo
    Ho, Wo = (H + 2 * ph - fh) // sv + 1, (W + 2 * pw - fw) // sh + 1
    fl = 2.0 * N * Ho * Wo * Co * Ci * fh * fw
    per = ms / cnt
    print("%-16s Ci%4d Co%4d %3dx%-3d f%d s%d  x%d  %.3f ms/call  %6.1f TF/s  %.3f ms/step" % (n, Ci, Co, H, W, fh, sv, cnt // steps, per, fl / per / 1e9, ms / steps))
