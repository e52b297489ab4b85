#!/usr/bin/env python
"""Host-side cost of enqueueing one training step: cProfile of NetTrainer.step() (no sync inside), top functions by own time.
usage: host_profile.py [alexnet|googlenet|lenet|mlp] [steps]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import minerva_b200.owl as owl
import minerva_b200.owl.net as onet
name = sys.argv[1] if len(sys.argv) > 1 else "googlenet"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
wl = bench.WORKLOADS[name]
owl.set_device(owl.create_gpu_device(0))
owl.set_seed(1)
net = getattr(onet, wl["builder"])()
net.batch_size = wl["batch"]
x, onehot = bench.host_batch(wl, net.input_shape, wl["batch"], 100)
du = net.get_data_unit()
du.data, du.label = owl.from_numpy(x), owl.from_numpy(onehot)
tr = onet.NetTrainer(net, None)
for _ in range(5):
    tr.step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(steps):
    tr.step()
t_enq = time.perf_counter() - t
torch.cuda.synchronize()
t_all = time.perf_counter() - t
print("%s: enqueue %.3f ms/step, with drain %.3f ms/step" % (name, 1e3 * t_enq / steps, 1e3 * t_all / steps))
# one step at a time from an empty launch queue (a full queue makes the enqueue loop above wait for the device)
enq, tot = [], []
for _ in range(steps):
    torch.cuda.synchronize()
    t = time.perf_counter()
    tr.step()
    enq.append(time.perf_counter() - t)
    torch.cuda.synchronize()
    tot.append(time.perf_counter() - t)
enq.sort(); tot.sort()
print("%s: single step from an empty queue: enqueue median %.3f ms, enqueue+drain median %.3f ms" % (name, 1e3 * enq[len(enq) // 2], 1e3 * tot[len(tot) // 2]))
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    tr.step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
