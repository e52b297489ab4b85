#!/usr/bin/env python
"""Where does the end-to-end AlexNet step lose time against the resident-input step?  Same recorded graph, loops that add one
thing at a time: (A) replay only, (B) + device-to-device copy of a fresh input, (C) + loss copied to pinned memory and read one
step late, (D) the uint8 feed (upload + device transform) without the loss read, (E) everything (= bench e2e)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import minerva_b200.owl as owl
import minerva_b200.owl.net as onet
from minerva_b200.owl import _runtime as rt
from minerva_b200.owl.net.data import HostFeed
wl = bench.WORKLOADS["alexnet"]
owl.set_device(owl.create_gpu_device(0))
owl.set_seed(1)
net = getattr(onet, wl["builder"])()
B = wl["batch"]
net.batch_size = B
x, onehot = bench.host_batch(wl, net.input_shape, B, 100)
du = net.get_data_unit()
du.data, du.label = owl.from_numpy(x), owl.from_numpy(onehot)
tr = onet.NetTrainer(net, None, graph=True)
for _ in range(5):
    tr.step()
torch.cuda.synchronize()
static = (du.data, du.label)
other = (owl.from_numpy(x), owl.from_numpy(onehot))
dev = rt.current_device()
host_loss = [torch.zeros(1).pin_memory() for _ in range(2)]
ev = [torch.cuda.Event(), torch.cuda.Event()]
steps = 30


def run(name, body, setup=None, teardown=None):
    if setup:
        setup()
    for _ in range(3):
        body(0)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for it in range(steps):
        body(it)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / steps
    if teardown:
        teardown()
    print("%-52s %.3f ms/step  %.0f images/s" % (name, 1e3 * dt, B / dt), flush=True)


def a(it):
    du.data, du.label = static
    tr.step()


def b(it):
    du.data, du.label = other
    tr.step()


pend = [None]


def loss_read(it):
    cur = it % 2
    dsum, n = tr.loss_device
    host_loss[cur].copy_(dsum.as_torch(), non_blocking=True)
    ev[cur].record(dev.stream)
    if pend[0] is not None:
        ev[pend[0]].synchronize()
        float(host_loss[pend[0]][0])
    pend[0] = cur


def c(it):
    b(it)
    loss_read(it)


rs = np.random.RandomState(7)
stored = rs.randint(0, 256, (B, 3, 256, 256), dtype=np.uint8)
feed = HostFeed(owl, rt, data_u8=stored, mean=np.full((3, 256, 256), 127.5, np.float32), scale=1.0 / 73.9, crop=(227, 227), mirror=True, label=onehot, seed=0)


def d(it):
    du.data, du.label = feed.next()
    tr.step()
    feed.done()


def e(it):
    d(it)
    loss_read(it)


run("A replay, resident input", a)
run("B + device copy of a fresh input", b)
run("C + loss to pinned memory, read one step late", c)
run("D uint8 feed (upload + transform), no loss read", d, setup=feed.start, teardown=feed.stop)
pend[0] = None
run("E uint8 feed + loss read (bench e2e)", e, setup=feed.start, teardown=feed.stop)
