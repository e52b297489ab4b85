"""Host-logic tests that need no GPU: the owl.net graph on the CPU oracle backend -- gradient check of
the layer graph, fused-vs-chained update, and the N>1 data-parallel path on gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny_net(backend):
    """AlexNet's layer types at toy size: conv-relu-lrn-pool-conv-relu-pool-fc-relu-dropout-fc-softmax."""
    from minerva_b200.owl.net.net import (Net, DataUnit, ConvConnection, ReluUnit, LRNUnit, PoolingUnit, FullyConnection,
                                          DropoutUnit, SoftmaxUnit)
    net = Net(backend)
    net.add_unit(DataUnit("data", ["data", "label"]))
    net.add_unit(ConvConnection("conv1", "data", "conv1", 6, 3, 2, 0, weight_std=0.3))
    net.add_unit(ReluUnit("relu1", "conv1", "c1r"))
    net.add_unit(LRNUnit("norm1", "c1r", "norm1", 5, 1e-2, 0.75))
    net.add_unit(PoolingUnit("pool1", "norm1", "pool1", 3, 2))
    net.add_unit(ConvConnection("conv2", "pool1", "conv2", 8, 3, 1, 1, weight_std=0.3, bias_value=0.1))
    net.add_unit(ReluUnit("relu2", "conv2", "c2r"))
    net.add_unit(PoolingUnit("pool2", "c2r", "pool2", 2, 2))
    net.add_unit(FullyConnection("fc6", "pool2", "fc6", 16, weight_std=0.2, bias_value=0.1))
    net.add_unit(ReluUnit("relu6", "fc6", "fc6r"))
    net.add_unit(DropoutUnit("drop6", "fc6r", "fc6d", 0.5))
    net.add_unit(FullyConnection("fc8", "fc6d", "fc8", 5, weight_std=0.2))
    net.add_unit(SoftmaxUnit("loss", "fc8", "label", "prob"))
    net.base_lr = net.current_lr = 0.05
    return net


def _batch(backend, n, seed=0, lo=0):
    rs = np.random.RandomState(seed)
    x = rs.normal(0, 1, (64, 3, 17, 17)).astype(np.float32)[lo:lo + n]
    lab = rs.randint(0, 5, 64)[lo:lo + n]
    onehot = np.zeros((n, 5), np.float32)
    onehot[np.arange(n), lab] = 1
    return backend.owl.from_numpy(x), backend.owl.from_numpy(onehot)


def test_numeric_gradient_of_the_graph():
    """The reference ships this as a tool (owl/net/gradient_checker.py); here it is a test."""
    from oracle import owl_cpu
    B = owl_cpu.Backend()
    owl_cpu.set_seed(3)
    net = _tiny_net(B)
    du = net.get_data_unit()
    du.data, du.label = _batch(B, 4)
    net.batch_size = 4
    net.forward("TEST")          # TEST: no dropout, deterministic
    net.backward("TEST")
    loss_unit = net.get_loss_units()[0]
    dsum, n = loss_unit.getloss_device()      # the reduction left on the device == the blocking getloss()
    assert abs(-float(dsum.to_numpy().reshape(-1)[0]) / n - loss_unit.getloss()) <= 1e-6 * abs(loss_unit.getloss())
    for uname, idxs in (("fc8", [0, 7, 33]), ("conv2", [1, 50, 200]), ("conv1", [0, 11, 100])):
        u = net.units[net.name_to_uid[uname]]
        g = u.weightgrad.a.copy()
        for i in idxs:
            h = 1e-2
            old = u.weight.a[i]
            u.weight.a[i] = old + h
            net.forward("TEST"); lp = loss_unit.getloss()
            u.weight.a[i] = old - h
            net.forward("TEST"); lm = loss_unit.getloss()
            u.weight.a[i] = old
            num = (lp - lm) / (2 * h) * 4    # getloss averages over the batch; grads are sums
            assert abs(num - g[i]) <= 2e-2 * max(1.0, abs(g[i])), (uname, i, num, g[i])


def test_fused_update_equals_reference_chain():
    from oracle import owl_cpu
    from minerva_b200.owl.net.trainer import NetTrainer
    B = owl_cpu.Backend()
    res = []
    for fused in (False, True):
        owl_cpu.set_seed(5)
        net = _tiny_net(B)
        du = net.get_data_unit()
        du.data, du.label = _batch(B, 8)
        net.batch_size = 8
        tr = NetTrainer(net, None, fused_update=fused)
        tr.step(); tr.step()
        res.append(np.concatenate([net.units[i].weight.a for i in net.get_weighted_unit_ids()]))
    np.testing.assert_allclose(res[0], res[1], rtol=1e-6, atol=1e-7)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import owl_cpu
    from minerva_b200.owl.net.trainer import NetTrainer
    B = owl_cpu.Backend()
    owl_cpu.set_seed(11)                       # identical initial weights on every rank
    net = _tiny_net(B)
    per = 8 // world
    du = net.get_data_unit()
    du.data, du.label = _batch(B, per, lo=rank * per)
    net.batch_size = 8                         # GLOBAL batch is the update divisor
    tr = NetTrainer(net, dist if world > 1 else None)
    assert tr.world == world
    net.forward("TEST"); net.backward("TEST"); tr._wait_merge()
    grads = np.concatenate([net.units[i].weightgrad.a for i in net.get_weighted_unit_ids()])
    net.forward("TEST"); net.backward("TEST"); tr._wait_merge()
    tr.fused_update = True
    # finish the iteration by hand (phase TEST keeps dropout out of the comparison)
    for uid in net.get_weighted_unit_ids():
        net.update(uid)
    w = np.concatenate([net.units[i].weight.a for i in net.get_weighted_unit_ids()])
    if rank == 0:
        np.save(out, np.stack([grads, w]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_data_parallel_matches_single_process(tmp_path):
    """2 ranks x 4 samples must give the merged gradient and the updated weights of 1 rank x 8 samples
    (the reference's claim, owl/owl/net/README.md:104), to 1e-5 relative (summation order differs)."""
    f1, f2 = str(tmp_path / "w1.npy"), str(tmp_path / "w2.npy")
    mp.spawn(_dp_worker, args=(1, _free_port(), f1), nprocs=1, join=True)
    mp.spawn(_dp_worker, args=(2, _free_port(), f2), nprocs=2, join=True)
    a, b = np.load(f1), np.load(f2)
    assert np.abs(a[0] - b[0]).max() <= 1e-5 * np.abs(a[0]).max()
    assert np.abs(a[1] - b[1]).max() <= 1e-5 * np.abs(a[1]).max()


def test_peer_merge_layout():
    """owl/net/merge.py plan_layout: slices are disjoint, 16-byte aligned, in backward order; every bucket splits into
    `world` shards of whole vectors; the FC gradients close the first bucket and the last unit has a bucket of its own."""
    from minerva_b200.owl.net.merge import plan_layout
    sizes = [("fc8", 4096000, 1000), ("fc7", 16777216, 4096), ("fc6", 37748736, 4096), ("conv5", 884736, 256),
             ("conv4", 1327104, 384), ("conv3", 884736, 384), ("conv2", 614400, 256), ("conv1", 34848, 96)]
    for world in (2, 3, 8):
        slices, buckets, total = plan_layout(sizes, world, last_fc=2)
        assert [b[2] for b in buckets] == ["fc6", "conv2", "conv1"]
        spans = sorted((o, o + n) for o, n in slices.values())
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and all(o % 4 == 0 for o, _ in spans)
        pos = 0
        for off, n, last in buckets:
            assert off == pos and n % (4 * world) == 0
            pos += n
        assert pos == total
        for (name, tag), (o, n) in slices.items():       # a slice lies inside the bucket its unit closes or precedes
            b = [i for i, (bo, bn, _) in enumerate(buckets) if bo <= o and o + n <= bo + bn]
            assert len(b) == 1
        fc_bytes = sum(n for (name, tag), (o, n) in slices.items() if name.startswith("fc"))
        assert buckets[0][1] >= fc_bytes
    # no fully-connected unit, two units: the first unit and the last one
    slices, buckets, total = plan_layout([("c2", 10, 2), ("c1", 7, 1)], 2, last_fc=-1)
    assert [b[2] for b in buckets] == ["c2", "c1"]
    slices, buckets, total = plan_layout([("only", 5, 1)], 4, last_fc=0)
    assert len(buckets) == 1 and buckets[0][1] % 16 == 0


def test_fusion_plan_flags():
    """Net._plan_fusion on a backend that advertises the fused entry points (no compute): a ReLU is folded into its
    producing convolution / its consuming LRN or max pooling only when it is the single reader / has a single reader."""
    from minerva_b200.owl.net.net import (Net, DataUnit, ConvConnection, ReluUnit, LRNUnit, PoolingUnit, FullyConnection,
                                          SoftmaxUnit, ConcatUnit)

    class _Co:
        FUSED_CONV_RELU = True
        FUSED_RELU_BACKWARD = True

    class _B:
        co, ele, owl = _Co, None, None

    def build():
        net = Net(_B())
        net.add_unit(DataUnit("data", ["data", "label"]))
        net.add_unit(ConvConnection("c1", "data", "c1", 8, 3))
        net.add_unit(ReluUnit("r1", "c1", "c1r"))              # single reader of c1; read by an LRN only
        net.add_unit(LRNUnit("n1", "c1r", "n1"))
        net.add_unit(ConvConnection("c2", "n1", "c2", 8, 3))
        net.add_unit(ReluUnit("r2", "c2", "c2r"))              # read by a max pool AND a conv: backward mask stays
        net.add_unit(PoolingUnit("p2", "c2r", "p2", 3, 2))
        net.add_unit(ConvConnection("c3", "c2r", "c3", 8, 1))
        net.add_unit(ReluUnit("r3", "c3", "c3r"))
        net.add_unit(PoolingUnit("p3", "c3r", "p3", 3, 2, pool="avg"))   # average pooling has no fused mask
        net.add_unit(ConvConnection("c4", "p2", "c4", 8, 1))
        net.add_unit(ReluUnit("r4a", "c4", "c4a"))             # c4 has two readers: no epilogue fusion
        net.add_unit(ReluUnit("r4b", "c4", "c4b"))
        net.add_unit(ConcatUnit("cat", ["c4a", "c4b", "p3"], "cat"))
        net.add_unit(FullyConnection("fc", "cat", "fc", 4))
        net.add_unit(SoftmaxUnit("loss", "fc", "label", "prob"))
        return net

    net = build()
    net._plan_fusion()
    u = {x.name: x for x in net.units}
    assert u["c1"].fuse_relu and u["r1"].fused and u["c2"].fuse_relu and u["c3"].fuse_relu and not u["c4"].fuse_relu
    assert not u["r4a"].fused and not u["r4b"].fused
    assert u["r1"].bp_fused and u["n1"].relu_bp                 # ReLU -> LRN
    assert not u["r2"].bp_fused and not u["p2"].relu_bp         # two readers
    assert not u["r3"].bp_fused and not u["p3"].relu_bp         # average pooling
    net = build()
    net.fuse_conv_relu = net.fuse_relu_backward = net.fuse_lrn_recompute = net.fuse_pool_index = net.fuse_conv_grads = False
    net._plan_fusion()
    for x in net.units:
        assert not getattr(x, "fuse_relu", False) and not getattr(x, "fused", False) and not getattr(x, "bp_fused", False)
        assert not getattr(x, "relu_bp", False)
        if isinstance(x, LRNUnit):
            assert not x.lite
        if isinstance(x, PoolingUnit):
            assert not x.use_idx
        if isinstance(x, ConvConnection):
            assert not x.fuse_grads


def test_in_place_units_match_distinct_names():
    """Caffe's in-place naming (relu / dropout with top == bottom, what every reference prototxt uses): the
    sensitivities are tracked per blob VERSION, so gradients equal those of the same net with distinct names
    (owl/owl/net/net.py:1102-1114 keeps per-unit dicts for the same reason)."""
    from oracle import owl_cpu
    from minerva_b200.owl.net.net import (Net, DataUnit, ConvConnection, ReluUnit, PoolingUnit, FullyConnection,
                                          DropoutUnit, SoftmaxUnit, ConcatUnit)

    def build(inplace):
        B = owl_cpu.Backend()
        owl_cpu.set_seed(9)
        n = (lambda a, b: a) if inplace else (lambda a, b: b)
        net = Net(B)
        net.add_unit(DataUnit("data", ["data", "label"]))
        net.add_unit(ConvConnection("conv1", "data", "conv1", 6, 3, 2, 0, weight_std=0.3))
        net.add_unit(ReluUnit("relu1", "conv1", n("conv1", "c1r")))
        # two consumers of the rectified blob: their contributions ARE summed
        net.add_unit(PoolingUnit("poolA", n("conv1", "c1r"), "pa", 2, 2))
        net.add_unit(PoolingUnit("poolB", n("conv1", "c1r"), "pb", 2, 2, pool="avg"))
        net.add_unit(ConcatUnit("cat", ["pa", "pb"], "cat"))
        net.add_unit(FullyConnection("fc1", "cat", "fc1", 12, weight_std=0.2, bias_value=0.1))
        net.add_unit(ReluUnit("relu2", "fc1", n("fc1", "f1r")))
        net.add_unit(DropoutUnit("drop", n("fc1", "f1r"), n("fc1", "f1d"), 0.5))
        net.add_unit(FullyConnection("fc2", n("fc1", "f1d"), "fc2", 5, weight_std=0.2))
        net.add_unit(SoftmaxUnit("loss", "fc2", "label", "prob"))
        du = net.get_data_unit()
        du.data, du.label = _batch(B, 6)
        net.batch_size = 6
        net.forward("TRAIN")
        net.backward("TRAIN")
        return net

    a, b = build(True), build(False)
    assert abs(a.get_loss_units()[0].getloss() - b.get_loss_units()[0].getloss()) == 0.0
    for uid in a.get_weighted_unit_ids():
        np.testing.assert_array_equal(a.units[uid].weightgrad.a, b.units[uid].weightgrad.a, err_msg=a.units[uid].name)
        np.testing.assert_array_equal(a.units[uid].biasgrad.a, b.units[uid].biasgrad.a, err_msg=a.units[uid].name)


def test_xavier_seeds_and_rank_salt():
    """Same-length unit names draw different xavier weights; ranks share weights but not dropout masks."""
    from oracle import owl_cpu
    from minerva_b200.owl.net.net import ConvConnection
    B = owl_cpu.Backend()
    ws = []
    for name in ("inception_4b/5x5_reduce", "inception_4c/5x5_reduce"):
        u = ConvConnection(name, "x", name, 24, 1, weight_filler="xavier")
        u.B = B
        u.wshape, u.bshape, u.fan_in = [1, 1, 512, 24], [24], 512
        u.init_weights_with_filler()
        ws.append(u.weight.a.copy())
    assert not np.array_equal(ws[0], ws[1])
    masks, weights = [], []
    for rank in (0, 1):
        owl_cpu.set_seed(4)
        owl_cpu.set_rank_salt(rank)
        weights.append(B.owl.randn([64], 0.0, 1.0).a.copy())
        masks.append(B.owl.randb([4096], 0.5).a.copy())
    owl_cpu.set_rank_salt(0)
    np.testing.assert_array_equal(weights[0], weights[1])
    assert not np.array_equal(masks[0], masks[1])


def test_graph_seeds_reproduce_next_seed():
    """_runtime.GraphSeeds (keys of a step recorded into a CUDA graph): word + k * stride, xor salt == what next_seed() hands
    the k-th generator call of that step; arm() leaves the host counter where the eager step would have left it."""
    import torch
    from minerva_b200.owl import _runtime as rt

    class Dev:
        device = torch.device("cpu")
    saved = list(rt._seed)
    try:
        for salt in (0, 5):
            rt.set_seed(0xC0FFEE)
            rt.set_rank_salt(salt)
            rt._seed[1] = 0xFFFFFFF0            # the 32-bit products wrap inside the step
            eager = [rt.next_seed(salted=True) for _ in range(6)]
            after = rt._seed[1]
            rt._seed[1] = 0xFFFFFFF0
            gs = rt.GraphSeeds(Dev())
            recorded = [gs.next(salted=True) for _ in range(6)]     # what the recording bakes in: (add, xor) per call
            gs.arm()
            word = int(gs.word[0]) & 0xFFFFFFFF
            assert [((word + add) & 0xFFFFFFFF) ^ xor for add, xor in recorded] == eager
            assert rt._seed[1] == after
    finally:
        rt._seed[:] = saved


def test_small_nets_merge_with_one_all_reduce():
    """merge.plan_layout + the SMALL_BYTES rule: LeNet / the MLP (1-2 MB of gradient) take the single all-reduce, AlexNet the
    bucketed peer exchange."""
    from minerva_b200.owl.net.merge import plan_layout, PeerGradMerge
    lenet = [("ip2", 5000, 10), ("ip1", 400000, 500), ("conv2", 25000, 50), ("conv1", 500, 20)]
    _, _, total = plan_layout(lenet, 8, 1)
    assert total * 4 <= PeerGradMerge.SMALL_BYTES
    alex = [("fc8", 4096000, 1000), ("fc7", 16777216, 4096), ("fc6", 37748736, 4096), ("conv5", 884736, 256), ("conv1", 34848, 96)]
    _, buckets, total = plan_layout(alex, 8, 2)
    assert total * 4 > PeerGradMerge.SMALL_BYTES and len(buckets) == 3
