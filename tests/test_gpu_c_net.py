"""-m gpu: a whole training step of the owl.net graph on the B200 kernels vs the same graph on the CPU
oracle (same Philox seeds => same initial weights and dropout masks).  Also the owl API surface."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_owl():
    import minerva_b200.owl as owl
    dev = owl.create_gpu_device(0)
    owl.set_device(dev)
    return owl


def test_owl_api_surface(gpu_owl):
    owl = gpu_owl
    import minerva_b200.owl.conv as co
    import minerva_b200.owl.elewise as ele
    a = owl.from_numpy(np.arange(6, dtype=np.float32).reshape(2, 3))     # numpy (2,3) -> owl [3,2]
    assert a.shape == [3, 2]
    np.testing.assert_array_equal(a.to_numpy(), np.arange(6, dtype=np.float32).reshape(2, 3))
    b = owl.ones([3, 2])
    np.testing.assert_array_equal((a + b).to_numpy(), a.to_numpy() + 1)
    np.testing.assert_array_equal((a * 2 - 1).to_numpy(), a.to_numpy() * 2 - 1)
    np.testing.assert_array_equal((3 - a).to_numpy(), 3 - a.to_numpy())
    np.testing.assert_array_equal((-a).to_numpy(), -a.to_numpy())
    np.testing.assert_array_equal(ele.mult(a, a).to_numpy(), a.to_numpy() ** 2)
    np.testing.assert_array_equal(a.trans().to_numpy(), a.to_numpy().T)
    # matrix product: owl [3,2] * [2,4] -> [3,4]; numpy sees the transposes
    c = owl.from_numpy(np.arange(8, dtype=np.float32).reshape(4, 2))      # owl [2,4]
    np.testing.assert_allclose((a * c).to_numpy(), c.to_numpy() @ a.to_numpy(), rtol=2e-3)
    # broadcast add of a {3,1} bias over columns (NormArithmetic)
    bias = owl.from_numpy(np.array([[10.0, 20.0, 30.0]], np.float32))    # owl [3,1]
    np.testing.assert_array_equal((a + bias).to_numpy(), a.to_numpy() + np.array([10, 20, 30], np.float32))
    np.testing.assert_array_equal(a.sum(0).to_numpy(), a.to_numpy().sum(1, keepdims=True))
    np.testing.assert_array_equal(a.max(1).to_numpy(), a.to_numpy().max(0, keepdims=True))
    np.testing.assert_array_equal(a.max_index(0).to_numpy(), a.to_numpy().argmax(1)[:, None].astype(np.float32))
    assert owl.zeros([4, 4]).count_zero() == 16
    x = owl.randn([8, 8, 3, 2], 0.0, 1.0)
    y = co.Pooler(2, 2, 2, 2).ff(x)
    assert y.shape == [4, 4, 3, 2]
    cc = owl.concat([x, x], 2)
    assert cc.shape == [8, 8, 6, 2]
    np.testing.assert_array_equal(owl.slice(cc, 2, 3, 3).to_numpy(), x.to_numpy())
    s = co.softmax(owl.randn([10, 4], 0, 1))
    np.testing.assert_allclose(s.to_numpy().sum(1), 1.0, rtol=1e-5)
    from minerva_b200 import _lib
    cpu = owl.create_cpu_device()
    owl.set_device(cpu)
    with pytest.raises(_lib.MnvError):        # no CPU fallback on the product path
        owl.zeros([2, 2])
    owl.set_device(0)


@pytest.mark.parametrize("m,k,n", [(8, 12, 4), (7, 5, 3), (256, 128, 64), (300, 200, 36), (1000, 256, 4096)])
def test_lazy_transpose(gpu_owl, m, k, n):
    """`trans()` defers its copy: a product consumes it through mnv_matmult_ex, anything else materialises it, and the
    in-place update flushes pending views of the array it overwrites.  Values must match the eager semantics."""
    owl = gpu_owl
    from minerva_b200.owl import NArray
    rng = np.random.default_rng(7)
    A = rng.normal(0, 1, (k, m)).astype(np.float32)      # owl [m,k]
    B = rng.normal(0, 1, (n, k)).astype(np.float32)      # owl [k,n]
    a, b = owl.from_numpy(A), owl.from_numpy(B)
    want = (B.astype(np.float64) @ A.astype(np.float64))             # numpy view of owl a*b
    tol = 5e-3 * np.linalg.norm(want)
    at, bt = owl.from_numpy(np.ascontiguousarray(A.T)), owl.from_numpy(np.ascontiguousarray(B.T))   # owl [k,m], [n,k]
    assert at.trans()._lazy_src is at
    for got in (at.trans() * b, a * bt.trans(), at.trans() * bt.trans()):
        assert got.shape == [m, n]
        assert np.linalg.norm(got.to_numpy() - want) <= tol
    # non-product consumers see the transposed values
    np.testing.assert_array_equal(at.trans().to_numpy(), A)
    np.testing.assert_array_equal((at.trans() + a).to_numpy(), A + A)
    np.testing.assert_array_equal(at.trans().trans().to_numpy(), A.T)
    # a pending view survives an in-place update of its source with the OLD values (eager semantics)
    w = owl.from_numpy(A.copy())
    view = w.trans()
    NArray.sgd_update(w, owl.zeros(w.shape), owl.ones(w.shape), 0.0, 1.0, 0.0)     # w -= 1
    np.testing.assert_array_equal(view.to_numpy(), A.T)
    np.testing.assert_array_equal(w.to_numpy(), A - 1)


def test_training_step_matches_cpu_oracle(gpu_owl):
    from tests.test_net_cpu import _tiny_net, _batch
    from oracle import owl_cpu
    from minerva_b200.owl.net.net import _default_backend
    from minerva_b200.owl.net.trainer import NetTrainer
    nets = []
    for B, seed_fn in ((owl_cpu.Backend(), owl_cpu.set_seed), (_default_backend(), gpu_owl.set_seed)):
        seed_fn(21)
        net = _tiny_net(B)
        du = net.get_data_unit()
        du.data, du.label = _batch(B, 8)
        net.batch_size = 8
        net.forward("TRAIN")
        net.backward("TRAIN")
        nets.append(net)
    cpu, gpu = nets
    lc, lg = cpu.get_loss_units()[0].getloss(), gpu.get_loss_units()[0].getloss()
    assert abs(lc - lg) <= 5e-3 * abs(lc), (lc, lg)          # TF32 conv/GEMM tolerance carried to the loss
    for uid in cpu.get_weighted_unit_ids():
        for attr in ("weight", "weightgrad", "biasgrad"):
            a = getattr(cpu.units[uid], attr).to_numpy().astype(np.float64)
            b = getattr(gpu.units[uid], attr).to_numpy().astype(np.float64)
            err = np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-12)
            assert err < 5e-3, (cpu.units[uid].name, attr, err)     # TF32 conv/GEMM tolerance
    # dropout masks are bit-identical (same Philox stream)
    dc = [u for u in cpu.units if u.name == "drop6"][0].dropmask.to_numpy()
    dg = [u for u in gpu.units if u.name == "drop6"][0].dropmask.to_numpy()
    np.testing.assert_array_equal(dc, dg)
    # and a full trainer step (fused update) runs and changes the weights identically to the op chain
    w0 = gpu.units[1].weight.to_numpy().copy()
    NetTrainer(gpu, None, fused_update=True).step()
    assert not np.array_equal(w0, gpu.units[1].weight.to_numpy())


def test_conv_relu_fusion_is_bit_identical(gpu_owl):
    """Net._plan_fusion: the fused graph (conv epilogue rectifies, ReluUnit passes through) gives the same bits."""
    from tests.test_net_cpu import _tiny_net, _batch
    from minerva_b200.owl.net.net import _default_backend, ConvConnection, ReluUnit
    res = []
    for fuse in (True, False):
        gpu_owl.set_seed(5)
        net = _tiny_net(_default_backend())
        net.fuse_conv_relu = net.fuse_relu_backward = net.fuse_lrn_recompute = net.fuse_pool_index = net.fuse_conv_grads = net.fuse_conv_twins = fuse
        du = net.get_data_unit()
        du.data, du.label = _batch(net.B, 8)
        net.batch_size = 8
        net.forward("TRAIN")
        net.backward("TRAIN")
        n_fused = sum(1 for u in net.units if isinstance(u, ConvConnection) and u.fuse_relu)
        assert (n_fused > 0) == fuse and n_fused == sum(1 for u in net.units if isinstance(u, ReluUnit) and u.fused)
        # relu1 -> LRN and relu2 -> max pool hand their backward masks over; relu6 -> dropout keeps its own pass
        assert sum(1 for u in net.units if isinstance(u, ReluUnit) and u.bp_fused) == (2 if fuse else 0)
        res.append([net.units[uid].weightgrad.to_numpy() for uid in net.get_weighted_unit_ids()]
                   + [net.get_loss_units()[0].ff_y.to_numpy()])
    for a, b in zip(*res):
        np.testing.assert_array_equal(a, b)


def test_graph_step_is_bit_identical(gpu_owl):
    """NetTrainer(graph=True): the step recorded into a CUDA graph and replayed leaves, step after step, the bits of the eager
    trainer -- weights, loss and the dropout masks (whose keys the replay reads from a device word) -- also when the
    caller swaps the input arrays and changes the learning rate (a new recording)."""
    from tests.test_net_cpu import _tiny_net, _batch
    from minerva_b200.owl.net.net import _default_backend
    from minerva_b200.owl.net.trainer import NetTrainer
    out = []
    for graph in (False, True):
        gpu_owl.set_seed(21)
        net = _tiny_net(_default_backend())
        net.batch_size = 8
        du = net.get_data_unit()
        tr = NetTrainer(net, None, graph=graph)
        losses, masks = [], []
        for it in range(7):
            if it == 4:
                net.current_lr = net.current_lr * 0.5
            du.data, du.label = _batch(net.B, 8, seed=it % 3)         # a fresh input array every step
            tr.step()
            if graph:
                dsum, n = tr.loss_device
                losses.append(-float(dsum.to_numpy().reshape(-1)[0]) / n)
            else:
                losses.append(net.get_loss_units()[0].getloss())
            masks.append([u for u in net.units if u.name == "drop6"][0].dropmask.to_numpy().copy())
        if graph:
            assert tr.graph_replays == 6 and tr.graph_launches_per_step > 20      # the first step ran eagerly (lazy initialisation)
        out.append((losses, masks, [net.units[uid].weight.to_numpy() for uid in net.get_weighted_unit_ids()]
                    + [net.units[uid].biasdelta.to_numpy() for uid in net.get_weighted_unit_ids()]))
    (l0, m0, w0), (l1, m1, w1) = out
    assert l0 == l1
    for a, b in zip(m0 + w0, m1 + w1):
        np.testing.assert_array_equal(a, b)
    assert not np.array_equal(m0[0], m0[1])          # the masks do change from step to step
    # an eager draw after the replays continues the same stream as after the eager steps (host counter kept in step)


def test_graph_step_with_feed(gpu_owl):
    """The recorded step driven by a FeedDataUnit (uint8 upload + device transform): same weights as the eager loop."""
    from tests.test_net_cpu import _tiny_net
    from minerva_b200.owl.net.net import _default_backend
    from minerva_b200.owl.net.trainer import NetTrainer
    from minerva_b200.owl.net.data import HostFeed, FeedDataUnit
    import minerva_b200.owl._runtime as rt
    res = []
    for graph in (False, True):
        gpu_owl.set_seed(3)
        net = _tiny_net(_default_backend())
        net.batch_size = 8
        rs = np.random.RandomState(1)
        stored = rs.randint(0, 256, (8, 3, 17, 17), dtype=np.uint8)
        onehot = np.zeros((8, 5), np.float32)
        onehot[np.arange(8), rs.randint(0, 5, 8)] = 1
        feed = HostFeed(gpu_owl, rt, data_u8=stored, scale=1.0 / 255.0, label=onehot)
        old = net.units[0]
        fu = FeedDataUnit(old.name, old.top_names, feed)
        fu.B = old.B
        net.units[0] = fu
        tr = NetTrainer(net, None, graph=graph)
        for _ in range(9):
            tr.step()
        fu.close()
        if graph:      # step 1 eager, 2-3 through the copying recording, then one recording bound to each of the feed's two device slots
            assert len(tr._bound) == 2 and tr.graph_replays == 8
        res.append([net.units[uid].weight.to_numpy() for uid in net.get_weighted_unit_ids()])
    for a, b in zip(*res):
        np.testing.assert_array_equal(a, b)


def test_relu_backward_twin_is_bit_identical(gpu_owl):
    """conv -> relu -> conv chains (AlexNet conv3..5, every GoogLeNet reduce -> 3x3 / 5x5 pair): the ReluUnit's backward leaves
    the channels-last twin and channel sums of its result (mnv_relu_backward_tw) for the convolution below; gradients are
    the bits of the graph without twins."""
    from minerva_b200.owl.net.net import (_default_backend, Net, DataUnit, ConvConnection, ReluUnit, FullyConnection, SoftmaxUnit)
    res = []
    for twins in (True, False):
        gpu_owl.set_seed(9)
        net = Net(_default_backend())
        net.add_unit(DataUnit("data", ["data", "label"]))
        net.add_unit(ConvConnection("c1", "data", "c1", 64, 3, 1, 1, weight_std=0.1))
        net.add_unit(ReluUnit("r1", "c1", "c1"))                      # in place, the Caffe convention
        net.add_unit(ConvConnection("c2", "c1", "c2", 96, 3, 1, 1, weight_std=0.05, bias_value=0.1))
        net.add_unit(ReluUnit("r2", "c2", "c2r"))
        net.add_unit(ConvConnection("c3", "c2r", "c3", 48, 1, 1, 0, weight_std=0.05))
        net.add_unit(ReluUnit("r3", "c3", "c3"))
        net.add_unit(FullyConnection("fc", "c3", "fc", 5, weight_std=0.05))
        net.add_unit(SoftmaxUnit("loss", "fc", "label", "prob"))
        net.fuse_conv_twins = twins
        rs = np.random.RandomState(0)
        x = rs.normal(0, 1, (8, 32, 14, 14)).astype(np.float32)
        onehot = np.zeros((8, 5), np.float32)
        onehot[np.arange(8), rs.randint(0, 5, 8)] = 1
        du = net.get_data_unit()
        du.data, du.label = gpu_owl.from_numpy(x), gpu_owl.from_numpy(onehot)
        net.batch_size = 8
        net.forward("TRAIN")
        made = []
        orig = gpu_owl.NArray.relu_back_tw

        def spy(diff, top, geo):
            out = orig(diff, top, geo)
            made.append(out._twin[1].value if out._twin is not None else 0)
            return out
        gpu_owl.NArray.relu_back_tw = staticmethod(spy)
        try:
            net.backward("TRAIN")
        finally:
            gpu_owl.NArray.relu_back_tw = staticmethod(orig)
        assert (len(made) == 3) == twins
        if twins:
            assert 3 in made          # at least one convolution here takes a top_diff twin (and its sums)
        res.append([g for uid in net.get_weighted_unit_ids() for g in (net.units[uid].weightgrad.to_numpy(), net.units[uid].biasgrad.to_numpy())])
    for a, b in zip(*res):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("builder,shape,batch", [("build_lenet", [28, 28, 1], 16), ("build_mnist_mlp", [784], 16)])
def test_small_configs_train(gpu_owl, builder, shape, batch):
    """configs[0..1] of BASELINE.json at reduced batch: loss goes down on a fixed synthetic batch."""
    import minerva_b200.owl.net as onet
    gpu_owl.set_seed(1)
    net = getattr(onet, builder)()
    rs = np.random.RandomState(0)
    x = rs.uniform(0, 1, [batch] + list(reversed(shape))).astype(np.float32)
    lab = rs.randint(0, 10, batch)
    onehot = np.zeros((batch, 10), np.float32)
    onehot[np.arange(batch), lab] = 1
    du = net.get_data_unit()
    du.data, du.label = gpu_owl.from_numpy(x), gpu_owl.from_numpy(onehot)
    net.batch_size = batch
    net.base_lr = net.current_lr = 0.1
    tr = onet.NetTrainer(net, None)
    tr.step()
    l0 = net.get_loss_units()[0].getloss()
    for _ in range(30):
        tr.step()
    l1 = net.get_loss_units()[0].getloss()
    assert np.isfinite(l0) and np.isfinite(l1) and l1 < 0.7 * l0, (l0, l1)


def test_googlenet_graph_runs(gpu_owl):
    """BASELINE config 5 (GoogLeNet, 3 loss heads, inception concat/slice, avg pools) at reduced batch: one
    full training step runs, every weighted unit gets finite gradients, and the loss is ~ln(1000) at init."""
    import minerva_b200.owl.net as onet
    gpu_owl.set_seed(3)
    net = onet.build_googlenet()
    batch = 4
    rs = np.random.RandomState(0)
    x = rs.standard_normal((batch, 3, 224, 224)).astype(np.float32)
    onehot = np.zeros((batch, 1000), np.float32)
    onehot[np.arange(batch), rs.randint(0, 1000, batch)] = 1
    du = net.get_data_unit()
    du.data, du.label = gpu_owl.from_numpy(x), gpu_owl.from_numpy(onehot)
    net.batch_size = batch
    tr = onet.NetTrainer(net, None)
    net.forward("TRAIN")
    net.backward("TRAIN")
    for uid in net.get_weighted_unit_ids():
        u = net.units[uid]
        g = u.weightgrad.to_numpy()
        assert np.isfinite(g).all() and np.abs(g).max() > 0, u.name
    losses = [lu.getloss() for lu in net.get_loss_units()]
    assert len(losses) == 3 and all(np.isfinite(l) and np.log(1000) - 1.0 < l < np.log(1000) + 3.0 for l in losses), losses
    tr.step()


def test_image_transform_matches_host_data_layer(gpu_owl):
    """mnv_image_transform_u8 == the reference data layer's host computation (owl/owl/net/netio.py:300-311):
    (uint8 image - mean image)[crop window][mirrored] as float32, bit for bit."""
    import torch
    from tests import gpu_util as g
    rs = np.random.RandomState(4)
    for (N, C, S, crop, scale, use_mean, use_off) in ((5, 3, 40, 33, 1.0, True, True), (3, 1, 28, 28, 1.0 / 255, False, False),
                                                      (4, 3, 256, 227, 0.017, True, True), (2, 2, 9, 7, 1.0, True, False)):
        img = rs.randint(0, 256, (N, C, S, S), dtype=np.uint8)
        mean = (rs.uniform(90, 130, (C, S, S))).astype(np.float32)
        off = np.stack([rs.randint(0, S - crop + 1, N), rs.randint(0, S - crop + 1, N), rs.randint(0, 2, N)], 1).astype(np.int32)
        want = np.empty((N, C, crop, crop), np.float32)
        for n in range(N):
            oy, ox, mir = off[n] if use_off else (0, 0, 0)
            im = img[n].astype(np.float32) - (mean if use_mean else np.float32(0))        # netio.py:300
            w = im[:, oy:oy + crop, ox:ox + crop]
            want[n] = (w[:, :, ::-1] if mir else w) * np.float32(scale)
        dst = torch.full((N, C, crop, crop), float("nan"), device="cuda")
        g.run("mnv_image_transform_u8", torch.from_numpy(img).cuda(), torch.from_numpy(mean).cuda() if use_mean else 0,
              torch.from_numpy(off).cuda() if use_off else 0, dst, N, C, S, S, crop, crop, float(scale))
        g.assert_bits_equal(dst.cpu().numpy(), want, "image transform %r" % ((N, C, S, crop),))


def test_feed_data_unit_trains(gpu_owl):
    """A net whose data unit is a HostFeed over a provider (uint8 stored images, random crop + mirror on the device,
    double-buffered uploads on a copy stream): steps run, every step sees a freshly uploaded batch, the loss is finite."""
    import torch
    import minerva_b200.owl.net as onet
    from minerva_b200.owl import _runtime as rt
    from minerva_b200.owl.net.data import HostFeed, FeedDataUnit, SyntheticImageProvider
    from tests.test_net_cpu import _tiny_net
    from minerva_b200.owl.net.net import _default_backend
    gpu_owl.set_seed(2)
    net = _tiny_net(_default_backend())
    prov = SyntheticImageProvider(8, 3, (21, 21), 5, seed=1, pool=3)
    feed = HostFeed(gpu_owl, rt, provider=prov, mean=[100.0, 110.0, 120.0], scale=1.0 / 64, crop=(17, 17), mirror=True)
    du = FeedDataUnit("data", ["data", "label"], feed)
    du.B = net.B
    net.units[0] = du
    net.batch_size = 8
    tr = onet.NetTrainer(net, None)
    seen = []
    for _ in range(5):
        tr.step()
        seen.append(du.data.to_numpy().copy())
        assert du.data.shape == [17, 17, 3, 8] and np.isfinite(net.get_loss_units()[0].getloss())
    du.close()
    assert not np.array_equal(seen[0], seen[1])                      # a new minibatch every step
    lo, hi = (0 - 120.0) / 64, (255 - 100.0) / 64
    assert all(s.min() >= lo and s.max() <= hi for s in seen)


def test_prototxt_alexnet_trains_from_the_feed(gpu_owl):
    """models/bvlc_alexnet_nogroup_{solver,train_val}.prototxt through CaffeNetBuilder: the Data layer becomes a
    FeedDataUnit (synthetic 256 x 256 uint8 images, random 227 crop + mirror + mean on the device), in-place ReLU / Dropout
    layers, fused conv+ReLU planned over the in-place names; two trainer steps at batch 4."""
    import os
    import minerva_b200.owl.net as onet
    from minerva_b200.owl.net import net as N
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gpu_owl.set_seed(9)
    cb = onet.CaffeNetBuilder(os.path.join(root, "models", "bvlc_alexnet_nogroup_solver.prototxt"))
    net = cb.build_net(N.Net(), num_gpu=64)                     # 256 / 64 = 4 images per replica
    du = net.get_data_unit()
    assert du.geometry["batch"] == 4 and du.geometry["crop"] == (227, 227) and du.geometry["mirror"]
    du.attach(1000, seed=1)
    net.batch_size = 4
    tr = onet.NetTrainer(net, None)
    tr.step()
    tr.step()
    assert sum(1 for u in net.units if isinstance(u, N.ConvConnection) and u.fuse_relu) == 5
    assert du.data.shape == [227, 227, 3, 4]
    loss = net.get_loss_units()[0].getloss()
    assert np.isfinite(loss) and 5.0 < loss < 9.0, loss
    du.close()
