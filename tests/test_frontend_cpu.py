"""No-GPU tests of the Caffe front end: prototxt text parser, CaffeNetBuilder over the model files under
models/, raw-fp32 snapshots, and the .caffemodel converter (reference: owl/owl/net/net_helper.py:11-317)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prototxt_parser_round_trip():
    from minerva_b200.owl.net import prototxt
    text = '''
    name: "t"   # a comment
    input_dim: [1, 3, 8, 8]
    layers { name: 'c' type: CONVOLUTION bottom: "data" top: "c" blobs_lr: 1 blobs_lr: 2
             convolution_param { num_output: 4 kernel_size: 3 weight_filler { type: "gaussian" std: 1e-2 } } }
    layer { name: "r" type: "ReLU" bottom: "c" top: "c" include { phase: TEST } }
    flag: true  ratio: -0.5  hex: 0x10
    '''
    m = prototxt.parse(text)
    assert m.get("name") == "t" and m.all("input_dim") == [1, 3, 8, 8] and m.get("flag") is True
    assert m.get("ratio") == -0.5 and m.get("hex") == 16
    c = m.get("layers")
    assert c.get("type") == "CONVOLUTION" and isinstance(c.get("type"), prototxt.Enum) and c.all("blobs_lr") == [1, 2]
    assert c.sub("convolution_param").sub("weight_filler").get("std") == 0.01
    assert c.sub("pooling_param").get("pad", 7) == 7                     # absent block reads as defaults
    assert m.get("layer").sub("include").get("phase") == "TEST"
    again = prototxt.parse(prototxt.dump(m))
    assert prototxt.dump(again) == prototxt.dump(m)
    with pytest.raises(ValueError):
        prototxt.parse("layer { name: 'x' ")


@pytest.mark.parametrize("stem,builder,stored", [("bvlc_alexnet_nogroup", "build_alexnet", None), ("bvlc_googlenet", "build_googlenet", None),
                                                 ("mnist_lenet", "build_lenet", (1, 28, 28)), ("mnist_mlp", "build_mnist_mlp", (1, 28, 28))])
def test_prototxt_nets_equal_the_programmatic_builders(stem, builder, stored):
    """The unit list built from models/*.prototxt has the types, geometry, fillers and multipliers of the builder's."""
    import minerva_b200.owl.net as onet
    from minerva_b200.owl.net.net_helper import CaffeNetBuilder
    from minerva_b200.owl.net import net as N

    class _B:
        owl = co = ele = None
    ref = getattr(onet, builder)(_B())
    cb = CaffeNetBuilder(os.path.join(ROOT, "models", stem + "_solver.prototxt"))
    got = cb.build_net(N.Net(_B()), stored_shape=stored, feed=False)
    assert [type(u) for u in got.units] == [type(u) for u in ref.units]
    assert [u.name for u in got.units] == [u.name for u in ref.units]
    assert (got.base_lr, got.momentum, got.base_weight_decay) == (ref.base_lr, ref.momentum, ref.base_weight_decay)
    assert int(np.prod(got.input_shape)) == int(np.prod(ref.input_shape))     # the MLP's 784-vector is a 1 x 28 x 28 image in the file
    for a, b in zip(got.units, ref.units):
        for attr in ("num_output", "kernel_size", "stride", "pad", "geom", "pool", "args", "keep_ratio", "loss_weight", "lr_mult_w",
                     "lr_mult_b", "decay_mult_w", "decay_mult_b", "weight_std", "bias_value", "weight_filler", "need_bp"):
            if hasattr(b, attr):
                if attr == "weight_std" and b.weight_filler == "xavier":
                    continue                                  # unused by the xavier filler
                assert getattr(a, attr) == pytest.approx(getattr(b, attr)), (a.name, attr)
        assert len(a.btm_names) == len(b.btm_names) and len(a.top_names) == len(b.top_names)
    # in-place ReLU / Dropout layers keep their bottom's name
    relus = [u for u in got.units if isinstance(u, N.ReluUnit)]
    assert relus and all(u.top_names == u.btm_names for u in relus)


def test_prototxt_alexnet_step_equals_builder_step():
    """Same initial weights (same generator order), same data -> the in-place prototxt graph and the distinct-name builder
    graph give bit-identical losses and gradients on the CPU oracle."""
    import minerva_b200.owl.net as onet
    from minerva_b200.owl.net.net_helper import CaffeNetBuilder
    from minerva_b200.owl.net import net as N
    from oracle import owl_cpu
    rs = np.random.RandomState(2)
    x = rs.standard_normal((1, 3, 227, 227)).astype(np.float32)
    lab = np.zeros((1, 1000), np.float32)
    lab[0, 7] = 1
    nets = []
    for how in ("builder", "prototxt"):
        B = owl_cpu.Backend()
        owl_cpu.set_seed(6)
        if how == "builder":
            net = onet.build_alexnet(B)
        else:
            net = CaffeNetBuilder(os.path.join(ROOT, "models", "bvlc_alexnet_nogroup_solver.prototxt")).build_net(N.Net(B), feed=False)
        du = net.get_data_unit()
        du.data, du.label = B.owl.from_numpy(x), B.owl.from_numpy(lab)
        net.batch_size = 1
        net.forward("TRAIN")
        net.backward("TRAIN")
        nets.append(net)
    a, b = nets
    assert a.get_loss_units()[0].getloss() == b.get_loss_units()[0].getloss()
    for ua, ub in zip([a.units[i] for i in a.get_weighted_unit_ids()], [b.units[i] for i in b.get_weighted_unit_ids()]):
        np.testing.assert_array_equal(ua.weightgrad.a, ub.weightgrad.a, err_msg=ua.name)
        np.testing.assert_array_equal(ua.biasgrad.a, ub.biasgrad.a, err_msg=ua.name)


def test_snapshot_round_trip(tmp_path):
    """save_net_to_file / init_net_from_file: raw fp32 .dat per tensor under snapshot<idx>/, '/' in layer names -> '_';
    a missing or mis-sized file leaves the filler-initialised tensor and is reported."""
    from minerva_b200.owl.net.net_helper import CaffeNetBuilder
    from tests.test_net_cpu import _tiny_net, _batch
    from oracle import owl_cpu
    cb = CaffeNetBuilder(os.path.join(ROOT, "models", "mnist_mlp_solver.prototxt"))

    def trained(seed, steps):
        B = owl_cpu.Backend()
        owl_cpu.set_seed(seed)
        net = _tiny_net(B)
        net.units[1].name = "scope/conv1"
        net.name_to_uid["scope/conv1"] = 1
        du = net.get_data_unit()
        du.data, du.label = _batch(B, 4)
        net.batch_size = 4
        for _ in range(steps):
            net.forward("TEST"); net.backward("TEST"); net.weight_update()
        net.forward("TEST")
        return net
    a = trained(1, 2)
    cb.save_net_to_file(a, str(tmp_path), 3)
    files = sorted(os.listdir(tmp_path / "snapshot3"))
    assert "scope_conv1_weights.dat" in files and "fc8_biasdelta.dat" in files and len(files) == 4 * len(a.get_weighted_unit_ids())
    b = trained(2, 0)                                       # different weights
    assert cb.init_net_from_file(b, str(tmp_path), 3) == []
    for i in a.get_weighted_unit_ids():
        for attr in ("weight", "weightdelta", "bias", "biasdelta"):
            np.testing.assert_array_equal(getattr(a.units[i], attr).a, getattr(b.units[i], attr).a)
            assert getattr(a.units[i], attr).shape == getattr(b.units[i], attr).shape
    b.forward("TEST")
    assert b.get_loss_units()[0].getloss() == a.get_loss_units()[0].getloss()
    # a truncated file and a missing file are reported, the tensors keep their values
    os.remove(tmp_path / "snapshot3" / "fc8_weights.dat")
    with open(tmp_path / "snapshot3" / "fc6_bias.dat", "wb") as f:
        f.write(b"\0" * 8)
    c = trained(2, 0)
    keep = c.units[c.name_to_uid["fc8"]].weight.a.copy()
    assert sorted(cb.init_net_from_file(c, str(tmp_path), 3)) == [("fc6", "bias"), ("fc8", "weights")]
    np.testing.assert_array_equal(c.units[c.name_to_uid["fc8"]].weight.a, keep)


@pytest.mark.parametrize("v1", [False, True])
def test_caffemodel_converter(tmp_path, v1):
    """CaffeModelLoader: filters rotated by 180 degrees per (co, ci) plane, inner-product weights transposed
    (net_helper.py:296-313); both the V1 `layers` and the new `layer` encodings of a .caffemodel."""
    from minerva_b200.owl.net.caffemodel import write_caffemodel, read_caffemodel
    from minerva_b200.owl.net.net_helper import CaffeModelLoader
    from oracle import pyoracle as orc
    rs = np.random.RandomState(8)
    cw, cb_ = rs.standard_normal((6, 3, 3, 5)).astype(np.float32), rs.standard_normal(6).astype(np.float32)
    fw, fb = rs.standard_normal((4, 10)).astype(np.float32), rs.standard_normal(4).astype(np.float32)
    path = str(tmp_path / "m.caffemodel")
    write_caffemodel(path, [{"name": "data", "type": "Data", "blobs": []},
                            {"name": "incep/conv", "type": "Convolution", "blobs": [cw, cb_]},
                            {"name": "fc", "type": "InnerProduct", "blobs": [fw.reshape(1, 1, 4, 10) if v1 else fw, fb]}], v1=v1)
    layers = read_caffemodel(path)
    assert [l["name"] for l in layers] == ["data", "incep/conv", "fc"] and layers[1]["type"] == "Convolution"
    np.testing.assert_array_equal(layers[1]["blobs"][0]["data"], cw.ravel())
    conv = CaffeModelLoader(path, str(tmp_path / "w"), 0)
    assert conv.converted == ["incep/conv", "fc"]
    d = tmp_path / "w" / "snapshot0"
    filt = np.fromfile(d / "incep_conv_weights.dat", np.float32).reshape(6, 3, 3, 5)
    np.testing.assert_array_equal(filt, cw[:, :, ::-1, ::-1])
    np.testing.assert_array_equal(np.fromfile(d / "fc_weights.dat", np.float32).reshape(10, 4), fw.T)
    np.testing.assert_array_equal(np.fromfile(d / "fc_bias.dat", np.float32), fb)
    # semantics: Minerva's convolution with the rotated filter == Caffe's cross-correlation with the original one
    x = rs.standard_normal((1, 3, 7, 9)).astype(np.float32)
    y = orc.conv_forward(x.ravel(), filt.ravel(), cb_, 1, 3, 6, 7, 9, 0, 0, 1, 1, 3, 5).reshape(6, 5, 5)
    want = np.zeros((6, 5, 5))
    for co in range(6):
        for i in range(5):
            for j in range(5):
                want[co, i, j] = (x[0, :, i:i + 3, j:j + 5].astype(np.float64) * cw[co]).sum() + cb_[co]
    np.testing.assert_allclose(y, want, rtol=1e-4, atol=1e-4)
    # owl's {num_output, input_dim} matrix is column-major: W * a == Caffe's fw @ a
    a = rs.standard_normal(10).astype(np.float32)
    w_owl = np.fromfile(d / "fc_weights.dat", np.float32)
    np.testing.assert_allclose(orc.matmult(w_owl, a, 4, 1, 10), fw @ a, rtol=1e-5, atol=1e-5)
