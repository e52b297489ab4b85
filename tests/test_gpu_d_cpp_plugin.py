"""-m gpu: the C++ drop-in surface (minerva/op ComputeFn::Execute + minerva/device GpuDevice) driven by a C++
test program, tests/cpp/test_host_plugin.cpp, built by minerva_b200.build.build_host()."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "minerva_b200", "lib", "test_host_plugin")


@pytest.mark.gpu
def test_cpp_plugin_surface():
    if not os.path.exists(BIN):
        from minerva_b200 import build
        build.build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PASS" in out.stdout


APPS = os.path.join(ROOT, "minerva_b200", "lib", "mnist_apps")


def _run_app(*args):
    import json
    out = subprocess.run([APPS] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("net,builder", [("lenet", "build_lenet"), ("mlp", "build_mnist_mlp")])
def test_cpp_apps_match_the_owl_net_twin(tmp_path, net, builder):
    """configs[0..1] through the C++ host path (Task -> StreamDevice::PushTask -> ComputeFn::Execute, event-ordered,
    size-class pool) vs the same net in owl.net on the same initial parameters and batch: after 6 plain-SGD steps the
    parameters are BIT-IDENTICAL (the owl side runs its fused kernels, which are bit-identical to the op chains), in all
    three completion modes; the steady state allocates nothing."""
    import numpy as np
    import minerva_b200.owl as owl
    import minerva_b200.owl.net as onet
    steps = 6
    finals = {}
    for mode in ("enqueue", "event", "blocking"):
        d = tmp_path / mode
        d.mkdir()
        r = _run_app("--net", net, "--mb", 32, "--steps", steps - 2, "--warmup", 2, "--completion", mode, "--dump-dir", d, "--seed", 5)
        assert r["cuda_mallocs_in_timed_region"] == 0 and r["pool_hits_in_timed_region"] > 0, r
        assert r["listener_completions"] >= r["ops_per_step"] * steps
        assert np.isfinite(r["loss_last"]) and r["loss_last"] < r["loss_first"], r
        finals[mode] = d
    owl.set_device(owl.create_gpu_device(0))
    d = finals["enqueue"]
    rd = lambda name: np.fromfile(d / (name + ".dat"), np.float32)      # noqa: E731
    g = getattr(onet, builder)()
    wu = [g.units[i] for i in g.get_weighted_unit_ids()]
    shape = [28, 28, 1] if net == "lenet" else [784]
    du = g.get_data_unit()
    du.data = owl.from_numpy(rd("data").reshape([32] + list(reversed(shape))))
    du.label = owl.from_numpy(rd("label").reshape(32, 10))
    g.batch_size = 32
    g.forward("TRAIN")                                                 # shapes known; now overwrite the fillers' values
    for i, u in enumerate(wu):
        u.weight = owl.from_numpy(rd("w%d_init" % i).reshape(list(reversed(u.wshape))))
        u.bias = owl.from_numpy(rd("b%d_init" % i).reshape(list(reversed(u.bshape))))
        u.lr_mult_b = 1.0                                              # the apps use one alpha for weights and biases
    tr = onet.NetTrainer(g, None)
    for _ in range(steps):
        tr.step()
    for i, u in enumerate(wu):
        for mode, dd in finals.items():
            np.testing.assert_array_equal(u.weight.to_numpy().ravel(), np.fromfile(dd / ("w%d_final.dat" % i), np.float32),
                                          err_msg="%s w%d %s" % (net, i, mode))
            np.testing.assert_array_equal(u.bias.to_numpy().ravel(), np.fromfile(dd / ("b%d_final.dat" % i), np.float32),
                                          err_msg="%s b%d %s" % (net, i, mode))


def test_cpp_apps_refuse_to_run_without_a_gpu():
    """No CPU fallback on the product path: without a CUDA device the C++ app exits with an error, it does not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    from minerva_b200 import build
    build.build()
    out = subprocess.run([APPS, "--net", "mlp"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 3 and "no CUDA device" in out.stderr


def test_size_classes():
    """StreamDevice::SizeClass: 256 B floor, 8 classes per power of two (<= 12.5 % slack), monotone."""
    import ctypes
    from minerva_b200 import build
    build.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "minerva_b200", "lib", "libminerva_b200_host.so"))
    fn = lib.mnv_host_size_class
    fn.restype = ctypes.c_size_t
    fn.argtypes = [ctypes.c_size_t]
    assert fn(1) == 256 and fn(256) == 256 and fn(257) == 288 and fn(4096) == 4096 and fn(4097) == 4608
    prev = 0
    for b in list(range(1, 5000, 7)) + [10 ** 6 + 3, 151 * 2 ** 20 + 1, 3 * 2 ** 30]:
        c = fn(b)
        assert c >= b and c <= max(256, b + b // 8 + 1) and c >= prev
        prev = c
