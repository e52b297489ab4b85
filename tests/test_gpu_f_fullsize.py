"""-m gpu parity at BASELINE.json's FULL sizes: every AlexNet layer at batch 256 and the GoogLeNet geometries at batch 120
run on the GPU at full size; the CPU oracle checks a strided sub-batch (images {0, 17, N-1}), which it finishes in seconds.

Why a sub-batch is a full check of the index math: convolution forward / backward-data, pooling and LRN treat images
independently, so image n of the full-size result must equal the oracle on image n alone.  Backward-filter / -bias sum
over the batch: top_diff is zero outside the selected images, so the full-size result (whose K loop still walks all N
images, every tile, every split) must equal the oracle on the selected images.

Tolerances (north_star): bit-exact for max pooling, <= 1e-5 relative for LRN / average pooling / bias sums,
<= 5e-3 norm-relative for the TF32 convolutions."""
import numpy as np
import pytest
import torch

from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu
TOL = 5e-3


@pytest.fixture(scope="module")
def g():
    from tests import gpu_util
    return gpu_util


def _sel(N):
    return [0, 17 % N, N - 1] if N > 2 else list(range(N))


def _rand(shape, seed, relu=False):
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    t = torch.randn(shape, generator=gen, device="cuda", dtype=torch.float32)
    return torch.relu(t) if relu else t


ALEX_CONVS = [
    # name, (N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw)
    ("alexnet conv1", (256, 3, 96, 227, 227, 0, 0, 4, 4, 11, 11)),
    ("alexnet conv2", (256, 96, 256, 27, 27, 2, 2, 1, 1, 5, 5)),
    ("alexnet conv3", (256, 256, 384, 13, 13, 1, 1, 1, 1, 3, 3)),
    ("alexnet conv4", (256, 384, 384, 13, 13, 1, 1, 1, 1, 3, 3)),
    ("alexnet conv5", (256, 384, 256, 13, 13, 1, 1, 1, 1, 3, 3)),
]
GOOG_CONVS = [
    ("googlenet conv1/7x7_s2", (120, 3, 64, 224, 224, 3, 3, 2, 2, 7, 7)),
    ("googlenet conv2/3x3_reduce", (120, 64, 64, 56, 56, 0, 0, 1, 1, 1, 1)),
    ("googlenet conv2/3x3", (120, 64, 192, 56, 56, 1, 1, 1, 1, 3, 3)),
    ("googlenet 3a/5x5_reduce", (120, 192, 16, 28, 28, 0, 0, 1, 1, 1, 1)),
    ("googlenet 3a/5x5", (120, 16, 32, 28, 28, 2, 2, 1, 1, 5, 5)),
    ("googlenet 3b/3x3", (120, 128, 192, 28, 28, 1, 1, 1, 1, 3, 3)),
    ("googlenet 4a/1x1", (120, 480, 192, 14, 14, 0, 0, 1, 1, 1, 1)),
    ("googlenet 4e/3x3", (120, 160, 320, 14, 14, 1, 1, 1, 1, 3, 3)),
    ("googlenet 5b/1x1 (Ci 832)", (120, 832, 384, 7, 7, 0, 0, 1, 1, 1, 1)),
    ("googlenet 5b/3x3", (120, 192, 384, 7, 7, 1, 1, 1, 1, 3, 3)),
    ("googlenet 5b/5x5", (120, 48, 128, 7, 7, 2, 2, 1, 1, 5, 5)),
    ("googlenet loss1/conv", (120, 512, 128, 4, 4, 0, 0, 1, 1, 1, 1)),
]
LENET_CONVS = [
    ("lenet conv1", (256, 1, 16, 28, 28, 0, 0, 1, 1, 5, 5)),
    ("lenet conv2", (256, 16, 32, 12, 12, 2, 2, 1, 1, 5, 5)),
]


@pytest.mark.parametrize("name,case", ALEX_CONVS + GOOG_CONVS + LENET_CONVS, ids=lambda v: v if isinstance(v, str) else "")
def test_conv_full_size_three_directions(g, name, case):
    N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw = case
    Ho, Wo = orc.conv_out(H, ph, fh, sv), orc.conv_out(W, pw, fw, sh)
    sel = _sel(N)
    sub = (len(sel),) + tuple(case[1:])
    ws = g.workspace()
    x = _rand((N, Ci, H, W), 1)
    w = _rand((Co, Ci, fh, fw), 2) * float(1.0 / np.sqrt(Ci * fh * fw))
    b = _rand((Co,), 3)
    xs, wh, bh = g.host(x[sel]).ravel(), g.host(w).ravel(), g.host(b)
    # forward (+ the fused ReLU epilogue owl.net uses)
    y = torch.full((N, Co, Ho, Wo), float("nan"), device="cuda")
    g.run("mnv_conv_forward", x, w, b, y, *case, ws, ws.numel())
    want = orc.conv_forward(xs, wh, bh, *sub)
    assert g.norm_rel(g.host(y[sel]).ravel(), want) < TOL, name + " forward"
    assert torch.isfinite(y).all()
    yr = torch.full((N, Co, Ho, Wo), float("nan"), device="cuda")
    g.run("mnv_conv_forward_relu", x, w, b, yr, *case, ws, ws.numel())
    assert torch.equal(yr, torch.relu(y)), name + " fused ReLU epilogue"
    del yr
    # backward data: per image
    dy = _rand((N, Co, Ho, Wo), 4)
    dx = torch.full((N, Ci, H, W), float("nan"), device="cuda")
    g.run("mnv_conv_backward_data", dy, w, dx, *case, ws, ws.numel())
    want = orc.conv_backward_data(g.host(dy[sel]).ravel(), wh, *sub)
    assert g.norm_rel(g.host(dx[sel]).ravel(), want) < TOL, name + " backward data"
    assert torch.isfinite(dx).all()
    del dx
    # backward filter + bias: top_diff zero outside the selected images, the K loop still spans the whole batch
    dyz = torch.zeros_like(dy)
    dyz[sel] = dy[sel]
    dw = torch.full((Co, Ci, fh, fw), float("nan"), device="cuda")
    db = torch.full((Co,), float("nan"), device="cuda")
    g.run("mnv_conv_backward_filter_bias", x, dyz, dw, db, *case, ws, ws.numel())
    dys = g.host(dy[sel]).ravel()
    want = orc.conv_backward_filter(xs, dys, *sub)
    assert g.norm_rel(g.host(dw).ravel(), want) < TOL, name + " backward filter"
    wb = orc.conv_backward_bias(dys, len(sel), Co, Ho, Wo)
    assert np.abs(g.host(db) - wb).max() <= 1e-5 * np.abs(dys).reshape(len(sel), Co, -1).sum((0, 2)).max(), name + " bias"
    dw2 = torch.full((Co, Ci, fh, fw), float("nan"), device="cuda")
    g.run("mnv_conv_backward_filter", x, dyz, dw2, *case, ws, ws.numel())
    assert torch.equal(dw, dw2), name + " filter gradient with and without the fused bias sums"
    # and with a dense top_diff the result stays finite and deterministic (split-K order is fixed)
    g.run("mnv_conv_backward_filter", x, dy, dw, *case, ws, ws.numel())
    g.run("mnv_conv_backward_filter", x, dy, dw2, *case, ws, ws.numel())
    assert torch.isfinite(dw).all() and torch.equal(dw, dw2)


POOLS = [
    # name, kind, (N, C, H, W, sv, sh, wh, ww, ph, pw)
    ("alexnet pool1", "max", (256, 96, 55, 55, 2, 2, 3, 3, 0, 0)),
    ("alexnet pool2", "max", (256, 256, 27, 27, 2, 2, 3, 3, 0, 0)),
    ("alexnet pool5", "max", (256, 256, 13, 13, 2, 2, 3, 3, 0, 0)),
    ("lenet pool1", "max", (256, 16, 24, 24, 2, 2, 2, 2, 0, 0)),
    ("lenet pool2", "max", (256, 32, 12, 12, 3, 3, 3, 3, 0, 0)),
    ("googlenet pool1 (overhang)", "max", (120, 64, 112, 112, 2, 2, 3, 3, 0, 0)),
    ("googlenet pool2 (overhang)", "max", (120, 192, 56, 56, 2, 2, 3, 3, 0, 0)),
    ("googlenet pool3 (overhang)", "max", (120, 480, 28, 28, 2, 2, 3, 3, 0, 0)),
    ("googlenet pool4 (overhang)", "max", (120, 832, 14, 14, 2, 2, 3, 3, 0, 0)),
    ("googlenet 3a/pool 3x3/1 pad 1", "max", (120, 192, 28, 28, 1, 1, 3, 3, 1, 1)),
    ("googlenet 4a/pool 3x3/1 pad 1", "max", (120, 480, 14, 14, 1, 1, 3, 3, 1, 1)),
    ("googlenet 5a/pool 3x3/1 pad 1", "max", (120, 832, 7, 7, 1, 1, 3, 3, 1, 1)),
    ("googlenet loss1/ave_pool 5x5/3", "average", (120, 512, 14, 14, 3, 3, 5, 5, 0, 0)),
    ("googlenet pool5 7x7/1", "average", (120, 1024, 7, 7, 1, 1, 7, 7, 0, 0)),
]


@pytest.mark.parametrize("name,kind,case", POOLS, ids=lambda v: v if isinstance(v, str) and " " in v else "")
def test_pooling_full_size(g, name, kind, case):
    N, C, H, W, sv, sh, wh, ww, ph, pw = case
    Ho, Wo = orc.pooled_size(H, ph, wh, sv), orc.pooled_size(W, pw, ww, sh)
    sel = _sel(N)
    sub = (len(sel),) + tuple(case[1:])
    x = _rand((N, C, H, W), 5, relu=True)          # post-ReLU input: ~50 % exact zeros => ties
    y = torch.full((N, C, Ho, Wo), float("nan"), device="cuda")
    g.run("mnv_%s_pooling_forward" % kind, x, y, *case)
    xs = g.host(x[sel]).ravel()
    wy = getattr(orc, kind + "_pooling_forward")(xs, *sub)
    dy = _rand((N, C, Ho, Wo), 6)
    dx = torch.full((N, C, H, W), float("nan"), device="cuda")
    g.run("mnv_%s_pooling_backward" % kind, x, y, dy, dx, *case)
    wdx = getattr(orc, kind + "_pooling_backward")(xs, wy, g.host(dy[sel]).ravel(), *sub)
    if kind == "max":
        g.assert_bits_equal(g.host(y[sel]).ravel(), wy, name + " forward")
        g.assert_bits_equal(g.host(dx[sel]).ravel(), wdx, name + " backward")
        # the fused ReLU-backward variant and the arg-max-byte pair owl.net actually runs give the same bits
        dxr = torch.full((N, C, H, W), float("nan"), device="cuda")
        g.run("mnv_max_pooling_backward_relu", x, y, dy, dxr, *case)
        assert torch.equal(dxr, torch.where(x > 0, dx, torch.zeros_like(dx))), name + " backward + ReLU mask"
        from minerva_b200 import _lib
        if _lib.load().mnv_max_pooling_idx_supported(*case):
            idx = torch.empty(N * C * Ho * Wo, dtype=torch.uint8, device="cuda")
            y2 = torch.full((N, C, Ho, Wo), float("nan"), device="cuda")
            g.run("mnv_max_pooling_forward_idx", x, y2, idx, *case)
            assert torch.equal(y2, y), name + " forward with arg-max bytes"
            g.run("mnv_max_pooling_backward_idx", dy, idx, 0, dxr, *case)
            assert torch.equal(dxr, dx), name + " backward from arg-max bytes"
            g.run("mnv_max_pooling_backward_idx", dy, idx, y, dxr, *case)
            assert torch.equal(dxr, torch.where(x > 0, dx, torch.zeros_like(dx))), name + " backward from bytes + ReLU"
    else:
        assert np.abs(g.host(y[sel]).ravel() - wy).max() <= 1e-5 * np.abs(wy).max(), name + " forward"
        assert np.abs(g.host(dx[sel]).ravel() - wdx).max() <= 1e-5 * np.abs(wdx).max(), name + " backward"
    assert torch.isfinite(dx).all()


LRNS = [("alexnet norm1", (256, 96, 55, 55)), ("alexnet norm2", (256, 256, 27, 27)),
        ("googlenet pool1/norm1", (120, 64, 56, 56)), ("googlenet conv2/norm2", (120, 192, 56, 56))]


@pytest.mark.parametrize("name,shape", LRNS, ids=lambda v: v if isinstance(v, str) else "")
def test_lrn_full_size(g, name, shape):
    N, C, H, W = shape
    size, alpha, beta = 5, 1e-4, 0.75
    sel = _sel(N)
    x = _rand(shape, 7, relu=True) * 30.0           # large enough that scale departs from 1
    dy = _rand(shape, 8)
    y, sc, dx = (torch.full(shape, float("nan"), device="cuda") for _ in range(3))
    g.run("mnv_lrn_forward", x, sc, y, size, alpha, beta, N, C, W, H)
    g.run("mnv_lrn_backward", x, y, sc, dy, dx, size, alpha, beta, N, C, W, H)
    xs = g.host(x[sel]).ravel()
    wy, wsc = orc.lrn_forward(xs, size, alpha, beta, len(sel), C, W, H)
    wdx = orc.lrn_backward(xs, wy, wsc, g.host(dy[sel]).ravel(), size, alpha, beta, len(sel), C, W, H)
    g.assert_bits_equal(g.host(sc[sel]).ravel(), wsc, name + " scale")
    assert np.abs(g.host(y[sel]).ravel() - wy).max() <= 1e-5 * np.abs(wy).max(), name + " forward"
    assert np.abs(g.host(dx[sel]).ravel() - wdx).max() <= 1e-5 * np.abs(wdx).max(), name + " backward"
    # the scale-less pair owl.net runs: forward bit-identical to the three-array form's output, backward to its result
    y2, dx2 = (torch.full(shape, float("nan"), device="cuda") for _ in range(2))
    g.run("mnv_lrn_forward_lite", x, y2, size, alpha, beta, N, C, W, H)
    assert torch.equal(y2, y), name + " lite forward"
    g.run("mnv_lrn_backward_lite", x, dy, dx2, size, alpha, beta, N, C, W, H, 0)
    assert torch.equal(dx2, dx), name + " lite backward"
    g.run("mnv_lrn_backward_lite", x, dy, dx2, size, alpha, beta, N, C, W, H, 1)
    assert torch.equal(dx2, torch.where(x > 0, dx, torch.zeros_like(dx))), name + " lite backward + ReLU mask"


def test_concat_slice_full_size(gpu_owl_f):
    """GoogLeNet inception_3a output at batch 120: four branches concatenated on the channel dim and sliced back."""
    owl = gpu_owl_f
    rs = np.random.RandomState(3)
    chans = (64, 128, 32, 32)
    parts = [rs.standard_normal((120, c, 28, 28)).astype(np.float32) for c in chans]
    arrs = [owl.from_numpy(p) for p in parts]
    cat = owl.concat(arrs, 2)
    assert cat.shape == [28, 28, sum(chans), 120]
    np.testing.assert_array_equal(cat.to_numpy(), np.concatenate(parts, 1))
    off = 0
    for p, c in zip(parts, chans):
        np.testing.assert_array_equal(owl.slice(cat, 2, off, c).to_numpy(), p)
        off += c
    # all four slices in one launch (mnv_copy_strided_n), what ConcatUnit.backward uses; odd extents take the scalar path
    for piece, p in zip(owl.NArray.split(cat, 2, list(chans)), parts):
        np.testing.assert_array_equal(piece.to_numpy(), p)
    odd = [rs.standard_normal((3, c, 5, 7)).astype(np.float32) for c in (3, 1, 6)]
    cat2 = owl.concat([owl.from_numpy(p) for p in odd], 2)
    np.testing.assert_array_equal(cat2.to_numpy(), np.concatenate(odd, 1))
    for piece, p in zip(owl.NArray.split(cat2, 2, [3, 1, 6]), odd):
        np.testing.assert_array_equal(piece.to_numpy(), p)


@pytest.fixture(scope="module")
def gpu_owl_f():
    import minerva_b200.owl as owl
    owl.set_device(owl.create_gpu_device(0))
    return owl


# ---- whole training step of the other BASELINE configs against the CPU oracle -------------------------------------
# Two oracles per net.  (1) The reference's fp32 algorithm on TF32-ROUNDED operands for MatMult / convolution
# (owl_cpu.set_tf32_operands): what north_star's "TF32 conv and GEMM" computes; the GPU must match it to 5e-3 on the
# losses and on EVERY gradient.  (2) The plain fp32 oracle: losses to 5e-3; the gradients of the first layers differ by
# what operand rounding does to a tiny batch through ReLU / arg-max switches (LeNet conv1 3e-2, GoogLeNet conv1 9e-2 --
# the same numbers oracle (1) shows against oracle (2) on the CPU alone, asserted below), so against (2) the bound is
# "no worse than 1.5x the TF32-operand oracle's own distance".
def _step_nets(gpu_owl, builder, shape, classes, batch, seed, uniform=False):
    import minerva_b200.owl.net as onet
    from minerva_b200.owl.net.net import _default_backend
    from oracle import owl_cpu
    rs = np.random.RandomState(seed)
    x = (rs.uniform(0, 1, [batch] + list(reversed(shape))) if uniform
         else rs.standard_normal([batch] + list(reversed(shape)))).astype(np.float32)
    onehot = np.zeros((batch, classes), np.float32)
    onehot[np.arange(batch), rs.randint(0, classes, batch)] = 1
    nets = []
    for B, seed_fn, tf32 in ((owl_cpu.Backend(), owl_cpu.set_seed, False), (owl_cpu.Backend(), owl_cpu.set_seed, True),
                             (_default_backend(), gpu_owl.set_seed, False)):
        owl_cpu.set_tf32_operands(tf32)
        try:
            seed_fn(seed)
            net = getattr(onet, builder)(B)
            du = net.get_data_unit()
            du.data, du.label = B.owl.from_numpy(x), B.owl.from_numpy(onehot)
            net.batch_size = batch
            net.forward("TRAIN")
            net.backward("TRAIN")
        finally:
            owl_cpu.set_tf32_operands(False)
        nets.append(net)
    return nets


def _errs(a_net, b_net):
    out = {}
    for uid in a_net.get_weighted_unit_ids():
        for attr in ("weight", "weightgrad", "biasgrad"):
            a = getattr(a_net.units[uid], attr).to_numpy().astype(np.float64)
            b = getattr(b_net.units[uid], attr).to_numpy().astype(np.float64)
            out[a_net.units[uid].name + "." + attr] = float(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-12))
    return out


def _check_step(name, fp32, tf32, gpu, deep=False):
    """deep=False: every gradient within 5e-3 of the TF32-operand oracle.  deep=True (AlexNet, GoogLeNet at batch 2): the
    forward pass is held to 5e-3 at EVERY unit against both oracles; gradients that crossed many ReLU / arg-max layers at
    batch 2 amplify any forward difference (the tensor core's fp32 accumulation alone puts relu_conv5 1e-4 from the
    TF32-operand oracle, tools/step_diag.py), so there the bound is relative: tensor by tensor the GPU is within 3x, and
    summed over the net within 1.25x, of what TF32 operand rounding alone does to the reference fp32 algorithm on the CPU."""
    for lf, lt, lg in zip(fp32.get_loss_units(), tf32.get_loss_units(), gpu.get_loss_units()):
        a, t, b = lf.getloss(), lt.getloss(), lg.getloss()
        assert abs(a - b) <= TOL * abs(a) and abs(t - b) <= TOL * abs(t), (lf.name, a, t, b)
    worst_fwd = 0.0
    for uf, ut, ug in zip(fp32.units, tf32.units, gpu.units):
        if uf.out is None or getattr(ug, "fuse_relu", False):     # a fused conv's `out` is already rectified: checked at its ReLU unit
            continue
        a, t, b = (u.out.to_numpy().astype(np.float64) for u in (uf, ut, ug))
        for ref in (a, t):
            err = np.linalg.norm(ref - b) / max(np.linalg.norm(ref), 1e-30)
            worst_fwd = max(worst_fwd, err)
            assert err < TOL, (uf.name, "forward output", err)
    e_gt, e_gf, e_tf = _errs(tf32, gpu), _errs(fp32, gpu), _errs(fp32, tf32)
    worst = max(e_gt, key=e_gt.get)
    print("%s: forward worst %.2e; GPU vs TF32-operand oracle worst %s %.2e; GPU vs fp32 oracle worst %.2e (TF32-operand oracle vs "
          "fp32 oracle %.2e)" % (name, worst_fwd, worst, e_gt[worst], max(e_gf.values()), max(e_tf.values())))
    for k, v in e_gt.items():
        assert v < (max(TOL, 3.0 * e_tf[k]) if deep else TOL), (k, v, e_tf[k])
    for k, v in e_gf.items():
        assert v < max(TOL, 3.0 * e_tf[k] if deep else 1.5 * e_tf[k]), (k, v, e_tf[k])
    if deep:
        # a single tensor's distance is one draw of the ReLU / arg-max switching noise (the factor 3 above); summed over the
        # net's ~130 tensors the GPU must be no further from either oracle than the two oracles are from each other
        tot_tf, tot_gt, tot_gf = sum(e_tf.values()), sum(e_gt.values()), sum(e_gf.values())
        print("%s: sum over tensors: TF32-operand oracle vs fp32 oracle %.3f, GPU vs TF32-operand oracle %.3f, GPU vs fp32 oracle %.3f"
              % (name, tot_tf, tot_gt, tot_gf))
        assert tot_gt <= 1.25 * tot_tf and tot_gf <= 1.25 * tot_tf, (tot_tf, tot_gt, tot_gf)


def test_lenet_step_matches_cpu_oracle(gpu_owl_f):
    """configs[1] (apps/mnist_cnn, apps/mnist_common.h:123-222) at batch 16: loss and every gradient vs the oracle."""
    _check_step("lenet", *_step_nets(gpu_owl_f, "build_lenet", [28, 28, 1], 10, 16, 11, uniform=True))


def test_mlp_step_matches_cpu_oracle(gpu_owl_f):
    """configs[0] (apps/mnist_mlp, apps/mnist_common.h:224-288) at batch 16."""
    _check_step("mlp", *_step_nets(gpu_owl_f, "build_mnist_mlp", [784], 10, 16, 12, uniform=True))


def test_googlenet_step_matches_cpu_oracle(gpu_owl_f):
    """configs[4] at batch 2: three loss heads, nine inception modules, 57 convolutions, both LRNs, concat / slice,
    dropout masks from the same Philox stream."""
    _check_step("googlenet", *_step_nets(gpu_owl_f, "build_googlenet", [224, 224, 3], 1000, 2, 13), deep=True)


def test_alexnet_step_matches_cpu_oracle(gpu_owl_f):
    """configs[3] -- the headline net itself -- at batch 2 and full 227 x 227 geometry (the bench's batch 256 is covered
    layer by layer above)."""
    _check_step("alexnet", *_step_nets(gpu_owl_f, "build_alexnet", [227, 227, 3], 1000, 2, 14), deep=True)
