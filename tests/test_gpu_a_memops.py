"""-m gpu parity: the HBM-bound ops through the C ABI vs the CPU oracle on identical inputs.
Tolerances are north_star's: bit-exact for elementwise / activation / max-pooling / index outputs,
<= 1e-5 relative for fp32 reductions and softmax."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu

rng = np.random.default_rng(2024)
SIZES = [1, 3, 720, 4099, (1 << 20) + 3]   # 720 = Scale{2,3,4,5,6} of the reference tests


@pytest.fixture(scope="module")
def g():
    from tests import gpu_util
    return gpu_util


def _edge():
    return np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, 1.17e-38, 2.5,
                     -7.25, 1e-20, 1e20, 0.1], np.float32)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("name,ofn", [("mnv_add", orc.add), ("mnv_sub", orc.sub), ("mnv_dot_mult", orc.dot_mult),
                                      ("mnv_dot_div", orc.dot_div)])
def test_arithmetic_bit_exact(g, name, ofn, n):
    a = rng.normal(0, 1, n).astype(np.float32)
    b = rng.normal(0, 5, n).astype(np.float32)
    da, db, dc = g.dev(a), g.dev(b), g.empty(n)
    g.run(name, da, db, dc, n)
    g.assert_bits_equal(g.host(dc), ofn(a, b), name)


def test_arithmetic_edges_and_misaligned(g):
    e = _edge()
    a, b = np.repeat(e, e.size), np.tile(e, e.size)
    for name, ofn in (("mnv_add", orc.add), ("mnv_sub", orc.sub), ("mnv_dot_mult", orc.dot_mult), ("mnv_dot_div", orc.dot_div)):
        dc = g.empty(a.size)
        g.run(name, g.dev(a), g.dev(b), dc, a.size)
        g.assert_bits_equal(g.host(dc), ofn(a, b), name + " edge")
    # views that start 4 bytes into an allocation take the scalar path
    n = 5001
    a, b = rng.normal(0, 1, n + 1).astype(np.float32), rng.normal(0, 1, n + 1).astype(np.float32)
    da, db, dc = g.dev(a), g.dev(b), g.empty(n + 1)
    g.run("mnv_add", da[1:], db[1:], dc[1:], n)
    g.assert_bits_equal(g.host(dc)[1:], orc.add(a[1:], b[1:]), "misaligned add")
    # inputs may alias each other (relu_back(y, x, x) in the reference)
    g.run("mnv_dot_mult", da, da, dc, n + 1)
    g.assert_bits_equal(g.host(dc), orc.dot_mult(a, a), "aliased inputs")


@pytest.mark.parametrize("n", [1, 720, 4099, (1 << 20) + 3])
def test_in_place_forms(g, n):
    """mnv_accumulate / mnv_relu_mask_inplace: the explicit in-place entries (the reference-shaped ones never alias an
    output with an input) give the bits of the out-of-place ops."""
    acc = rng.normal(0, 1, n + 1).astype(np.float32)
    x = rng.normal(0, 1, n + 1).astype(np.float32)
    for off in (0, 1):           # off = 1: a view 4 bytes into the allocation takes the scalar path
        da, dx = g.dev(acc), g.dev(x)
        g.run("mnv_accumulate", da[off:], dx[off:], n)
        g.assert_bits_equal(g.host(da)[off:off + n], orc.add(acc[off:off + n], x[off:off + n]), "accumulate")
        da = g.dev(acc)
        g.run("mnv_relu_mask_inplace", da[off:], dx[off:], n)
        g.assert_bits_equal(g.host(da)[off:off + n], orc.relu_backward(x[off:off + n], x[off:off + n], acc[off:off + n]), "relu mask")


@pytest.mark.parametrize("n", [720, 4099, (1 << 18) + 1])
def test_arithmetic_const_bit_exact(g, n):
    x = rng.normal(0, 5, n).astype(np.float32)
    dx, dy = g.dev(x), g.empty(n)
    for v in (0.37, -3.0, 1e-3, 7.0):
        g.run("mnv_const_add", dx, dy, v, n); g.assert_bits_equal(g.host(dy), orc.const_add(x, v), "const_add")
        g.run("mnv_const_add", dx, dy, -v, n); g.assert_bits_equal(g.host(dy), orc.const_sub(x, v), "x - v as x + (-v)")
        g.run("mnv_left_const_sub", dx, dy, v, n); g.assert_bits_equal(g.host(dy), orc.left_const_sub(x, v), "left sub")
        g.run("mnv_left_const_div", dx, dy, v, n); g.assert_bits_equal(g.host(dy), orc.left_const_div(x, v), "left div")
        g.run("mnv_scale", dx, dy, n, v); g.assert_bits_equal(g.host(dy), orc.scale(x, v), "scale")
        g.run("mnv_const_div", dx, dy, v, n); g.assert_bits_equal(g.host(dy), orc.const_div(x, v), "true division")


def test_elewise(g):
    n = 100003
    x = rng.normal(0, 1, n).astype(np.float32)
    dx, dy = g.dev(x), g.empty(n)
    g.run("mnv_elewise_negative", dx, dy, n)
    g.assert_bits_equal(g.host(dy), orc.elewise_negative(x), "neg")
    g.run("mnv_elewise_exp", dx, dy, n)
    assert g.ulp_diff(g.host(dy), orc.elewise_exp(x)).max() <= 4   # EXPECT_FLOAT_EQ, unittest_elewise.cpp:17
    xl = rng.normal(500, 1, n).astype(np.float32)                   # unittest_elewise.cpp:43
    g.run("mnv_elewise_ln", g.dev(xl), dy, n)
    assert g.ulp_diff(g.host(dy), orc.elewise_ln(xl)).max() <= 4


def test_exact_mode_transcendentals_bit_exact(g):
    """mnv_*_exact (exp, ln, sigmoid, tanh) return the CPU reference's bits: compared with the reference's own basic::
    functions (oracle/_ref, which call this box's libm) where built, and with the oracle restatement, on (a) every 1021st
    of all 2^32 float bit patterns (4.2 M values: all exponents, subnormals, both signs, inf / nan), (b) the reference
    tests' input ranges (unittest_elewise.cpp:13,43, unittest_activation.cpp:15,45), (c) the edge vector."""
    sweep = np.arange(0, 1 << 32, 1021, dtype=np.uint64).astype(np.uint32).view(np.float32)
    x = np.concatenate([sweep, rng.normal(0, 1, 200003).astype(np.float32), rng.normal(500, 1, 100003).astype(np.float32),
                        rng.normal(0, 5, 100003).astype(np.float32), rng.uniform(-104, 89, 200003).astype(np.float32), _edge()])
    n = x.size
    dx, dy = g.dev(x), g.empty(n)
    cases = [("mnv_elewise_exp_exact", (n,), orc.elewise_exp, lambda v: orc.Ref.elewise("exp", v)),
             ("mnv_elewise_ln_exact", (n,), orc.elewise_ln, lambda v: orc.Ref.elewise("ln", v)),
             ("mnv_sigmoid_forward_exact", (1, 1, 1, n), orc.sigmoid_forward, lambda v: orc.Ref.activation("sigmoid", v)),
             ("mnv_tanh_forward_exact", (1, 1, 1, n), orc.tanh_forward, lambda v: orc.Ref.activation("tanh", v))]
    with np.errstate(all="ignore"):
        for name, dims, oracle_fn, ref_fn in cases:
            dy.fill_(float("nan"))
            g.run(name, dx, dy, *dims)
            got = g.host(dy)
            g.assert_bits_equal(got, oracle_fn(x), name + " vs oracle")
            if orc.have_ref():
                g.assert_bits_equal(got, ref_fn(x), name + " vs the compiled reference")
    # the default (fast) entries stay within the reference tests' 4 ulp on the reference tests' ranges
    xs = rng.normal(0, 1, 100003).astype(np.float32)
    d2 = g.empty(xs.size)
    g.run("mnv_sigmoid_forward", g.dev(xs), d2, 1, 1, 1, xs.size)
    assert g.ulp_diff(g.host(d2), orc.sigmoid_forward(xs)).max() <= 4


def test_activation(g):
    n = 50001
    x = np.concatenate([rng.normal(0, 1, n).astype(np.float32), _edge()])
    n = x.size
    dy_ = rng.normal(0, 1, n).astype(np.float32)
    dx, dout, ddy = g.dev(x), g.empty(n), g.dev(dy_)
    g.run("mnv_relu_forward", dx, dout, 1, 1, 1, n)
    g.assert_bits_equal(g.host(dout), orc.relu_forward(x), "relu fwd")          # incl. -0.0 -> +0.0, NaN -> 0
    y = orc.relu_forward(x)
    dres = g.empty(n)
    g.run("mnv_relu_backward", dx, g.dev(y), ddy, dres, 1, 1, 1, n)
    g.assert_bits_equal(g.host(dres), orc.relu_backward(x, y, dy_), "relu bwd")
    fin = np.isfinite(x)
    for kind in ("sigmoid", "tanh"):
        g.run("mnv_%s_forward" % kind, dx, dout, 1, 1, 1, n)
        want = getattr(orc, kind + "_forward")(x)
        got = g.host(dout)
        assert g.ulp_diff(got[fin], want[fin]).max() <= 4, kind                 # unittest_activation.cpp:19,49
        yk = want
        g.run("mnv_%s_backward" % kind, dx, g.dev(yk), ddy, dres, 1, 1, 1, n)
        g.assert_bits_equal(g.host(dres)[fin], getattr(orc, kind + "_backward")(x, yk, dy_)[fin], kind + " bwd")


@pytest.mark.parametrize("m,n", [(9, 7), (10, 256), (4096, 256), (1000, 33), (1, 5), (5, 1)])
def test_norm_arithmetic_bit_exact(g, m, n):
    mat = rng.normal(0, 1, m * n).astype(np.float32)
    vc, vr = rng.normal(0, 5, n).astype(np.float32), rng.normal(0, 5, m).astype(np.float32)
    dm, dres = g.dev(mat), g.empty(m * n)
    for op, sfx in (("add", "add"), ("sub", "sub"), ("mult", "mult"), ("div", "div")):
        g.run("mnv_norm_%s_on_col" % sfx, dm, g.dev(vc), dres, m, n)
        g.assert_bits_equal(g.host(dres), orc.norm_on_col(op, mat, vc, m, n), "on_col " + op)
        g.run("mnv_norm_%s_on_row" % sfx, dm, g.dev(vr), dres, m, n)
        g.assert_bits_equal(g.host(dres), orc.norm_on_row(op, mat, vr, m, n), "on_row " + op)


@pytest.mark.parametrize("m,n", [(5, 3), (9, 7), (10, 256), (1000, 256), (4096, 256), (1, 256), (300, 1), (33, 1000)])
def test_reduction_and_max_index(g, m, n):
    x = rng.normal(0, 1, m * n).astype(np.float32)
    if m * n > 8:
        x[5] = x[2] = x.max() + 1   # a tie for the maximum: first index must win
    dx = g.dev(x)
    oc, orow = g.empty(n), g.empty(m)
    g.run("mnv_reduction_max_on_col", dx, oc, m, n); g.assert_bits_equal(g.host(oc), orc.reduction_on_col("max", x, m, n), "max col")
    g.run("mnv_reduction_max_on_row", dx, orow, m, n); g.assert_bits_equal(g.host(orow), orc.reduction_on_row("max", x, m, n), "max row")
    g.run("mnv_max_index_on_col", dx, oc, m, n); g.assert_bits_equal(g.host(oc), orc.max_index_on_col(x, m, n), "argmax col")
    g.run("mnv_max_index_on_row", dx, orow, m, n); g.assert_bits_equal(g.host(orow), orc.max_index_on_row(x, m, n), "argmax row")
    X = np.abs(x.astype(np.float64)).reshape(n, m)
    g.run("mnv_reduction_sum_on_col", dx, oc, m, n)
    assert np.all(np.abs(g.host(oc).astype(np.float64) - orc.reduction_on_col("sum", x, m, n)) <= 1e-5 * X.sum(1) + 1e-30)
    g.run("mnv_reduction_sum_on_row", dx, orow, m, n)
    assert np.all(np.abs(g.host(orow).astype(np.float64) - orc.reduction_on_row("sum", x, m, n)) <= 1e-5 * X.sum(0) + 1e-30)


def test_reduction_golden(g, golden_dir):
    gd = json.load(open(os.path.join(golden_dir, "reduction.json")))
    m, n = gd["size"]
    dx = g.dev(gd["input"])
    oc, orow = g.empty(n), g.empty(m)
    g.run("mnv_reduction_max_on_col", dx, oc, m, n); assert g.host(oc).tolist() == gd["max_dim0"]
    g.run("mnv_reduction_max_on_row", dx, orow, m, n); assert g.host(orow).tolist() == gd["max_dim1"]
    g.run("mnv_reduction_sum_on_col", dx, oc, m, n); assert g.host(oc).tolist() == gd["sum_dim0"]
    g.run("mnv_reduction_sum_on_row", dx, orow, m, n); assert g.host(orow).tolist() == gd["sum_dim1"]


@pytest.mark.parametrize("m,n", [(9, 7), (33, 65), (4096, 1000), (1, 17), (256, 9216)])
def test_transpose_copy(g, m, n):
    a = rng.normal(0, 1, m * n).astype(np.float32)
    da, dc = g.dev(a), g.empty(m * n)
    g.run("mnv_transpose", da, dc, m, n)
    g.assert_bits_equal(g.host(dc), orc.transpose(a, m, n), "transpose")
    g.run("mnv_copy", da, dc, m * n); g.assert_bits_equal(g.host(dc), a, "copy")
    dc.zero_()
    g.run("mnv_reshape", da, dc, m * n * 4); g.assert_bits_equal(g.host(dc), a, "reshape")


def test_concat_slice_select(g):
    N, H, W, C1, C2 = 5, 3, 7, 2, 4   # inner sizes not multiples of 4 -> scalar path
    for (h, w) in ((H, W), (4, 8)):   # and a vectorisable one
        a = rng.normal(0, 1, (N, C1, h, w)).astype(np.float32)
        b = rng.normal(0, 1, (N, C2, h, w)).astype(np.float32)
        out = g.empty(N * (C1 + C2) * h * w)
        g.run("mnv_copy_strided", g.dev(a), out, C1 * h * w, N, C1 * h * w, (C1 + C2) * h * w)
        g.run("mnv_copy_strided", g.dev(b), out[C1 * h * w:], C2 * h * w, N, C2 * h * w, (C1 + C2) * h * w)
        g.assert_bits_equal(g.host(out).reshape(N, C1 + C2, h, w), np.concatenate([a, b], 1), "concat")
        sl = g.empty(N * 2 * h * w)
        g.run("mnv_copy_strided", out[h * w:], sl, 2 * h * w, N, (C1 + C2) * h * w, 2 * h * w)
        g.assert_bits_equal(g.host(sl).reshape(N, 2, h, w), np.concatenate([a, b], 1)[:, 1:3], "slice")
    import torch
    rows, cols = 13, 9
    src = rng.normal(0, 1, rows * cols).astype(np.float32)
    idx = np.array([8, 0, 3, 3], np.int32)
    dst = g.empty(rows * idx.size)
    g.run("mnv_select", dst, g.dev(src), torch.from_numpy(idx).cuda(), idx.size, cols, rows)
    g.assert_bits_equal(g.host(dst).reshape(idx.size, rows), src.reshape(cols, rows)[idx], "select")


@pytest.mark.parametrize("N,C,H,W", [(8, 1, 1, 10), (256, 1, 1, 1000), (120, 1, 1, 1000), (3, 5, 2, 4), (2, 7, 13, 13)])
def test_softmax(g, N, C, H, W):
    x = rng.normal(0, 3, N * C * H * W).astype(np.float32)
    dy_ = rng.normal(0, 1, x.size).astype(np.float32)
    dx, dout = g.dev(x), g.empty(x.size)
    for mode in ("instance", "channel"):
        g.run("mnv_%s_softmax_forward" % mode, dx, dout, N, C, H, W)
        want = getattr(orc, mode + "_softmax_forward")(x, N, C, H, W)
        np.testing.assert_allclose(g.host(dout), want, rtol=1e-5, atol=1e-30, err_msg=mode)
        dres = g.empty(x.size)
        g.run("mnv_%s_softmax_backward" % mode, g.dev(dy_), g.dev(want), dres, N, C, H, W)
        wb = getattr(orc, mode + "_softmax_backward")(dy_, want, N, C, H, W)
        assert np.abs(g.host(dres) - wb).max() <= 1e-5 * np.abs(wb).max() + 1e-12, mode + " bwd"


POOL_CASES = [
    (2, 3, 4, 4, 1, 1, 3, 3, 0, 0), (2, 3, 4, 4, 2, 2, 3, 3, 1, 1), (1, 2, 4, 4, 3, 3, 4, 4, 2, 2),
    (4, 8, 55, 55, 2, 2, 3, 3, 0, 0), (4, 8, 27, 27, 2, 2, 3, 3, 0, 0), (4, 8, 13, 13, 2, 2, 3, 3, 0, 0),
    (4, 16, 24, 24, 2, 2, 2, 2, 0, 0), (4, 32, 12, 12, 3, 3, 3, 3, 0, 0), (2, 3, 14, 14, 2, 2, 3, 3, 0, 0),
    (2, 3, 7, 9, 1, 1, 3, 3, 1, 1), (2, 4, 14, 14, 3, 3, 5, 5, 0, 0), (2, 4, 7, 7, 1, 1, 7, 7, 0, 0),
    (5, 37, 28, 28, 1, 1, 3, 3, 1, 1), (3, 11, 14, 14, 1, 1, 3, 3, 1, 1), (2, 300, 7, 7, 1, 1, 3, 3, 1, 1), (2, 2, 1, 5, 1, 1, 3, 3, 1, 1),   # inception pools (column-strip kernels)
    (3, 5, 8, 11, 2, 2, 3, 3, 0, 0), (2, 3, 3, 3, 2, 2, 3, 3, 0, 0), (300, 7, 6, 6, 2, 2, 3, 3, 0, 0),   # 3x3/2: overhanging windows, one window, many planes
    (1, 2, 70, 131, 2, 2, 3, 3, 0, 0), (2, 3, 9, 64, 2, 2, 3, 3, 0, 0), (2, 2, 5, 4, 2, 2, 3, 3, 0, 0),    # wider than 64 columns (thread-per-block kernel), exactly 64, tiny
]


def test_max_pooling_goldens(g, golden_dir):
    for c in json.load(open(os.path.join(golden_dir, "pooling_forward.json"))):
        W, H, C, N = c["input_size"]
        out = g.empty(np.prod(c["correct_size"]))
        g.run("mnv_max_pooling_forward", g.dev(c["input"]), out, N, C, H, W, c["stride_vertical"], c["stride_horizontal"],
              c["height"], c["width"], c["pad_height"], c["pad_width"])
        assert g.host(out).tolist() == c["correct"], c["name"]


@pytest.mark.parametrize("case", POOL_CASES)
@pytest.mark.parametrize("relu_input", [False, True])
def test_pooling_bit_exact(g, case, relu_input):
    N, C, H, W, sv, sh, wh, ww, ph, pw = case
    x = rng.normal(0, 1, N * C * H * W).astype(np.float32)
    if relu_input:
        x = np.maximum(x, 0)   # ~50% exact zeros: windows full of ties
    Ho, Wo = orc.pooled_size(H, ph, wh, sv), orc.pooled_size(W, pw, ww, sh)
    dy_ = rng.normal(0, 1, N * C * Ho * Wo).astype(np.float32)
    dx, dyd = g.dev(x), g.dev(dy_)
    for kind, fname in (("max", "max"), ("average", "average")):
        out = g.empty(N * C * Ho * Wo)
        g.run("mnv_%s_pooling_forward" % fname, dx, out, N, C, H, W, sv, sh, wh, ww, ph, pw)
        want = getattr(orc, kind + "_pooling_forward")(x, N, C, H, W, sv, sh, wh, ww, ph, pw)
        g.assert_bits_equal(g.host(out), want, kind + " fwd")
        dres = g.empty(x.size)
        g.run("mnv_%s_pooling_backward" % fname, dx, g.dev(want), dyd, dres, N, C, H, W, sv, sh, wh, ww, ph, pw)
        wb = getattr(orc, kind + "_pooling_backward")(x, want, dy_, N, C, H, W, sv, sh, wh, ww, ph, pw)
        g.assert_bits_equal(g.host(dres), wb, kind + " bwd")
        if kind == "max" and relu_input:   # fused ReLU backward: mask by bottom > 0 in the same kernel (or a second pass)
            dfus = g.empty(x.size)
            g.run("mnv_max_pooling_backward_relu", dx, g.dev(want), dyd, dfus, N, C, H, W, sv, sh, wh, ww, ph, pw)
            g.assert_bits_equal(g.host(dfus), np.where(x > 0, wb, np.float32(0)).astype(np.float32), "max bwd + relu bwd")
        if kind == "max" and (sv, sh, wh, ww, ph, pw) in ((2, 2, 3, 3, 0, 0), (1, 1, 3, 3, 1, 1)):
            # arg-max remembering pair: same top, same bottom_diff, with and without the folded ReLU mask
            import torch
            out2 = g.empty(N * C * Ho * Wo)
            idx = torch.full((N * C * Ho * Wo,), 77, dtype=torch.uint8, device="cuda")
            g.run("mnv_max_pooling_forward_idx", dx, out2, idx, N, C, H, W, sv, sh, wh, ww, ph, pw)
            g.assert_bits_equal(g.host(out2), want, "max fwd (idx variant)")
            assert int(idx.max()) <= 8
            didx = g.empty(x.size); didx.fill_(float("nan"))
            g.run("mnv_max_pooling_backward_idx", dyd, idx, 0, didx, N, C, H, W, sv, sh, wh, ww, ph, pw)
            g.assert_bits_equal(g.host(didx), wb, "max bwd from arg-max bytes")
            if relu_input:
                didx.fill_(float("nan"))
                g.run("mnv_max_pooling_backward_idx", dyd, idx, out2, didx, N, C, H, W, sv, sh, wh, ww, ph, pw)
                g.assert_bits_equal(g.host(didx), np.where(x > 0, wb, np.float32(0)).astype(np.float32), "max bwd from arg-max bytes + relu bwd")
        elif kind == "max":
            from minerva_b200._lib import MnvError
            import torch
            with pytest.raises(MnvError):
                g.run("mnv_max_pooling_forward_idx", dx, g.empty(N * C * Ho * Wo), torch.empty(N * C * Ho * Wo, dtype=torch.uint8, device="cuda"),
                      N, C, H, W, sv, sh, wh, ww, ph, pw)


@pytest.mark.parametrize("N,C,H,W,size", [(2, 7, 3, 4, 5), (4, 96, 27, 27, 5), (2, 256, 13, 13, 5), (2, 5, 2, 3, 3),
                                          (1, 6, 2, 2, 4), (2, 3, 4, 4, 5), (3, 96, 55, 55, 5), (2, 5, 16, 16, 5), (2, 19, 8, 7, 5)])
def test_lrn(g, N, C, H, W, size):
    alpha, beta = 1e-4, 0.75
    x = rng.normal(0, 20, N * C * H * W).astype(np.float32)
    dy_ = rng.normal(0, 1, x.size).astype(np.float32)
    dscale, dout = g.empty(x.size), g.empty(x.size)
    g.run("mnv_lrn_forward", g.dev(x), dscale, dout, size, alpha, beta, N, C, W, H)
    y, scale = orc.lrn_forward(x, size, alpha, beta, N, C, W, H)
    g.assert_bits_equal(g.host(dscale), scale, "lrn scale")
    np.testing.assert_allclose(g.host(dout), y, rtol=1e-5, atol=1e-30)
    dres = g.empty(x.size)
    g.run("mnv_lrn_backward", g.dev(x), g.dev(y), g.dev(scale), g.dev(dy_), dres, size, alpha, beta, N, C, W, H)
    wb = orc.lrn_backward(x, y, scale, dy_, size, alpha, beta, N, C, W, H)
    assert np.abs(g.host(dres) - wb).max() <= 1e-5 * np.abs(wb).max()
    # fused ReLU backward == the unfused pair, bit for bit (x here has both signs: the mask is bottom > 0)
    dfus = g.empty(x.size)
    g.run("mnv_lrn_backward_relu", g.dev(x), g.dev(y), g.dev(scale), g.dev(dy_), dfus, size, alpha, beta, N, C, W, H)
    g.assert_bits_equal(g.host(dfus), np.where(x > 0, g.host(dres), np.float32(0)).astype(np.float32), "lrn bwd + relu bwd")
    # scale-less pair: forward output and backward result bit-identical to the three-array form fed with the GPU's own arrays
    if size == 5 and C >= 5:
        dout2 = g.empty(x.size)
        g.run("mnv_lrn_forward_lite", g.dev(x), dout2, size, alpha, beta, N, C, W, H)
        g.assert_bits_equal(g.host(dout2), g.host(dout), "lrn forward without scale")
        dref = g.empty(x.size)
        g.run("mnv_lrn_backward", g.dev(x), dout, dscale, g.dev(dy_), dref, size, alpha, beta, N, C, W, H)
        for relu in (0, 1):
            dlite = g.empty(x.size); dlite.fill_(float("nan"))
            g.run("mnv_lrn_backward_lite", g.dev(x), g.dev(dy_), dlite, size, alpha, beta, N, C, W, H, relu)
            want = np.where(x > 0, g.host(dref), np.float32(0)).astype(np.float32) if relu else g.host(dref)
            g.assert_bits_equal(g.host(dlite), want, "lrn backward from (bottom, top_diff), relu=%d" % relu)
    else:
        from minerva_b200._lib import MnvError
        with pytest.raises(MnvError):
            g.run("mnv_lrn_forward_lite", g.dev(x), g.empty(x.size), size, alpha, beta, N, C, W, H)


@pytest.mark.parametrize("N,C,H,W", [(2, 5, 4, 4), (16, 96, 55, 55), (8, 256, 13, 13), (256, 10, 1, 1), (3, 1000, 1, 1)])
def test_conv_backward_bias(g, N, C, H, W):
    dy_ = rng.normal(0.1, 1, N * C * H * W).astype(np.float32)
    want = orc.conv_backward_bias(dy_, N, C, H, W)
    l1 = np.abs(dy_.astype(np.float64)).reshape(N, C, H * W).sum((0, 2))
    ws = g.workspace()
    for wsp, wsb in ((ws, ws.numel()), (0, 0)):
        db = g.empty(C)
        g.run("mnv_conv_backward_bias", g.dev(dy_), db, N, C, H, W, wsp, wsb)
        assert np.all(np.abs(g.host(db).astype(np.float64) - want) <= 1e-5 * l1)


def test_add_n_is_the_chained_add(g):
    """mnv_add_n: ((a0 + a1) + a2) + ... -- bit-identical to the reference's chain of Add ops, aligned and unaligned."""
    import ctypes
    for n in (1, 7, 4096, 100003):
        for count in (1, 2, 4, 8):
            srcs = [rng.normal(0, 1, n + 1).astype(np.float32) for _ in range(count)]
            want = srcs[0].copy()
            for v in srcs[1:]:
                want = (want + v).astype(np.float32)
            for off in (0, 1):
                devs = [g.dev(v) for v in srcs]
                out = g.empty(n + 1)
                ptrs = (ctypes.c_void_p * count)(*[d[off:].data_ptr() for d in devs])
                g.run("mnv_add_n", ptrs, count, out[off:], n)
                g.assert_bits_equal(g.host(out)[off:off + n], want[off:off + n], "add_n %d x %d" % (count, n))


def test_generators_and_fill(g):
    for n in (1, 5, 1000, 100003):
        out = g.empty(n + 1)
        g.run("mnv_fill", out[1:], n, 0.25)     # misaligned start
        g.assert_bits_equal(g.host(out)[1:], orc.fill(n, 0.25), "fill")
        g.run("mnv_rand_bernoulli", out, n, 42, 0.3)
        g.assert_bits_equal(g.host(out)[:n], orc.rand_bernoulli(n, 42, 0.3), "bernoulli")
        g.run("mnv_randn", out, n, 7, 1.5, 0.5)
        np.testing.assert_allclose(g.host(out)[:n], orc.randn(n, 7, 1.5, 0.5), rtol=0, atol=2e-5)
        # key read on the device (the form a recorded CUDA graph replays): (*base + add) ^ xor, 32-bit wrap-around
        import torch
        base = torch.tensor([0xFFFFFFF0 - (1 << 32)], dtype=torch.int32, device="cuda")
        g.run("mnv_rand_bernoulli_ds", out, n, base, 0x35, 0x5A5A, 0.3)
        seed = ((0xFFFFFFF0 + 0x35) & 0xFFFFFFFF) ^ 0x5A5A
        g.assert_bits_equal(g.host(out)[:n], orc.rand_bernoulli(n, seed, 0.3), "bernoulli, device seed")
    n = 1 << 20
    out = g.empty(n)
    g.run("mnv_randn", out, n, 99, 0.0, 0.01)   # AlexNet gaussian filler
    z = g.host(out)
    assert abs(z.mean()) < 1e-4 and abs(z.std() - 0.01) < 1e-4
    g.run("mnv_rand_bernoulli", out, n, 99, 0.5)
    assert abs(g.host(out).mean() - 0.5) < 2e-3


def test_sgd_update_bit_exact(g):
    for n in (7, 4096, 100003):
        w, d, gr = (rng.normal(0, 1, n).astype(np.float32) for _ in range(3))
        dw, dd = g.dev(w), g.dev(d)
        g.run("mnv_sgd_momentum_update", dw, dd, g.dev(gr), n, 0.9, 0.01 / 256, 0.01 * 5e-4)
        w2, d2 = orc.sgd_momentum_update(w, d, gr, 0.9, 0.01 / 256, 0.01 * 5e-4)
        g.assert_bits_equal(g.host(dd), d2, "delta")
        g.assert_bits_equal(g.host(dw), w2, "w")


def test_sgd_update_multi_equals_single_calls(g):
    """mnv_sgd_momentum_update_multi: one launch over many tensors (sizes from 1 element to several chunks, one of them
    misaligned, one empty) == one mnv_sgd_momentum_update per tensor, bit for bit."""
    import ctypes
    import torch
    from minerva_b200 import _lib
    from minerva_b200.owl.narray import NArray
    sizes = [1, 96, 0, 4096, 4097, 1000, 37748736 // 64, 12289, 5] * 9          # 81 tensors: more than one 64-entry table
    ws, ds, gs = [], [], []
    arr = (NArray._SgdTensor * len(sizes))()
    for i, n in enumerate(sizes):
        off = 1 if i == 4 else 0
        w, d, gr = (torch.randn(n + off, device="cuda")[off:] for _ in range(3))
        ws.append(w); ds.append(d); gs.append(gr)
        arr[i] = NArray._SgdTensor(w.data_ptr(), d.data_ptr(), gr.data_ptr(), n, 0.01 / 256 * (1 + i % 3), 5e-6 * (i % 2))
    w1, d1 = [w.clone() for w in ws], [d.clone() for d in ds]
    lib = _lib.load()
    for i, n in enumerate(sizes):
        if n:
            _lib.check(lib.mnv_sgd_momentum_update(w1[i].data_ptr(), d1[i].data_ptr(), gs[i].data_ptr(), n, 0.9, 0.01 / 256 * (1 + i % 3),
                                                   5e-6 * (i % 2), g.stream()), "single")
    _lib.check(lib.mnv_sgd_momentum_update_multi(ctypes.cast(arr, ctypes.c_void_p), len(sizes), 0.9, g.stream()), "multi")
    torch.cuda.synchronize()
    for i in range(len(sizes)):
        assert torch.equal(ws[i], w1[i]) and torch.equal(ds[i], d1[i]), i


def test_bad_arguments_fail_loudly(g):
    from minerva_b200 import _lib
    with pytest.raises(_lib.MnvError):
        g.run("mnv_add", 0, 0, 0, 16)
    with pytest.raises(_lib.MnvError):
        g.run("mnv_transpose", g.empty(4), g.empty(4), -1, 4)
