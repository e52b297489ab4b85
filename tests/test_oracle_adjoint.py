"""Derive the "parity unpinned" backward restatements from forward ops that ARE pinned by reference goldens.

The reference has neither CPU code nor tests for conv backward-data / -filter / -bias and pooling backward
(minerva/op/impl/bundle.h:35-44).  Their forward ops are pinned by the reference's own golden vectors
(tests/unittest_conv_forward.cpp:7-68, tests/unittest_pooling_forward.cpp:7-81 -> tests/test_oracle_golden.py).  A
backward op is by definition the adjoint of the forward op's linear part, so on the ORACLE (no GPU, no torch):

    <conv_fwd(x; w, 0), dy> == <x, conv_bwd_data(dy; w)> == <w, conv_bwd_filter(x, dy)>,   <1_c, dy> == conv_bwd_bias(dy)_c
    <avg_pool_fwd(x), dy> == <x, avg_pool_bwd(dy)>
    max pooling is piecewise linear: for a direction d and a step that switches no arg-max,
        <max_pool_fwd(x + t d) - max_pool_fwd(x), dy> == t <d, max_pool_bwd(dy)>      (exactly linear in t)

ties every backward restatement to a golden-pinned row.  Sums are taken in float64; the residual is the fp32 rounding
of the oracle's own accumulations."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.test_gpu_b_gemm_conv import CONV_CASES

rng = np.random.default_rng(31)


def _dot(a, b):
    return float(np.dot(np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()))


def _scale(a, b):
    return float(np.dot(np.abs(np.asarray(a, np.float64).ravel()), np.abs(np.asarray(b, np.float64).ravel())))


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_backward_is_the_adjoint_of_the_pinned_forward(case):
    N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw = case
    Ho, Wo = orc.conv_out(H, ph, fh, sv), orc.conv_out(W, pw, fw, sh)
    x = rng.normal(0, 1, N * Ci * H * W).astype(np.float32)
    w = rng.normal(0, 1, Co * Ci * fh * fw).astype(np.float32)
    dy = rng.normal(0, 1, N * Co * Ho * Wo).astype(np.float32)
    zero_b = np.zeros(Co, np.float32)
    y = orc.conv_forward(x, w, zero_b, *case)
    lhs = _dot(y, dy)
    dx = orc.conv_backward_data(dy, w, *case)
    dw = orc.conv_backward_filter(x, dy, *case)
    tol = 1e-6 * _scale(y, dy)           # 1e-6 of sum |y||dy|: the fp32 rounding of the oracle's own sums
    assert abs(lhs - _dot(x, dx)) <= tol, ("backward data", lhs, _dot(x, dx), tol)
    assert abs(lhs - _dot(w, dw)) <= tol, ("backward filter", lhs, _dot(w, dw), tol)
    # bias: y(b) - y(0) = b broadcast, so <b broadcast, dy> = <b, db>
    b = rng.normal(0, 1, Co).astype(np.float32)
    db = orc.conv_backward_bias(dy, N, Co, Ho, Wo)
    yb = orc.conv_forward(np.zeros_like(x), w, b, *case)          # = b broadcast over (n, i, j), exactly
    np.testing.assert_array_equal(yb.reshape(N, Co, Ho * Wo), np.broadcast_to(b[None, :, None], (N, Co, Ho * Wo)))
    assert abs(_dot(yb, dy) - _dot(b, db)) <= 1e-6 * _scale(yb, dy)


POOL_CASES = [
    # N, C, H, W, sv, sh, wh, ww, ph, pw
    (2, 3, 4, 4, 1, 1, 3, 3, 0, 0),     # the four golden geometries of unittest_pooling_forward.cpp
    (2, 3, 4, 4, 1, 1, 3, 3, 1, 1),
    (2, 3, 4, 4, 2, 2, 3, 3, 1, 1),
    (1, 2, 4, 4, 3, 3, 4, 4, 2, 2),
    (2, 4, 13, 13, 2, 2, 3, 3, 0, 0),   # AlexNet pool5
    (2, 4, 27, 27, 2, 2, 3, 3, 0, 0),   # AlexNet pool2
    (2, 2, 12, 12, 3, 3, 3, 3, 0, 0),   # LeNet pool2
    (2, 2, 24, 24, 2, 2, 2, 2, 0, 0),   # LeNet pool1
    (1, 3, 14, 14, 2, 2, 3, 3, 0, 0),   # GoogLeNet overhang
    (1, 3, 7, 9, 1, 1, 3, 3, 1, 1),     # GoogLeNet inception pool 3x3/1 pad 1
    (2, 3, 14, 14, 3, 3, 5, 5, 0, 0),   # GoogLeNet aux head avg 5x5/3
    (2, 3, 7, 7, 1, 1, 7, 7, 0, 0),     # GoogLeNet pool5 avg 7x7/1
]


@pytest.mark.parametrize("case", POOL_CASES)
def test_pooling_backward_is_the_adjoint_of_the_pinned_forward(case):
    N, C, H, W, sv, sh, wh, ww, ph, pw = case
    geo = (N, C, H, W, sv, sh, wh, ww, ph, pw)
    Ho, Wo = orc.pooled_size(H, ph, wh, sv), orc.pooled_size(W, pw, ww, sh)
    # distinct values => unique arg-max in every window, with a gap of at least 1/8 between any two
    x = (rng.permutation(N * C * H * W).astype(np.float32) / 8.0).astype(np.float32)
    dy = rng.normal(0, 1, N * C * Ho * Wo).astype(np.float32)
    # average pooling is linear
    y = orc.average_pooling_forward(x, *geo)
    dx = orc.average_pooling_backward(x, y, dy, *geo)
    assert abs(_dot(y, dy) - _dot(x, dx)) <= 1e-6 * max(_scale(y, dy), 1.0)
    # max pooling: a step of |t d| <= 1/32 < gap/2 cannot switch an arg-max, so the map is linear along d
    y = orc.max_pooling_forward(x, *geo)
    dx = orc.max_pooling_backward(x, y, dy, *geo)
    d = rng.integers(-4, 5, x.size).astype(np.float32) / 128.0          # multiples of 1/128: x + d is exact in fp32
    y2 = orc.max_pooling_forward((x + d).astype(np.float32), *geo)
    lhs, rhs = _dot(y2.astype(np.float64) - y.astype(np.float64), dy), _dot(d, dx)
    assert abs(lhs - rhs) <= 1e-6 * max(_scale(d, dx), 1e-3), (lhs, rhs)
    # every unit of top_diff is routed to exactly one bottom element
    assert abs(float(dx.astype(np.float64).sum()) - float(dy.astype(np.float64).sum())) <= 1e-5 * np.abs(dy).sum()


@pytest.mark.parametrize("N,C,H,W,size", [(2, 7, 3, 4, 5), (1, 16, 5, 5, 5), (2, 5, 2, 3, 3)])
def test_lrn_backward_is_the_gradient_of_the_forward(N, C, H, W, size):
    """LRN has no pinned row at all (cuda_kernel.h:223-331 restated); at least backward must be forward's gradient:
    central differences of <lrn_fwd(x), dy> along random directions, in float64 steps."""
    alpha, beta = 1e-1, 0.75                      # a large alpha so the cross-channel term matters
    x = rng.normal(0, 1, N * C * H * W).astype(np.float32)
    dy = rng.normal(0, 1, x.size).astype(np.float32)
    y, scale = orc.lrn_forward(x, size, alpha, beta, N, C, W, H)
    dx = orc.lrn_backward(x, y, scale, dy, size, alpha, beta, N, C, W, H)
    for _ in range(3):
        d = rng.normal(0, 1, x.size).astype(np.float32)
        t = 1e-2
        yp, _ = orc.lrn_forward((x + t * d).astype(np.float32), size, alpha, beta, N, C, W, H)
        ym, _ = orc.lrn_forward((x - t * d).astype(np.float32), size, alpha, beta, N, C, W, H)
        num = (_dot(yp, dy) - _dot(ym, dy)) / (2 * t)
        assert abs(num - _dot(d, dx)) <= 2e-3 * max(abs(num), 1.0), (num, _dot(d, dx))
