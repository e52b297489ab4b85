"""The oracle against every golden vector the reference's tests hold for the hot path
(SURVEY.md 8c): conv forward x2, max-pool forward x4, reduction closed forms, Flatten."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as orc


def _load(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def test_conv_forward_goldens(golden_dir):
    cases = _load(golden_dir, "conv_forward.json")
    assert len(cases) == 2
    for c in cases:
        W, H, Ci, N = c["input_size"]
        fw, fh, Ci2, Co = c["weight_size"]
        assert Ci == Ci2
        y = orc.conv_forward(c["input"], c["weight"], c["bias"], N, Ci, Co, H, W, c["pad_height"],
                             c["pad_width"], c["stride_vertical"], c["stride_horizontal"], fh, fw)
        want = np.array(c["correct"], np.float32)
        Wo, Ho, Co2, N2 = c["correct_size"]
        assert y.size == want.size == Wo * Ho * Co2 * N2
        if c["correct_excludes_bias"]:  # unittest_conv_forward.cpp:66
            want = want + np.array(c["bias"], np.float32)[(np.arange(want.size) // (Wo * Ho)) % Co]
        assert np.max(np.abs(y - want)) < c["tolerance"], c["name"]
        # the goldens carry ~1e-6 of noise; the restatement is far inside the reference's 1e-3
        assert np.max(np.abs(y - want)) < 2e-5, c["name"]


def test_conv_is_true_convolution_not_correlation(golden_dir):
    """an un-flipped kernel misses the golden by O(10)."""
    c = _load(golden_dir, "conv_forward.json")[0]
    W, H, Ci, N = c["input_size"]
    fw, fh, _, Co = c["weight_size"]
    w = np.array(c["weight"], np.float32).reshape(Co, Ci, fh, fw)[:, :, ::-1, ::-1].copy()
    y = orc.conv_forward(c["input"], w.ravel(), c["bias"], N, Ci, Co, H, W, 0, 0, 1, 1, fh, fw)
    assert np.max(np.abs(y - np.array(c["correct"], np.float32))) > 5.0


def test_max_pooling_goldens(golden_dir):
    cases = _load(golden_dir, "pooling_forward.json")
    assert len(cases) == 4
    for c in cases:
        W, H, Cc, N = c["input_size"]
        Wo, Ho, _, _ = c["correct_size"]
        assert orc.pooled_size(H, c["pad_height"], c["height"], c["stride_vertical"]) == Ho, c["name"]
        assert orc.pooled_size(W, c["pad_width"], c["width"], c["stride_horizontal"]) == Wo, c["name"]
        y = orc.max_pooling_forward(c["input"], N, Cc, H, W, c["stride_vertical"], c["stride_horizontal"],
                                    c["height"], c["width"], c["pad_height"], c["pad_width"])
        np.testing.assert_array_equal(y, np.array(c["correct"], np.float32), err_msg=c["name"])


def test_reduction_goldens(golden_dir):
    g = _load(golden_dir, "reduction.json")
    m, n = g["size"]
    x = np.array(g["input"], np.float32)
    np.testing.assert_array_equal(orc.reduction_on_col("max", x, m, n), np.array(g["max_dim0"], np.float32))
    np.testing.assert_array_equal(orc.reduction_on_row("max", x, m, n), np.array(g["max_dim1"], np.float32))
    np.testing.assert_array_equal(orc.reduction_on_col("sum", x, m, n), np.array(g["sum_dim0"], np.float32))
    np.testing.assert_array_equal(orc.reduction_on_row("sum", x, m, n), np.array(g["sum_dim1"], np.float32))


def test_flatten_goldens(golden_dir):
    """Layout rule every kernel relies on: index (i0,i1,..) -> i0 + d0*(i1 + d1*(...))."""
    g = _load(golden_dir, "scale_flatten.json")
    for c in g["cases"]:
        flat, mul = 0, 1
        for d, i in zip(c["dims"], c["idx"]):
            flat += i * mul
            mul *= d
        assert flat == c["flat"]
        if orc.have_ref():
            assert orc.Ref.flatten(c["dims"], c["idx"]) == c["flat"]


def test_pooled_size_matches_c():
    import ctypes
    fn = orc.lib().orc_pooled_size
    fn.restype = ctypes.c_int
    for x in range(1, 40):
        for k in range(1, 8):
            for s in range(1, 5):
                for p in range(0, k):
                    if x + 2 * p < k:
                        continue
                    assert fn(x, p, k, s) == orc.pooled_size(x, p, k, s)
