"""Pin the C restatement bit-for-bit against the REFERENCE's own CPU code (basic.cpp compiled
into oracle/_ref by oracle/Makefile) on the reference tests' shapes and distributions
(tests/unittest_arithmetic.cpp, unittest_arithmetic_const.cpp, unittest_elewise.cpp,
unittest_activation.cpp, unittest_norm_arithmetic.cpp, unittest_sgemm.cpp)."""
import numpy as np
import pytest

from oracle import pyoracle as orc

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (no /root/reference)")

rng = np.random.default_rng(1234)
N720 = 2 * 3 * 4 * 5 * 6  # Scale{2,3,4,5,6} of the reference tests


def _edge():
    return np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, 1.17e-38,
                     2.5, -7.25, 1e-20, 1e20, 0.1], np.float32)


@pytest.mark.parametrize("op,fn", [("add", orc.add), ("sub", orc.sub), ("mult", orc.dot_mult), ("div", orc.dot_div)])
def test_arithmetic(op, fn):
    a = rng.normal(0, 1, N720).astype(np.float32)
    b = rng.normal(0, 5, N720).astype(np.float32)
    np.testing.assert_array_equal(fn(a, b), orc.Ref.arithmetic(op, a, b))
    e = _edge()
    aa, bb = np.repeat(e, e.size), np.tile(e, e.size)
    np.testing.assert_array_equal(fn(aa, bb).view(np.uint32) & 0x7FFFFFFF | (np.isnan(fn(aa, bb)) * 0),
                                  orc.Ref.arithmetic(op, aa, bb).view(np.uint32) & 0x7FFFFFFF)


@pytest.mark.parametrize("op,side,fn", [
    ("add", 1, orc.const_add), ("add", 0, orc.const_add), ("sub", 1, orc.const_sub),
    ("sub", 0, orc.left_const_sub), ("mult", 1, orc.scale), ("mult", 0, orc.scale),
    ("div", 1, orc.const_div), ("div", 0, orc.left_const_div)])
def test_arithmetic_const(op, side, fn):
    x = rng.normal(0, 5, N720).astype(np.float32)
    for v in (0.37, -3.0, 1e-3, 7.0):
        np.testing.assert_array_equal(fn(x, v), orc.Ref.arithmetic_const(op, side, v, x))


def test_elewise():
    x = rng.normal(0, 1, N720).astype(np.float32)
    np.testing.assert_array_equal(orc.elewise_exp(x), orc.Ref.elewise("exp", x))
    np.testing.assert_array_equal(orc.elewise_negative(x), orc.Ref.elewise("negative", x))
    xl = rng.normal(500, 1, N720).astype(np.float32)  # unittest_elewise.cpp:43
    np.testing.assert_array_equal(orc.elewise_ln(xl), orc.Ref.elewise("ln", xl))


def test_activation_forward():
    x = np.concatenate([rng.normal(0, 1, N720).astype(np.float32), _edge()])
    for kind, fn in (("sigmoid", orc.sigmoid_forward), ("relu", orc.relu_forward), ("tanh", orc.tanh_forward)):
        a, b = fn(x), orc.Ref.activation(kind, x)
        np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32), err_msg=kind)


@pytest.mark.parametrize("m,n,k", [(3, 5, 2), (9, 7, 11), (64, 33, 100)])
def test_matmult(m, n, k):
    a = rng.normal(0, 1, m * k).astype(np.float32)
    b = rng.normal(0, 1, k * n).astype(np.float32)
    np.testing.assert_array_equal(orc.matmult(a, b, m, n, k), orc.Ref.matmult(a, b, m, n, k))


def test_transpose():
    a = rng.normal(0, 1, 9 * 7).astype(np.float32)
    np.testing.assert_array_equal(orc.transpose(a, 9, 7), orc.Ref.transpose(a, 9, 7))


@pytest.mark.parametrize("m,n", [(5, 3), (9, 7), (128, 33)])
def test_reduction_and_max_index(m, n):
    x = rng.normal(0, 1, m * n).astype(np.float32)
    x[3] = x[1]  # a tie
    for kind in ("sum", "max"):
        np.testing.assert_array_equal(orc.reduction_on_col(kind, x, m, n), orc.Ref.reduction(kind, 0, x, m, n))
        np.testing.assert_array_equal(orc.reduction_on_row(kind, x, m, n), orc.Ref.reduction(kind, 1, x, m, n))
    np.testing.assert_array_equal(orc.max_index_on_col(x, m, n), orc.Ref.max_index(0, x, m, n))
    np.testing.assert_array_equal(orc.max_index_on_row(x, m, n), orc.Ref.max_index(1, x, m, n))


@pytest.mark.parametrize("op", ["add", "sub", "mult", "div"])
def test_norm_arithmetic(op):
    m, n = 9, 7
    mat = rng.normal(0, 1, m * n).astype(np.float32)
    np.testing.assert_array_equal(orc.norm_on_col(op, mat, rng.normal(0, 5, n).astype(np.float32) * 0 + 2.5, m, n),
                                  orc.Ref.norm_arithmetic(op, 0, mat, np.full(n, 2.5, np.float32), m, n))
    vc, vr = rng.normal(0, 5, n).astype(np.float32), rng.normal(0, 5, m).astype(np.float32)
    np.testing.assert_array_equal(orc.norm_on_col(op, mat, vc, m, n), orc.Ref.norm_arithmetic(op, 0, mat, vc, m, n))
    np.testing.assert_array_equal(orc.norm_on_row(op, mat, vr, m, n), orc.Ref.norm_arithmetic(op, 1, mat, vr, m, n))


def test_softmax_instance():
    # {W=10,H=1,C=1,N=8}: the reference CPU path normalises over dim 0 (basic.cpp:231), which is
    # cuDNN's instance mode when H=C=1 -- the only way owl calls it (owl/conv.py:30-33).
    x = rng.normal(0, 3, 10 * 8).astype(np.float32)
    np.testing.assert_array_equal(orc.instance_softmax_forward(x, 8, 1, 1, 10), orc.Ref.softmax_forward(x, 10, 1, 1, 8))
    x = rng.normal(0, 3, 1000 * 16).astype(np.float32)
    np.testing.assert_array_equal(orc.instance_softmax_forward(x, 16, 1, 1, 1000), orc.Ref.softmax_forward(x, 1000, 1, 1, 16))


def test_fill():
    np.testing.assert_array_equal(orc.fill(37, 0.25), orc.Ref.fill(37, 0.25))


def test_exact_mode_math_is_glibc(tmp_path):
    """minerva_b200/csrc/glibc_math.h -- the text the exact-mode CUDA kernels compile -- built for the host and compared
    with this machine's libm on every 97th of the 2^32 float inputs (44 M values per function, ~1 s): 0 mismatches for expf,
    logf, tanhf and the reference's sigmoid formula.  The full sweep (stride 1: 0 mismatches, 30 s on 8 cores) is
    `tests/cpp/check_glibc_math.c 1`; MNV_CHECK_MATH_STRIDE overrides the stride."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "check_glibc_math")
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp", "-mfma", "-ffp-contract=off", "-I" + os.path.join(root, "minerva_b200", "csrc"),
                           os.path.join(root, "tests", "cpp", "check_glibc_math.c"), "-o", exe, "-lm"])
    out = subprocess.run([exe, os.environ.get("MNV_CHECK_MATH_STRIDE", "97")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    for fn in ("expf", "logf", "tanhf", "sigmoid"):
        assert fn + ": 0 mismatches" in out.stdout, out.stdout


def test_reference_stack_runs_config0():
    """BASELINE configs[0] through the REFERENCE'S OWN stack (its 27 CPU sources compiled into oracle/_ref/libminerva_cpu.so,
    driven by oracle/ref_mnist_mlp.cpp through NArray / DagScheduler / CpuDevice): trains, and its softmax is a softmax."""
    r = orc.run_reference_mlp(mb=32, steps=3, warmup=1)
    if r is None:
        pytest.skip("oracle/_ref was never built (no /root/reference on this box)")
    assert r["images_per_s"] > 0 and abs(r["softmax_column_sum"] - 1.0) < 1e-5 and r["cpu_worker_threads"] == 4
