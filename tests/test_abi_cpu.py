"""No-GPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/mnv.h
declares, the header covers the reference's function table, argument errors are reported without touching a
device, and the product never imports the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from minerva_b200 import _lib, build
    build.build()
    protos = _lib.parse_header()
    assert len(protos) >= 60
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.mnv_abi_version() >= 1
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for forbidden in ("cublas", "cudnn", "curand"):   # north_star: none of these behind the ABI
        assert forbidden not in nm.lower()
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout.lower()
    for forbidden in ("cublas", "cudnn", "curand", "libtorch"):
        assert forbidden not in ldd


def test_product_library_has_no_tuning_state():
    """SURVEY 8b "no global mutable state": the runtime-settable options and the SIMT checker exist only in the tuning
    build (include/mnv_debug.h); the product library does not export the hook."""
    from minerva_b200 import _lib, build
    build.build()
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "mnv_debug_set_option" not in nm
    nm_t = subprocess.run(["nm", "-D", "--defined-only", _lib.TUNING_LIB_PATH], capture_output=True, text=True).stdout
    assert "mnv_debug_set_option" in nm_t
    assert "mnv_debug_set_option" not in _lib.parse_header()
    hdr = open(os.path.join(ROOT, "include", "mnv_debug.h")).read()
    assert "int mnv_debug_set_option(const char* key, int value);" in hdr
    tun = _lib.load_tuning()
    assert tun.mnv_debug_set_option(b"no_such_key", 1) == -1
    assert tun.mnv_debug_set_option(b"no_tail", 0) == 0
    for name in _lib.parse_header():
        assert hasattr(tun, name), name


def test_header_covers_reference_table():
    """One entry per row of minerva/op/impl/cuda/cuda_perform.h:12-76 (53 functions)."""
    from minerva_b200 import _lib
    have = set(_lib.parse_header())
    rows = """dot_mult dot_div add copy sub matmult scale transpose const_add left_const_sub left_const_div
    norm_add_on_col norm_sub_on_col norm_mult_on_col norm_div_on_col norm_add_on_row norm_sub_on_row norm_mult_on_row
    norm_div_on_row reduction_sum_on_col reduction_max_on_col reduction_sum_on_row reduction_max_on_row max_index_on_col
    max_index_on_row reshape elewise_exp elewise_ln elewise_negative conv_forward conv_backward_data conv_backward_filter
    conv_backward_bias instance_softmax_forward channel_softmax_forward instance_softmax_backward channel_softmax_backward
    sigmoid_forward relu_forward tanh_forward sigmoid_backward relu_backward tanh_backward max_pooling_forward
    average_pooling_forward max_pooling_backward average_pooling_backward randn rand_bernoulli fill lrn_forward lrn_backward
    select""".split()
    assert len(rows) == 53
    missing = [r for r in rows if "mnv_" + r not in have]
    assert not missing, missing


def test_argument_errors_without_a_device():
    from minerva_b200 import _lib
    lib = _lib.load()
    assert lib.mnv_add(None, None, None, 16, None) == -1           # MNV_EINVAL: null pointers
    assert lib.mnv_add(None, None, None, 0, None) == 0             # empty input is a no-op
    assert lib.mnv_transpose(None, None, -1, 4, None) == -1
    assert lib.mnv_max_pooling_forward(None, None, 1, 1, 4, 4, 1, 1, 3, 3, 3, 3, None) == -2   # pad >= window
    assert lib.mnv_conv_forward(None, None, None, None, 1, 1, 1, 2, 2, 0, 0, 1, 1, 3, 3, None, 0, None) == -1
    assert lib.mnv_pooled_size(4, 1, 3, 2) == 3 and lib.mnv_pooled_size(4, 2, 4, 3) == 2   # reference goldens
    assert lib.mnv_workspace_bytes_hint() > 0
    with pytest.raises(_lib.MnvError):
        _lib.check(-3, "x")


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|pyoracle|owl_cpu", re.M)
    for base, _, files in os.walk(os.path.join(ROOT, "minerva_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(base, f)).read()
                assert not pat.search(src), os.path.join(base, f)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from minerva_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.MnvError):
        _lib.load()


def test_reference_side_binding_compiles_against_the_reference_headers():
    """the replacement bodies of minerva/op/impl/cuda.cpp's shims compile against the REFERENCE'S
    own headers (DataList, closures, Context, the declarations in op/impl/cuda.h) and call the C ABI of include/mnv.h."""
    ref = "/root/reference/minerva"
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is not present on this box")
    cmd = ["/usr/bin/g++", "-std=c++11", "-fsyntax-only", "-DHAS_CUDA", "-I" + ref, "-I" + os.path.join(ROOT, "oracle", "shim"),
           "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include", "-I/usr/include/x86_64-linux-gnu",
           os.path.join(ROOT, "tests", "cpp", "reference_binding_stub.cpp")]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
