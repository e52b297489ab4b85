"""numpy front end of the parity oracle (oracle/mnv_oracle.c) and of the compiled reference
(oracle/_ref/libminerva_ref.so, built from /root/reference/minerva/op/impl/basic.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing in minerva_b200/ imports this module.

Every function takes/returns C-contiguous float32 numpy arrays laid out exactly like the device
buffers (NCHW images, KCRS filters, column-major matrices stored as flat arrays).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libmnv_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libminerva_ref.so")
REF_MLP = os.path.join(_HERE, "_ref", "ref_mnist_mlp")     # configs[0] through the reference's own NArray / DagScheduler / CpuDevice stack

F = C.POINTER(C.c_float)


def build(force=False):
    """Compile the oracle (and, when /root/reference is present, the reference .so)."""
    if force or not os.path.exists(_LIB) or (
            os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "mnv_oracle.c"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libmnv_oracle.so"])
    if os.path.isdir("/root/reference/minerva") and (force or not os.path.exists(_REF) or not os.path.exists(REF_MLP)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def ref():
    """The compiled reference, or None when it was never built (no /root/reference, no prebuilt)."""
    global _ref
    if _ref is None and os.path.exists(_REF):
        _ref = C.CDLL(_REF)
    return _ref


def run_reference_mlp(mb=256, steps=20, warmup=3, timeout=600):
    """BASELINE configs[0] through the reference stack (oracle/ref_mnist_mlp.cpp): -> the program's JSON line as a dict,
    or None when oracle/_ref was never built."""
    import json
    build()
    if not os.path.exists(REF_MLP):
        return None
    out = subprocess.run([REF_MLP, str(mb), str(steps), str(warmup)], capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError("ref_mnist_mlp failed: " + out.stderr[-400:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def have_ref():
    build()
    return ref() is not None


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return a.ctypes.data_as(F)


def _call(name, *args):
    fn = getattr(lib(), name)
    fn.restype = None
    conv = []
    for a in args:
        if isinstance(a, np.ndarray):
            conv.append(a.ctypes.data_as(C.c_void_p))
        elif isinstance(a, float):
            conv.append(C.c_float(a))
        elif isinstance(a, _SizeT):
            conv.append(C.c_size_t(int(a)))
        else:
            conv.append(C.c_int(int(a)))
    fn(*conv)


class _SizeT(int):
    pass


def _n(a):
    return _SizeT(a.size)


# ---- elementwise ---------------------------------------------------------------------------
def _binary(name):
    def f(a, b):
        a, b = _f(a), _f(b)
        c = np.empty_like(a)
        _call(name, a, b, c, _n(a))
        return c
    return f


add = _binary("orc_add")
sub = _binary("orc_sub")
dot_mult = _binary("orc_dot_mult")
dot_div = _binary("orc_dot_div")


def _const(name):
    def f(x, v):
        x = _f(x)
        y = np.empty_like(x)
        _call(name, x, y, float(v), _n(x))
        return y
    return f


const_add = _const("orc_const_add")
const_sub = _const("orc_const_sub")
left_const_sub = _const("orc_left_const_sub")
const_div = _const("orc_const_div")
left_const_div = _const("orc_left_const_div")


def scale(x, v):
    x = _f(x)
    y = np.empty_like(x)
    _call("orc_scale", x, y, _n(x), float(v))
    return y


def _unary(name):
    def f(x):
        x = _f(x)
        y = np.empty_like(x)
        _call(name, x, y, _n(x))
        return y
    return f


elewise_exp = _unary("orc_elewise_exp")
elewise_ln = _unary("orc_elewise_ln")
elewise_negative = _unary("orc_elewise_negative")
sigmoid_forward = _unary("orc_sigmoid_forward")
relu_forward = _unary("orc_relu_forward")
tanh_forward = _unary("orc_tanh_forward")


def _act_back(name):
    def f(x, y, dy):
        x, y, dy = _f(x), _f(y), _f(dy)
        dx = np.empty_like(dy)
        _call(name, x, y, dy, dx, _n(dy))
        return dx
    return f


sigmoid_backward = _act_back("orc_sigmoid_backward")
relu_backward = _act_back("orc_relu_backward")
tanh_backward = _act_back("orc_tanh_backward")

OPS = {"add": 0, "sub": 1, "mult": 2, "div": 3}


# ---- 2-D column-major {m,n}: flat arrays of m*n floats, element (i,j) at i + j*m -----------
def norm_on_col(op, mat, vec, m, n):
    mat, vec = _f(mat), _f(vec)
    res = np.empty_like(mat)
    _call("orc_norm_on_col", OPS[op], mat, vec, res, m, n)
    return res


def norm_on_row(op, mat, vec, m, n):
    mat, vec = _f(mat), _f(vec)
    res = np.empty_like(mat)
    _call("orc_norm_on_row", OPS[op], mat, vec, res, m, n)
    return res


def reduction_on_col(kind, x, m, n):
    x = _f(x)
    out = np.empty(n, np.float32)
    _call("orc_reduction_on_col", int(kind == "max"), x, out, m, n)
    return out


def reduction_on_row(kind, x, m, n):
    x = _f(x)
    out = np.empty(m, np.float32)
    _call("orc_reduction_on_row", int(kind == "max"), x, out, m, n)
    return out


def max_index_on_col(x, m, n):
    x = _f(x)
    out = np.empty(n, np.float32)
    _call("orc_max_index_on_col", x, out, m, n)
    return out


def max_index_on_row(x, m, n):
    x = _f(x)
    out = np.empty(m, np.float32)
    _call("orc_max_index_on_row", x, out, m, n)
    return out


def matmult(a, b, m, n, k):
    a, b = _f(a), _f(b)
    c = np.empty(m * n, np.float32)
    _call("orc_matmult", a, b, c, m, n, k)
    return c


def transpose(a, m, n):
    a = _f(a)
    c = np.empty_like(a)
    _call("orc_transpose", a, c, m, n)
    return c


def copy_strided(src, dst_size, inner, outer, src_stride, dst_stride, src_off=0, dst_off=0, dst=None):
    src = _f(src)
    if dst is None:
        dst = np.zeros(dst_size, np.float32)
    _call("orc_copy_strided", src[src_off:], dst[dst_off:], _SizeT(inner), _SizeT(outer),
          _SizeT(src_stride), _SizeT(dst_stride))
    return dst


def fill(n, v):
    out = np.empty(n, np.float32)
    _call("orc_fill", out, _SizeT(n), float(v))
    return out


# ---- softmax ---------------------------------------------------------------------------------
def _softmax_f(name):
    def f(x, N, Cc, H, W):
        x = _f(x)
        y = np.empty_like(x)
        _call(name, x, y, N, Cc, H, W)
        return y
    return f


def _softmax_b(name):
    def f(dy, y, N, Cc, H, W):
        dy, y = _f(dy), _f(y)
        dx = np.empty_like(dy)
        _call(name, dy, y, dx, N, Cc, H, W)
        return dx
    return f


instance_softmax_forward = _softmax_f("orc_instance_softmax_forward")
channel_softmax_forward = _softmax_f("orc_channel_softmax_forward")
instance_softmax_backward = _softmax_b("orc_instance_softmax_backward")
channel_softmax_backward = _softmax_b("orc_channel_softmax_backward")


# ---- convolution (x: N,Ci,H,W ; w: Co,Ci,fh,fw) ------------------------------------------------
def conv_out(x, pad, f, stride):
    return (x + 2 * pad - f) // stride + 1


def conv_forward(x, w, b, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw):
    x, w, b = _f(x), _f(w), _f(b)
    y = np.empty(N * Co * conv_out(H, ph, fh, sv) * conv_out(W, pw, fw, sh), np.float32)
    _call("orc_conv_forward", x, w, b, y, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw)
    return y


def conv_backward_data(dy, w, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw):
    dy, w = _f(dy), _f(w)
    dx = np.empty(N * Ci * H * W, np.float32)
    _call("orc_conv_backward_data", dy, w, dx, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw)
    return dx


def conv_backward_filter(x, dy, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw):
    x, dy = _f(x), _f(dy)
    dw = np.empty(Co * Ci * fh * fw, np.float32)
    _call("orc_conv_backward_filter", x, dy, dw, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw)
    return dw


def conv_backward_bias(dy, N, Cc, H, W):
    dy = _f(dy)
    db = np.empty(Cc, np.float32)
    _call("orc_conv_backward_bias", dy, db, N, Cc, H, W)
    return db


# ---- pooling ---------------------------------------------------------------------------------
def pooled_size(x, pad, window, stride):
    p = (x + 2 * pad - window + stride - 1) // stride + 1
    if 0 <= (p - 1) * stride - x - pad:
        p -= 1
    return p


def _pool_f(name):
    def f(x, N, Cc, H, W, sv, sh, wh, ww, ph, pw):
        x = _f(x)
        y = np.empty(N * Cc * pooled_size(H, ph, wh, sv) * pooled_size(W, pw, ww, sh), np.float32)
        _call(name, x, y, N, Cc, H, W, sv, sh, wh, ww, ph, pw)
        return y
    return f


def _pool_b(name):
    def f(x, y, dy, N, Cc, H, W, sv, sh, wh, ww, ph, pw):
        x, y, dy = _f(x), _f(y), _f(dy)
        dx = np.empty(N * Cc * H * W, np.float32)
        _call(name, x, y, dy, dx, N, Cc, H, W, sv, sh, wh, ww, ph, pw)
        return dx
    return f


max_pooling_forward = _pool_f("orc_max_pooling_forward")
average_pooling_forward = _pool_f("orc_average_pooling_forward")
max_pooling_backward = _pool_b("orc_max_pooling_backward")
average_pooling_backward = _pool_b("orc_average_pooling_backward")


# ---- LRN ---------------------------------------------------------------------------------------
def lrn_forward(x, local_size, alpha, beta, N, Cc, W, H):
    x = _f(x)
    scale_ = np.empty_like(x)
    out = np.empty_like(x)
    _call("orc_lrn_forward", x, scale_, out, local_size, float(alpha), float(beta), N, Cc, W, H)
    return out, scale_


def lrn_backward(bottom, top, scale_, top_diff, local_size, alpha, beta, N, Cc, W, H):
    bottom, top, scale_, top_diff = _f(bottom), _f(top), _f(scale_), _f(top_diff)
    out = np.empty_like(bottom)
    _call("orc_lrn_backward", bottom, top, scale_, top_diff, out, local_size, float(alpha),
          float(beta), N, Cc, W, H)
    return out


# ---- generators / update -------------------------------------------------------------------
def rand_bernoulli(n, seed, p):
    out = np.empty(n, np.float32)
    fn = lib().orc_rand_bernoulli
    fn.restype = None
    fn(out.ctypes.data_as(C.c_void_p), C.c_size_t(n), C.c_uint(seed), C.c_float(p))
    return out


def randn(n, seed, mean, sd):
    out = np.empty(n, np.float32)
    fn = lib().orc_randn
    fn.restype = None
    fn(out.ctypes.data_as(C.c_void_p), C.c_size_t(n), C.c_uint(seed), C.c_float(mean), C.c_float(sd))
    return out


def sgd_momentum_update(w, delta, grad, mom, lr_over_batch, lr_times_wd):
    w, delta, grad = _f(w).copy(), _f(delta).copy(), _f(grad)
    _call("orc_sgd_momentum_update", w, delta, grad, _n(w), float(mom), float(lr_over_batch),
          float(lr_times_wd))
    return w, delta


def tf32_round(a):
    """fp32 -> TF32 operand rounding as the tensor-core operand path does it (cvt.rna.tf32.f32: nearest, ties away
    from zero; 10 explicit mantissa bits kept).  Used by the whole-step parity tests to state what "TF32 conv / GEMM with
    fp32 accumulation" (north_star) computes: the reference's fp32 algorithm on TF32-rounded operands."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    finite = (u & 0x7F800000) != 0x7F800000
    r = np.where(finite, (u + 0x1000) & 0xFFFFE000, u).astype(np.uint32)
    return r.view(np.float32)


# ---- the compiled reference (oracle/_ref) ---------------------------------------------------
class Ref:
    """Thin numpy wrappers over the reference's own basic:: functions."""

    @staticmethod
    def _c(name, *args):
        fn = getattr(ref(), name)
        fn.restype = None
        conv = []
        for a in args:
            if isinstance(a, np.ndarray):
                conv.append(a.ctypes.data_as(C.c_void_p))
            elif isinstance(a, float):
                conv.append(C.c_float(a))
            else:
                conv.append(C.c_int(int(a)))
        fn(*conv)

    @classmethod
    def arithmetic(cls, op, a, b):
        a, b = _f(a), _f(b)
        c = np.empty_like(a)
        cls._c("ref_arithmetic", OPS[op], a, b, c, a.size)
        return c

    @classmethod
    def arithmetic_const(cls, op, side, val, x):
        x = _f(x)
        y = np.empty_like(x)
        cls._c("ref_arithmetic_const", OPS[op], side, float(val), x, y, x.size)
        return y

    @classmethod
    def elewise(cls, kind, x):
        x = _f(x)
        y = np.empty_like(x)
        cls._c("ref_elewise", {"exp": 0, "ln": 1, "negative": 2}[kind], x, y, x.size)
        return y

    @classmethod
    def matmult(cls, a, b, m, n, k):
        a, b = _f(a), _f(b)
        c = np.empty(m * n, np.float32)
        cls._c("ref_matmult", a, b, c, m, n, k)
        return c

    @classmethod
    def transpose(cls, a, m, n):
        a = _f(a)
        c = np.empty_like(a)
        cls._c("ref_transpose", a, c, m, n)
        return c

    @classmethod
    def reduction(cls, kind, dim, x, m, n):
        x = _f(x)
        out = np.empty(n if dim == 0 else m, np.float32)
        cls._c("ref_reduction", int(kind == "max"), dim, x, out, m, n)
        return out

    @classmethod
    def max_index(cls, dim, x, m, n):
        x = _f(x)
        out = np.empty(n if dim == 0 else m, np.float32)
        cls._c("ref_max_index", dim, x, out, m, n)
        return out

    @classmethod
    def norm_arithmetic(cls, op, dim, mat, vec, m, n):
        mat, vec = _f(mat), _f(vec)
        res = np.empty_like(mat)
        cls._c("ref_norm_arithmetic", OPS[op], dim, mat, vec, res, m, n)
        return res

    @classmethod
    def activation(cls, kind, x):
        x = _f(x)
        y = np.empty_like(x)
        cls._c("ref_%s_forward" % kind, x, y, x.size)
        return y

    @classmethod
    def softmax_forward(cls, x, w, h, c, n):
        x = _f(x)
        y = np.empty_like(x)
        cls._c("ref_softmax_forward", x, y, w, h, c, n)
        return y

    @classmethod
    def fill(cls, n, v):
        out = np.empty(n, np.float32)
        cls._c("ref_fill", out, n, float(v))
        return out

    @staticmethod
    def flatten(dims, idx):
        fn = ref().ref_flatten
        fn.restype = C.c_long
        d = (C.c_int * len(dims))(*dims)
        i = (C.c_int * len(idx))(*idx)
        return fn(d, i, len(dims))
