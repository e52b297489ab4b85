"""CPU twin of the owl API (minerva_b200/owl) over the parity oracle -- TEST INFRASTRUCTURE ONLY.

Lets the very same owl.net graph (minerva_b200/owl/net) run on the CPU restatement, for (a) the
end-to-end parity test of a training step, (b) the world_size-2 gloo tests of the data-parallel
trainer and (c) bench.py's cpu_baseline / --impl reference legs.  The product never imports it:
the net takes its backend by injection.  Where the reference has CPU code (minerva/op/impl/basic.cpp)
and oracle/_ref is built, `use_reference(True)` routes those ops through the reference itself.
"""
import numpy as np

from . import pyoracle as orc

_seed = [0x5EED, 0, 0]
_use_ref = [False]
_tf32 = [False]


def set_tf32_operands(flag):
    """True: MatMult and the three convolution directions round their two operands to TF32 first (orc.tf32_round) and
    then run the reference's fp32 algorithm -- the arithmetic north_star specifies for the tensor-core ops.  Whole-step
    parity tests compare the GPU with this AND with the plain fp32 oracle (a tiny batch amplifies the operand rounding
    through ReLU / arg-max switches, which is a property of TF32, not of a kernel)."""
    _tf32[0] = bool(flag)


def _op(a):
    return orc.tf32_round(a) if _tf32[0] else a


def use_reference(flag):
    _use_ref[0] = bool(flag) and orc.have_ref()
    return _use_ref[0]


def set_seed(seed):
    _seed[0] = int(seed) & 0xFFFFFFFF
    _seed[1] = 0


def set_rank_salt(rank):
    _seed[2] = int(rank) & 0xFFFF


def _next_seed(salted=False):
    _seed[1] += 1
    s = (_seed[0] * 0x9E3779B1 + _seed[1] * 0x85EBCA77) & 0xFFFFFFFF
    return (s ^ (_seed[2] * 0xC2B2AE35)) & 0xFFFFFFFF if salted else s


def _prod(s):
    p = 1
    for v in s:
        p *= int(v)
    return p


class ConvInfo:
    def __init__(self, ph=0, pw=0, sv=1, sh=1):
        self.pad_height, self.pad_width, self.stride_vertical, self.stride_horizontal = ph, pw, sv, sh


class NArray:
    def __init__(self, a, shape):
        self.a = np.ascontiguousarray(a, np.float32).reshape(-1)
        self._shape = [int(s) for s in shape]

    @property
    def shape(self):
        return list(self._shape)

    @property
    def size(self):
        return self.a.size

    def as_torch(self):
        import torch
        return torch.from_numpy(self.a)

    def _bin(self, rhs, fn, ref_op, op):
        if self._shape == rhs._shape:
            if _use_ref[0]:
                return NArray(orc.Ref.arithmetic(ref_op, self.a, rhs.a), self._shape)
            return NArray(fn(self.a, rhs.a), self._shape)
        m, n = self._shape
        rep = [i for i in range(2) if self._shape[i] != rhs._shape[i]]
        assert len(rep) == 1 and rhs._shape[rep[0]] == 1
        if _use_ref[0]:
            return NArray(orc.Ref.norm_arithmetic(ref_op, rep[0], self.a, rhs.a, m, n), self._shape)
        f = orc.norm_on_col if rep[0] == 0 else orc.norm_on_row
        return NArray(f(op, self.a, rhs.a, m, n), self._shape)

    def __add__(self, r):
        return self._bin(r, orc.add, "add", "add") if isinstance(r, NArray) else NArray(orc.const_add(self.a, r), self._shape)

    __radd__ = __add__

    def __sub__(self, r):
        return self._bin(r, orc.sub, "sub", "sub") if isinstance(r, NArray) else NArray(orc.const_sub(self.a, r), self._shape)

    def __rsub__(self, l):
        return NArray(orc.left_const_sub(self.a, l), self._shape)

    def __mul__(self, r):
        if isinstance(r, NArray):
            m, k, n = self._shape[0], self._shape[1], r._shape[1]
            assert k == r._shape[0]
            f = orc.Ref.matmult if _use_ref[0] else orc.matmult
            return NArray(f(_op(self.a), _op(r.a), m, n, k), [m, n])
        return NArray(orc.scale(self.a, r), self._shape)

    def __rmul__(self, l):
        return NArray(orc.scale(self.a, l), self._shape)

    def __truediv__(self, r):
        return self._bin(r, orc.dot_div, "div", "div") if isinstance(r, NArray) else NArray(orc.const_div(self.a, r), self._shape)

    def __neg__(self):
        return NArray(orc.elewise_negative(self.a), self._shape)

    def trans(self):
        m, n = self._shape
        f = orc.Ref.transpose if _use_ref[0] else orc.transpose
        return NArray(f(self.a, m, n), [n, m])

    def reshape(self, s):
        assert _prod(s) == self.size
        return NArray(self.a.copy(), s)

    def _as_2d(self, dim):
        nd = len(self._shape)
        if nd == 2:
            return self._shape[0], self._shape[1], dim
        if dim == 0:
            return self._shape[0], _prod(self._shape[1:]), 0
        assert dim == nd - 1
        return _prod(self._shape[:-1]), self._shape[-1], 1

    def _reduce(self, dim, kind):
        m, n, d = self._as_2d(dim)
        osh = list(self._shape)
        osh[dim] = 1
        if kind == "argmax":
            f = orc.max_index_on_col if d == 0 else orc.max_index_on_row
            return NArray(f(self.a, m, n), osh)
        if _use_ref[0]:
            return NArray(orc.Ref.reduction(kind, d, self.a, m, n), osh)
        f = orc.reduction_on_col if d == 0 else orc.reduction_on_row
        return NArray(f(kind, self.a, m, n), osh)

    def sum(self, dim):
        return self._reduce(dim, "sum")

    def max(self, dim):
        return self._reduce(dim, "max")

    def max_index(self, dim):
        return self._reduce(dim, "argmax")

    def count_zero(self):
        return int((self.a == 0).sum())

    def to_numpy(self):
        return self.a.reshape(tuple(reversed(self._shape))).copy()

    def wait_for_eval(self):
        pass

    # statics used by the net
    @staticmethod
    def sigm_back(diff, top, bottom):
        return NArray(orc.sigmoid_backward(bottom.a, top.a, diff.a), diff._shape)

    @staticmethod
    def tanh_back(diff, top, bottom):
        return NArray(orc.tanh_backward(bottom.a, top.a, diff.a), diff._shape)

    @staticmethod
    def sgd_update(w, delta, grad, momentum, lr_over_batch, lr_times_wd):
        w2, d2 = orc.sgd_momentum_update(w.a, delta.a, grad.a, momentum, lr_over_batch, lr_times_wd)
        w.a[:] = w2
        delta.a[:] = d2


def zeros(shape):
    return NArray(np.zeros(_prod(shape), np.float32), shape)


def ones(shape):
    return NArray(np.ones(_prod(shape), np.float32), shape)


def randn(shape, mu, var):
    return NArray(orc.randn(_prod(shape), _next_seed(), mu, var), shape)


def randb(shape, prob):
    return NArray(orc.rand_bernoulli(_prod(shape), _next_seed(salted=True), prob), shape)


def from_numpy(n):
    n = np.require(n, dtype=np.float32, requirements=["C"])
    return NArray(n.reshape(-1).copy(), list(reversed(n.shape)))


def concat(arrays, dim):
    osh = list(arrays[0].shape)
    osh[dim] = sum(a.shape[dim] for a in arrays)
    inner_unit, outer = _prod(osh[:dim]), _prod(osh[dim + 1:])
    out = np.zeros(_prod(osh), np.float32)
    off = 0
    for a in arrays:
        inner = inner_unit * a.shape[dim]
        orc.copy_strided(a.a, out.size, inner, outer, inner, inner_unit * osh[dim], dst_off=off, dst=out)
        off += inner
    return NArray(out, osh)


def slice(src, slice_dim, st_off, slice_count):  # noqa: A001
    osh = list(src.shape)
    osh[slice_dim] = slice_count
    inner_unit, outer = _prod(osh[:slice_dim]), _prod(osh[slice_dim + 1:])
    out = orc.copy_strided(src.a, _prod(osh), inner_unit * slice_count, outer, inner_unit * src.shape[slice_dim],
                           inner_unit * slice_count, src_off=inner_unit * st_off)
    return NArray(out, osh)


def wait_for_all():
    pass


class _Ele:
    @staticmethod
    def mult(x, y):
        return x._bin(y, orc.dot_mult, "mult", "mult")

    @staticmethod
    def exp(x):
        return NArray(orc.elewise_exp(x.a), x.shape)

    @staticmethod
    def ln(x):
        return NArray(orc.elewise_ln(x.a), x.shape)

    @staticmethod
    def relu(x):
        f = (lambda a: orc.Ref.activation("relu", a)) if _use_ref[0] else orc.relu_forward
        return NArray(f(x.a), x.shape)

    @staticmethod
    def sigm(x):
        return NArray(orc.sigmoid_forward(x.a), x.shape)

    @staticmethod
    def tanh(x):
        return NArray(orc.tanh_forward(x.a), x.shape)

    @staticmethod
    def relu_back(y, x):
        return NArray(orc.relu_backward(x.a, x.a, y.a), y.shape)


class _PoolOp:
    max, avg = "max", "avg"


class _SoftOp:
    instance, channel = "instance", "channel"


class _Co:
    pool_op, soft_op = _PoolOp, _SoftOp

    @staticmethod
    def softmax(x, op="instance"):
        shp = x.shape
        if len(shp) != 4:
            shp4 = shp[0:-1] + [1] * (4 - len(shp)) + [shp[-1]]
        else:
            shp4 = shp
        W, H, C, N = shp4
        f = orc.instance_softmax_forward if op == "instance" else orc.channel_softmax_forward
        return NArray(f(x.a, N, C, H, W), shp)

    class Convolver:
        def __init__(self, pad_h, pad_w, stride_v, stride_h):
            self.g = (pad_h, pad_w, stride_v, stride_h)

        def _geo(self, x, w):
            W, H, Ci, N = x.shape
            fw, fh, _, Co = w.shape
            return (N, Ci, Co, H, W) + self.g + (fh, fw)

        def ff(self, x, w, b):
            geo = self._geo(x, w)
            N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw = geo
            y = orc.conv_forward(_op(x.a), _op(w.a), b.a, *geo)
            return NArray(y, [orc.conv_out(W, pw, fw, sh), orc.conv_out(H, ph, fh, sv), Co, N])

        def bp(self, y, x, w):
            return NArray(orc.conv_backward_data(_op(y.a), _op(w.a), *self._geo(x, w)), x.shape)

        def weight_grad(self, y, x, w):
            return NArray(orc.conv_backward_filter(_op(x.a), _op(y.a), *self._geo(x, w)), w.shape)

        def bias_grad(self, y):
            W, H, C, N = y.shape
            return NArray(orc.conv_backward_bias(y.a, N, C, H, W), [C])

    class Pooler:
        def __init__(self, h, w, stride_v, stride_h, pad_h=0, pad_w=0, op="max"):
            self.g = (stride_v, stride_h, h, w, pad_h, pad_w)
            self.kind = "max" if op == "max" else "average"

        def ff(self, x):
            W, H, C, N = x.shape
            sv, sh, wh, ww, ph, pw = self.g
            y = getattr(orc, self.kind + "_pooling_forward")(x.a, N, C, H, W, *self.g)
            return NArray(y, [orc.pooled_size(W, pw, ww, sh), orc.pooled_size(H, ph, wh, sv), C, N])

        def bp(self, y, ff_y, ff_x):
            W, H, C, N = ff_x.shape
            return NArray(getattr(orc, self.kind + "_pooling_backward")(ff_x.a, ff_y.a, y.a, N, C, H, W, *self.g), ff_x.shape)

    class Lrner:
        def __init__(self, local_size, alpha, beta):
            self.p = (local_size, alpha, beta)

        def ff(self, x, scale):
            W, H, C, N = x.shape
            y, sc = orc.lrn_forward(x.a, *self.p, N, C, W, H)
            scale.a[:] = sc      # written in place, like the reference
            return NArray(y, x.shape)

        def bp(self, bottom, top, scale, top_diff):
            W, H, C, N = bottom.shape
            return NArray(orc.lrn_backward(bottom.a, top.a, scale.a, top_diff.a, *self.p, N, C, W, H), bottom.shape)


class Backend:
    """What Net(backend=...) expects: .owl, .co, .ele"""

    def __init__(self):
        import sys
        self.owl = sys.modules[__name__]
        self.co = _Co
        self.ele = _Ele
