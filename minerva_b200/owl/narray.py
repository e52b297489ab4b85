"""owl.NArray -- same operators and static methods as the reference binding
(owl/owl/libowl.pyx:51-444 over minerva/narray/*.cpp); each maps to one C-ABI call."""
import ctypes
import weakref

import numpy as np
import torch

from .. import _lib
from . import _runtime as _rt


def _prod(shape):
    p = 1
    for s in shape:
        p *= int(s)
    return p


class ConvInfo:
    """minerva/narray/convolution_info.h:5-16"""
    __slots__ = ("pad_height", "pad_width", "stride_vertical", "stride_horizontal")

    def __init__(self, ph=0, pw=0, sv=1, sh=1):
        self.pad_height, self.pad_width, self.stride_vertical, self.stride_horizontal = ph, pw, sv, sh


class _Algo:
    def __init__(self, name, value):
        self.name, self.value = name, value

    def is_same(self, rhs):
        return self.value == rhs.value

    def __repr__(self):
        return self.name


class _Enum:
    def __init__(self, **kw):
        self._items = {k: _Algo(k, v) for k, v in kw.items()}
        self.__dict__.update(self._items)

    def find(self, a):
        for v in self._items.values():
            if v.is_same(a):
                return v
        raise TypeError("invalid algorithm")


pooling_algo = _Enum(max=0, average=1)
pooling_algo.avg = pooling_algo.average
softmax_algo = _Enum(instance=0, channel=1)
activation_algo = _Enum(sigmoid=0, relu=1, tanh=2)


class PoolingInfo:
    """minerva/narray/convolution_info.h:18-46"""
    __slots__ = ("algorithm", "height", "width", "stride_vertical", "stride_horizontal", "pad_height", "pad_width")

    def __init__(self, algorithm=None, h=0, w=0, sv=1, sh=1, ph=0, pw=0):
        self.algorithm = algorithm or pooling_algo.max
        self.height, self.width, self.stride_vertical, self.stride_horizontal = h, w, sv, sh
        self.pad_height, self.pad_width = ph, pw


def _check(cond, msg):
    if not cond:
        raise _lib.MnvError(msg)   # the reference CHECK-fails (dmlc::Error)


class NArray:
    # `_buf` is the flat device buffer.  `trans()` is lazy: its result holds `_lazy_src` (the untransposed array) and
    # no buffer until something other than a matrix product reads it (`_t`), because both operand orders are native
    # to the tensor-core kernel (mnv_matmult_ex) and the reference's explicit transposes in FullyConnected.bp
    # (owl/owl/net/net.py:612-614) were pure HBM traffic.  `_views` lists pending transposes of this array so that
    # the one in-place op (sgd_update) can materialise them before it overwrites their source.
    # `_twin` = (flat device buffer, ctypes.c_int state): the channels-last copy of a 4-D activation that the convolution
    # calls of one training step share (include/mnv.h, mnv_conv_*_tw): forward fills the bottom's twin and backward-filter
    # reuses it, backward-filter fills top_diff's and backward-data reuses it.  Arrays are immutable once produced (the only
    # in-place op, sgd_update, touches parameters), so a filled twin stays current for the array's lifetime.
    __slots__ = ("_buf", "_shape", "_dev", "_lazy_src", "_views", "_twin", "__weakref__")
    exact_math = False   # owl.set_exact_math(True): exp / ln / sigmoid / tanh return the CPU reference's (glibc's) bits
    use_twins = True     # Net.fuse_conv_twins: False = every convolution call makes its own workspace copy (the plain entries)
    _twin_wanted = {}    # geometry -> mnv_conv_twin_wanted bitmask

    def __init__(self, tensor, shape, dev):
        self._buf, self._shape, self._dev = tensor, [int(s) for s in shape], dev
        self._lazy_src, self._views, self._twin = None, None, None

    @staticmethod
    def _twins(geo, bottom, top_diff, dev):
        """-> [bottom twin ptr, byref(state), top_diff twin ptr, byref(state)] for one convolution geometry
        (None, None where a twin is not wanted / no array given); buffers are allocated on first use."""
        want = NArray._twin_wanted.get(geo)
        if want is None:
            want = NArray._twin_wanted[geo] = _lib.load().mnv_conv_twin_wanted(*geo) if NArray.use_twins else 0
        out = []
        for bit, arr in ((1, bottom), (2, top_diff)):
            if arr is None or not (want & bit) or not NArray.use_twins or arr._dev is not dev:
                out += [None, None]
                continue
            if arr._twin is None:
                W, H, C, N = arr._shape
                nbytes = _lib.load().mnv_conv_twin_bytes(N, C, H, W)
                arr._twin = (torch.empty(nbytes // 4, dtype=torch.float32, device=dev.device), ctypes.c_int(0))
            out += [arr._twin[0].data_ptr(), ctypes.byref(arr._twin[1])]
        return out

    @property
    def _t(self):
        if self._lazy_src is not None:
            self._materialize()
        return self._buf

    def _materialize(self):
        src, self._lazy_src = self._lazy_src, None
        m, n = src._shape
        # the C ABI wants pointers and stream on the CURRENT device: the view may be read first after
        # owl.set_device(other_gpu) (the reference's single-process multi-GPU loop), so materialise on its own device
        with torch.cuda.device(self._dev.index):
            self._buf = torch.empty(max(m * n, 0), dtype=torch.float32, device=self._dev.device)
            NArray._call("mnv_transpose", self._dev, src._on(self._dev).data_ptr(), self._buf.data_ptr(), m, n)

    def _flush_views(self):
        """Materialise pending lazy transposes of this array (called before it is modified in place)."""
        if self._views:
            for ref in self._views:
                v = ref()
                if v is not None and v._lazy_src is self:
                    v._materialize()
            self._views = None

    # ---- plumbing ---------------------------------------------------------------------------
    @property
    def shape(self):
        return list(self._shape)

    @property
    def size(self):
        return _prod(self._shape)

    @staticmethod
    def _new(shape, dev=None):
        dev = dev or _rt.current_device()
        if NArray._placed:       # place_next(): this result goes into a caller-owned buffer
            t = NArray._placed.pop(0)
            if t.numel() != max(_prod(shape), 0):
                NArray._placed = []
                raise _lib.MnvError("place_next: the next result has %d elements, the buffer %d" % (_prod(shape), t.numel()))
            return NArray(t, shape, dev)
        return NArray(torch.empty(max(_prod(shape), 0), dtype=torch.float32, device=dev.device), shape, dev)

    _placed = []

    @staticmethod
    def place_next(*tensors):
        """The next len(tensors) results allocated by NArray ops are written straight into these flat fp32 device
        buffers (in order) instead of fresh memory.  Used by the data-parallel trainer so that gradients are produced
        inside the peer-mapped merge buffer (owl/net/merge.py) -- the role of the reference's explicit output DataShards
        (op/compute_fn.h:9-12), which owl's value-semantics API otherwise hides."""
        NArray._placed = list(tensors)

    def _on(self, dev):
        """Pull a remote input onto `dev` (the reference's DoCopyRemoteData, device.cpp:75-91,209-212)."""
        if self._dev is dev:
            return self._t
        return self._t.to(dev.device, non_blocking=True)

    @staticmethod
    def _call(name, dev, *args, prof_args=None):
        prof = _rt.profiler
        if prof is not None:          # per-op device timing (the reference's ExecutionProfiler role)
            tok = prof.begin(name, prof_args if prof_args is not None else args, dev)
        rc = getattr(_lib.load(), name)(*args, dev.stream_ptr)
        if rc:
            _lib.check(rc, name)
        if prof is not None:
            prof.end(tok, dev)

    def wait_for_eval(self):
        self._dev.stream.synchronize()

    def as_torch(self):
        """The underlying flat device buffer (for torch.distributed collectives)."""
        return self._t

    @staticmethod
    def sgd_update(w, delta, grad, momentum, lr_over_batch, lr_times_wd):
        """the reference's momentum-SGD chain (owl/net/net.py:252-256, ten ops
        and 60 B/param per tensor) as one in-place kernel (20 B/param)."""
        dev = _rt.current_device()
        _check(w._dev is dev and delta._dev is dev, "sgd_update is in place: w and delta must live on this device")
        w._flush_views()
        delta._flush_views()
        NArray._call("mnv_sgd_momentum_update", dev, w._t.data_ptr(), delta._t.data_ptr(), grad._on(dev).data_ptr(),
                     w.size, float(momentum), float(lr_over_batch), float(lr_times_wd))

    class _SgdTensor(ctypes.Structure):      # mnv_sgd_tensor_t (include/mnv.h)
        _fields_ = [("w", ctypes.c_void_p), ("delta", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("n", ctypes.c_size_t),
                    ("lr_over_batch", ctypes.c_float), ("lr_times_wd", ctypes.c_float)]

    @staticmethod
    def sgd_update_multi(entries, momentum):
        """entries: [(w, delta, grad, lr_over_batch, lr_times_wd)] -- every parameter tensor of a net updated by ONE launch
        (mnv_sgd_momentum_update_multi); bit-identical to one sgd_update per tensor."""
        dev = _rt.current_device()
        arr = (NArray._SgdTensor * len(entries))()
        keep = []
        for i, (w, delta, grad, lrb, lrwd) in enumerate(entries):
            _check(w._dev is dev and delta._dev is dev, "sgd_update is in place: w and delta must live on this device")
            w._flush_views()
            delta._flush_views()
            g = grad._on(dev)
            keep.append(g)
            arr[i] = NArray._SgdTensor(w._t.data_ptr(), delta._t.data_ptr(), g.data_ptr(), w.size, float(lrb), float(lrwd))
        NArray._call("mnv_sgd_momentum_update_multi", dev, ctypes.cast(arr, ctypes.c_void_p), len(entries), float(momentum),
                     prof_args=(0, len(entries), float(momentum), sum(e[0].size for e in entries)))

    def start_eval(self):
        pass

    # ---- elementwise arithmetic (narray_elewise.cpp:17-158) ------------------------------------
    def _arith(self, rhs, fn, norm):
        dev = _rt.current_device()
        if self._shape == rhs._shape:
            out = NArray._new(self._shape, dev)
            NArray._call(fn, dev, self._on(dev).data_ptr(), rhs._on(dev).data_ptr(), out._t.data_ptr(), self.size)
            return out
        # NormArithmetic (narray.cpp:195-214): rhs dims of size 1 are replicated; 2-D, one dim
        _check(len(self._shape) == len(rhs._shape) == 2, "NormArithmetic: 2-D operands only on the GPU path")
        rep = [i for i in range(2) if self._shape[i] != rhs._shape[i]]
        _check(len(rep) == 1 and rhs._shape[rep[0]] == 1, "NormArithmetic cannot replicate a dimension that is not 1")
        m, n = self._shape
        out = NArray._new(self._shape, dev)
        name = "mnv_norm_%s_on_%s" % (norm, "col" if rep[0] == 0 else "row")
        NArray._call(name, dev, self._on(dev).data_ptr(), rhs._on(dev).data_ptr(), out._t.data_ptr(), m, n)
        return out

    def _const(self, fn, val):
        dev = _rt.current_device()
        out = NArray._new(self._shape, dev)
        if fn == "mnv_scale":
            NArray._call(fn, dev, self._on(dev).data_ptr(), out._t.data_ptr(), self.size, float(val))
        else:
            NArray._call(fn, dev, self._on(dev).data_ptr(), out._t.data_ptr(), float(val), self.size)
        return out

    def __add__(self, rhs):
        return self._arith(rhs, "mnv_add", "add") if isinstance(rhs, NArray) else self._const("mnv_const_add", rhs)

    __radd__ = __add__

    def __sub__(self, rhs):
        if isinstance(rhs, NArray):
            return self._arith(rhs, "mnv_sub", "sub")
        return self._const("mnv_const_add", -float(rhs))       # cuda.cpp:184-186

    def __rsub__(self, lhs):
        return self._const("mnv_left_const_sub", lhs)

    def __mul__(self, rhs):
        if isinstance(rhs, NArray):                              # matrix product (narray.cpp:116-123)
            return NArray.matmult(self, rhs)
        return self._const("mnv_scale", rhs)

    def __rmul__(self, lhs):
        return self._const("mnv_scale", lhs)

    def __truediv__(self, rhs):
        if isinstance(rhs, NArray):
            return self._arith(rhs, "mnv_dot_div", "div")
        return self._const("mnv_const_div", rhs)                # IEEE division, not x*(1/v)

    def __rtruediv__(self, lhs):
        return self._const("mnv_left_const_div", lhs)

    __div__, __rdiv__ = __truediv__, __rtruediv__

    def __neg__(self):
        return self._unary("mnv_elewise_negative")

    def _unary(self, fn):
        dev = _rt.current_device()
        out = NArray._new(self._shape, dev)
        NArray._call(fn, dev, self._on(dev).data_ptr(), out._t.data_ptr(), self.size)
        return out

    @staticmethod
    def mult(lhs, rhs):
        return lhs._arith(rhs, "mnv_dot_mult", "mult")

    @staticmethod
    def exp(x):
        return x._unary("mnv_elewise_exp_exact" if NArray.exact_math else "mnv_elewise_exp")

    @staticmethod
    def ln(x):
        return x._unary("mnv_elewise_ln_exact" if NArray.exact_math else "mnv_elewise_ln")

    # ---- activations (narray_elewise.cpp:51-82; argument order (diff, top, bottom)) --------------
    def _act(self, fn):
        dev = _rt.current_device()
        out = NArray._new(self._shape, dev)
        NArray._call(fn, dev, self._on(dev).data_ptr(), out._t.data_ptr(), 1, 1, 1, self.size)
        return out

    @staticmethod
    def _act_back(fn, diff, top, bottom):
        _check(diff._shape == top._shape == bottom._shape, "inputs size mismatch")
        dev = _rt.current_device()
        out = NArray._new(diff._shape, dev)
        NArray._call(fn, dev, bottom._on(dev).data_ptr(), top._on(dev).data_ptr(), diff._on(dev).data_ptr(),
                     out._t.data_ptr(), 1, 1, 1, diff.size)
        return out

    @staticmethod
    def sigm(x):
        return x._act("mnv_sigmoid_forward_exact" if NArray.exact_math else "mnv_sigmoid_forward")

    @staticmethod
    def relu(x):
        return x._act("mnv_relu_forward")

    @staticmethod
    def tanh(x):
        return x._act("mnv_tanh_forward_exact" if NArray.exact_math else "mnv_tanh_forward")

    @staticmethod
    def sigm_back(diff, top, bottom):
        return NArray._act_back("mnv_sigmoid_backward", diff, top, bottom)

    @staticmethod
    def relu_back(diff, top, bottom):
        return NArray._act_back("mnv_relu_backward", diff, top, bottom)

    @staticmethod
    def add_n(arrays):
        """((a0 + a1) + a2) + ... in one pass (mnv_add_n): the bits of the chained `+`."""
        arrays = list(arrays)
        _check(len(arrays) >= 1 and all(a._shape == arrays[0]._shape for a in arrays), "inputs size mismatch")
        if len(arrays) == 1:
            return arrays[0]
        dev = _rt.current_device()
        out = NArray._new(arrays[0]._shape, dev)
        done = 0
        acc = None
        while done < len(arrays):          # at most 8 sources per launch
            chunk = ([acc] if acc is not None else []) + arrays[done:done + (7 if acc is not None else 8)]
            done += len(chunk) - (1 if acc is not None else 0)
            ptrs = (ctypes.c_void_p * len(chunk))(*[a._on(dev).data_ptr() for a in chunk])
            NArray._call("mnv_add_n", dev, ptrs, len(chunk), out._t.data_ptr(), out.size, prof_args=(len(chunk), out.size))
            acc = out
        return out

    @staticmethod
    def relu_back_tw(diff, top, conv_geo):
        """ReLU backward for a 4-D activation that a convolution of geometry `conv_geo` produced: same result as relu_back,
        and the result carries its channels-last twin (the top_diff twin that convolution's backward calls will ask for)."""
        _check(diff._shape == top._shape and len(diff._shape) == 4, "inputs size mismatch")
        dev = _rt.current_device()
        out = NArray._new(diff._shape, dev)
        W, H, C, N = diff._shape
        _, _, tw, ts = NArray._twins(conv_geo, None, out, dev)
        NArray._call("mnv_relu_backward_tw", dev, top._on(dev).data_ptr(), diff._on(dev).data_ptr(), out._t.data_ptr(), N, C, H, W, tw, ts)
        return out

    @staticmethod
    def tanh_back(diff, top, bottom):
        return NArray._act_back("mnv_tanh_backward", diff, top, bottom)

    @staticmethod
    def activation_forward(src, algo):
        exact = NArray.exact_math and algo.name in ("sigmoid", "tanh")
        return src._act("mnv_%s_forward%s" % (algo.name, "_exact" if exact else ""))

    @staticmethod
    def activation_backward(diff, top, bottom, algo):
        return NArray._act_back("mnv_%s_backward" % algo.name, diff, top, bottom)

    # ---- matrix ops ---------------------------------------------------------------------------------
    @staticmethod
    def matmult(lhs, rhs):
        _check(len(lhs._shape) == 2 and len(rhs._shape) == 2, "eligible only for 2D")
        _check(lhs._shape[1] == rhs._shape[0], "size must match")
        dev = _rt.current_device()
        m, k, n = lhs._shape[0], lhs._shape[1], rhs._shape[1]
        out = NArray._new([m, n], dev)
        ta, tb = lhs._lazy_src is not None, rhs._lazy_src is not None
        if ta and tb:
            tb = False                                           # at most one operand stays virtual
        if ta or tb:
            a = lhs._lazy_src if ta else lhs
            b = rhs._lazy_src if tb else rhs
            NArray._call("mnv_matmult_ex", dev, a._on(dev).data_ptr(), b._on(dev).data_ptr(), out._buf.data_ptr(), m, n, k,
                         int(ta), int(tb), dev.ws_ptr, dev.ws_bytes)
            return out
        NArray._call("mnv_matmult", dev, lhs._on(dev).data_ptr(), rhs._on(dev).data_ptr(), out._t.data_ptr(), m, n, k,
                     dev.ws_ptr, dev.ws_bytes)
        return out

    def trans(self):
        """Lazy: the copy happens only if something other than a matrix product reads the result."""
        _check(len(self._shape) == 2, "eligible only for 2D")
        dev = _rt.current_device()
        m, n = self._shape
        out = NArray(None, [n, m], dev)
        src = self
        if src._lazy_src is not None:
            src._materialize()
        out._lazy_src = src
        if src._views is None:
            src._views = []
        src._views = [r for r in src._views if r() is not None]
        src._views.append(weakref.ref(out))
        return out

    def reshape(self, s):
        _check(_prod(s) == self.size, "dimension mismatch")
        dev = _rt.current_device()
        out = NArray._new(s, dev)                                # the reference's Reshape is a full copy (cuda.cpp:312-316)
        NArray._call("mnv_reshape", dev, self._on(dev).data_ptr(), out._t.data_ptr(), self.size * 4)
        return out

    def _as_2d(self, dim):
        """View an N-d reduction over its first or last dim as the 2-D case the kernels support
        (the reference CUDA path is 2-D only, cuda.cpp:268-270)."""
        nd = len(self._shape)
        _check(0 <= dim < nd, "dim out of bound")
        if nd == 2:
            return self._shape[0], self._shape[1], dim
        if dim == 0:
            return self._shape[0], _prod(self._shape[1:]), 0
        _check(dim == nd - 1, "reduction over an inner dimension is not supported on the GPU path")
        return _prod(self._shape[:-1]), self._shape[-1], 1

    def _reduce(self, dim, kind):
        if not isinstance(dim, int):
            _check(len(dim) == 1, "currently do reduction on one dimension only")
            dim = int(list(dim)[0])
        m, n, d = self._as_2d(dim)
        dev = _rt.current_device()
        oshape = list(self._shape)
        oshape[dim] = 1
        out = NArray._new(oshape, dev)
        NArray._call("mnv_%s_on_%s" % (kind, "col" if d == 0 else "row"), dev, self._on(dev).data_ptr(),
                     out._t.data_ptr(), m, n)
        return out

    def sum(self, dim):
        return self._reduce(dim, "reduction_sum")

    def max(self, dim):
        return self._reduce(dim, "reduction_max")

    def max_index(self, dim):
        return self._reduce(dim, "max_index")

    def count_zero(self):
        """Blocking, like the reference (narray_reduction.cpp:64-75)."""
        self._dev.stream.synchronize()
        return int((self._t == 0).sum().item())

    # ---- convolution family (narray/convolution.cpp) -----------------------------------------------
    @staticmethod
    def conv_geo(src, filt, info):
        W, H, Ci, N = src._shape
        fw, fh, _, Co = filt._shape
        return (N, Ci, Co, H, W, info.pad_height, info.pad_width, info.stride_vertical, info.stride_horizontal, fh, fw)

    @staticmethod
    def conv_forward(src, filt, bias, info, relu=False):
        W, H, Ci, N = src._shape
        fw, fh, Ci2, Co = filt._shape
        _check(Ci == Ci2, "#input channels mismatch")
        _check(len(bias._shape) == 1 and bias._shape[0] == Co, "bias size mismatch")
        Wo = (W + 2 * info.pad_width - fw) // info.stride_horizontal + 1
        Ho = (H + 2 * info.pad_height - fh) // info.stride_vertical + 1
        dev = _rt.current_device()
        out = NArray._new([Wo, Ho, Co, N], dev)
        geo = (N, Ci, Co, H, W, info.pad_height, info.pad_width, info.stride_vertical, info.stride_horizontal, fh, fw)
        xt, xs, _, _ = NArray._twins(geo, src, None, dev)
        NArray._call("mnv_conv_forward_tw", dev, src._on(dev).data_ptr(), filt._on(dev).data_ptr(), bias._on(dev).data_ptr(),
                     out._t.data_ptr(), *geo, 1 if relu else 0, xt, xs, dev.ws_ptr, dev.ws_bytes)
        return out

    @staticmethod
    def conv_backward_data(diff, bottom, filt, info):
        W, H, Ci, N = bottom._shape                             # output takes the bottom's shape (convolution.cpp:47)
        fw, fh, _, Co = filt._shape
        _check(diff._shape[2] == Co, "#output channels mismatch")
        dev = _rt.current_device()
        out = NArray._new(bottom._shape, dev)
        geo = (N, Ci, Co, H, W, info.pad_height, info.pad_width, info.stride_vertical, info.stride_horizontal, fh, fw)
        _, _, dt, ds = NArray._twins(geo, None, diff, dev)
        NArray._call("mnv_conv_backward_data_tw", dev, diff._on(dev).data_ptr(), filt._on(dev).data_ptr(), out._t.data_ptr(),
                     *geo, dt, ds, dev.ws_ptr, dev.ws_bytes)
        return out

    @staticmethod
    def conv_backward_filter(diff, bottom, filt, info):
        W, H, Ci, N = bottom._shape
        fw, fh, _, Co = filt._shape
        _check(diff._shape[3] == N, "#images mismatch")
        dev = _rt.current_device()
        out = NArray._new(filt._shape, dev)
        geo = (N, Ci, Co, H, W, info.pad_height, info.pad_width, info.stride_vertical, info.stride_horizontal, fh, fw)
        NArray._call("mnv_conv_backward_filter_tw", dev, bottom._on(dev).data_ptr(), diff._on(dev).data_ptr(), out._t.data_ptr(), None,
                     *geo, *NArray._twins(geo, bottom, diff, dev), dev.ws_ptr, dev.ws_bytes)
        return out

    @staticmethod
    def conv_backward_filter_bias(diff, bottom, filt, info):
        """Both parameter gradients of a convolution in one call (not in the reference API): mnv_conv_backward_filter_bias."""
        W, H, Ci, N = bottom._shape
        fw, fh, _, Co = filt._shape
        _check(diff._shape[3] == N, "#images mismatch")
        dev = _rt.current_device()
        dw, db = NArray._new(filt._shape, dev), NArray._new([Co], dev)
        geo = (N, Ci, Co, H, W, info.pad_height, info.pad_width, info.stride_vertical, info.stride_horizontal, fh, fw)
        NArray._call("mnv_conv_backward_filter_tw", dev, bottom._on(dev).data_ptr(), diff._on(dev).data_ptr(), dw._t.data_ptr(),
                     db._t.data_ptr(), *geo, *NArray._twins(geo, bottom, diff, dev), dev.ws_ptr, dev.ws_bytes)
        return dw, db

    @staticmethod
    def conv_backward_bias(diff):
        W, H, C, N = diff._shape
        dev = _rt.current_device()
        out = NArray._new([C], dev)
        NArray._call("mnv_conv_backward_bias", dev, diff._on(dev).data_ptr(), out._t.data_ptr(), N, C, H, W,
                     dev.ws_ptr, dev.ws_bytes)
        return out

    @staticmethod
    def softmax_forward(src, algo):
        W, H, C, N = src._shape
        dev = _rt.current_device()
        out = NArray._new(src._shape, dev)
        NArray._call("mnv_%s_softmax_forward" % algo.name, dev, src._on(dev).data_ptr(), out._t.data_ptr(), N, C, H, W)
        return out

    @staticmethod
    def softmax_backward(diff, top, algo):
        _check(diff._shape == top._shape, "inputs sizes mismatch")
        W, H, C, N = diff._shape
        dev = _rt.current_device()
        out = NArray._new(diff._shape, dev)
        NArray._call("mnv_%s_softmax_backward" % algo.name, dev, diff._on(dev).data_ptr(), top._on(dev).data_ptr(),
                     out._t.data_ptr(), N, C, H, W)
        return out

    @staticmethod
    def _pool_args(info):
        return (info.stride_vertical, info.stride_horizontal, info.height, info.width, info.pad_height, info.pad_width)

    @staticmethod
    def pooling_forward(src, info):
        W, H, C, N = src._shape
        lib = _lib.load()
        Ho = lib.mnv_pooled_size(H, info.pad_height, info.height, info.stride_vertical)
        Wo = lib.mnv_pooled_size(W, info.pad_width, info.width, info.stride_horizontal)
        dev = _rt.current_device()
        out = NArray._new([Wo, Ho, C, N], dev)
        name = "mnv_%s_pooling_forward" % ("max" if info.algorithm.value == 0 else "average")
        NArray._call(name, dev, src._on(dev).data_ptr(), out._t.data_ptr(), N, C, H, W, *NArray._pool_args(info))
        return out

    @staticmethod
    def pooling_backward(diff, top, bottom, info, relu=False):
        """relu=True (max pooling only; not in the reference signature): `bottom` is a ReLU output and the result is
        also masked by bottom > 0, i.e. ReLU backward is folded into the same kernel (mnv_max_pooling_backward_relu)."""
        _check(diff._shape == top._shape, "inputs sizes mismatch")
        _check(not relu or info.algorithm.value == 0, "the fused ReLU mask exists for max pooling only")
        W, H, C, N = bottom._shape
        dev = _rt.current_device()
        out = NArray._new(bottom._shape, dev)
        name = "mnv_%s_pooling_backward" % ("max" if info.algorithm.value == 0 else "average") + ("_relu" if relu else "")
        NArray._call(name, dev, bottom._on(dev).data_ptr(), top._on(dev).data_ptr(), diff._on(dev).data_ptr(),
                     out._t.data_ptr(), N, C, H, W, *NArray._pool_args(info))
        return out

    @staticmethod
    def pooling_idx_ok(info, shape):
        """3x3 / stride 2 / pad 0 max pooling of planes that fit the staged kernel: what mnv_max_pooling_*_idx handle."""
        W, H, C, N = shape
        return (info.algorithm.value == 0 and
                bool(_lib.load().mnv_max_pooling_idx_supported(N, C, H, W, *NArray._pool_args(info))))

    @staticmethod
    def pooling_forward_idx(src, info):
        """Max pooling that also returns the arg-max bytes (a raw uint8 device buffer; not in the reference API)."""
        W, H, C, N = src._shape
        lib = _lib.load()
        Ho = lib.mnv_pooled_size(H, info.pad_height, info.height, info.stride_vertical)
        Wo = lib.mnv_pooled_size(W, info.pad_width, info.width, info.stride_horizontal)
        dev = _rt.current_device()
        out = NArray._new([Wo, Ho, C, N], dev)
        idx = torch.empty(max(out.size, 1), dtype=torch.uint8, device=dev.device)
        NArray._call("mnv_max_pooling_forward_idx", dev, src._on(dev).data_ptr(), out._t.data_ptr(), idx.data_ptr(), N, C, H, W,
                     *NArray._pool_args(info))
        return out, idx

    @staticmethod
    def pooling_backward_idx(diff, idx, bottom_shape, info, relu_top=None):
        """Backward from (top_diff, arg-max bytes); relu_top (the pooled top) folds ReLU backward in."""
        W, H, C, N = bottom_shape
        dev = _rt.current_device()
        out = NArray._new(bottom_shape, dev)
        NArray._call("mnv_max_pooling_backward_idx", dev, diff._on(dev).data_ptr(), idx.data_ptr(),
                     relu_top._on(dev).data_ptr() if relu_top is not None else None, out._t.data_ptr(), N, C, H, W,
                     *NArray._pool_args(info))
        return out

    @staticmethod
    def lrn_forward(src, scale, local_size, alpha, beta):
        """`scale` is written in place, as in the reference (cuda.cpp:45-59; owl/net/net.py:496-500)."""
        W, H, C, N = src._shape
        dev = _rt.current_device()
        _check(scale._dev is dev and scale._shape == src._shape, "scale must be a same-shape array on this device")
        out = NArray._new(src._shape, dev)
        NArray._call("mnv_lrn_forward", dev, src._on(dev).data_ptr(), scale._t.data_ptr(), out._t.data_ptr(),
                     int(local_size), float(alpha), float(beta), N, C, W, H)
        return out

    @staticmethod
    def lrn_backward(bottom_data, top_data, scale, top_diff, local_size, alpha, beta, relu=False):
        """relu=True (not in the reference signature): `bottom_data` is a ReLU output, ReLU backward is folded in."""
        W, H, C, N = bottom_data._shape
        dev = _rt.current_device()
        out = NArray._new(bottom_data._shape, dev)
        NArray._call("mnv_lrn_backward_relu" if relu else "mnv_lrn_backward", dev, bottom_data._on(dev).data_ptr(), top_data._on(dev).data_ptr(),
                     scale._on(dev).data_ptr(), top_diff._on(dev).data_ptr(), out._t.data_ptr(), int(local_size),
                     float(alpha), float(beta), N, C, W, H)
        return out

    @staticmethod
    def lrn_lite_ok(shape, local_size):
        """Window 5, >= 5 channels, 32-bit offsets: the shapes mnv_lrn_*_lite handle (include/mnv.h)."""
        W, H, C, N = shape
        return int(local_size) == 5 and C >= 5 and (C + 32) * W * H * N < (1 << 31)

    @staticmethod
    def lrn_forward_lite(src, local_size, alpha, beta):
        """LRN forward without the `scale` array (not in the reference API): see mnv_lrn_forward_lite."""
        W, H, C, N = src._shape
        dev = _rt.current_device()
        out = NArray._new(src._shape, dev)
        NArray._call("mnv_lrn_forward_lite", dev, src._on(dev).data_ptr(), out._t.data_ptr(),
                     int(local_size), float(alpha), float(beta), N, C, W, H)
        return out

    @staticmethod
    def lrn_backward_lite(bottom_data, top_diff, local_size, alpha, beta, relu=False):
        """LRN backward from (bottom, top_diff) alone, scale and top recomputed in registers: see mnv_lrn_backward_lite."""
        W, H, C, N = bottom_data._shape
        dev = _rt.current_device()
        out = NArray._new(bottom_data._shape, dev)
        NArray._call("mnv_lrn_backward_lite", dev, bottom_data._on(dev).data_ptr(), top_diff._on(dev).data_ptr(),
                     out._t.data_ptr(), int(local_size), float(alpha), float(beta), N, C, W, H, 1 if relu else 0)
        return out

    # ---- constructors / host transfer ---------------------------------------------------------------
    @staticmethod
    def _filled(shape, val):
        dev = _rt.current_device()
        out = NArray._new(shape, dev)
        NArray._call("mnv_fill", dev, out._t.data_ptr(), out.size, float(val))
        return out

    @staticmethod
    def zeros(s):
        return NArray._filled(s, 0.0)

    @staticmethod
    def ones(s):
        return NArray._filled(s, 1.0)

    @staticmethod
    def randn(s, mean, var):
        dev = _rt.current_device()
        out = NArray._new(s, dev)
        NArray._call("mnv_randn", dev, out._t.data_ptr(), out.size, _rt.next_seed(), float(mean), float(var))
        return out

    @staticmethod
    def randb(s, p):
        dev = _rt.current_device()
        out = NArray._new(s, dev)
        cs = _rt.capture_seeds
        if cs is not None:      # the step is being recorded into a CUDA graph: the key is read on the device at replay time
            add, xor = cs.next(salted=True)
            NArray._call("mnv_rand_bernoulli_ds", dev, out._t.data_ptr(), out.size, cs.word.data_ptr(), add, xor, float(p))
        else:
            NArray._call("mnv_rand_bernoulli", dev, out._t.data_ptr(), out.size, _rt.next_seed(salted=True), float(p))
        return out

    @staticmethod
    def concat(arrays, dim):
        _check(len(arrays) > 1, "Concat more than one narray")
        nd = len(arrays[0]._shape)
        _check(nd - dim <= 2, "Currently only support concat on the last two dims!")   # cuda.cpp:83
        dev = _rt.current_device()
        oshape = list(arrays[0]._shape)
        oshape[dim] = sum(a._shape[dim] for a in arrays)
        out = NArray._new(oshape, dev)
        inner_unit = _prod(oshape[:dim])
        outer = _prod(oshape[dim + 1:])
        dst_stride = inner_unit * oshape[dim]
        off, segs = 0, []
        for a in arrays:
            inner = inner_unit * a._shape[dim]
            segs.append((a._on(dev).data_ptr(), out._t.data_ptr() + 4 * off, inner, inner, dst_stride))
            off += inner
        NArray._copy_segs(dev, segs, outer)
        return out

    class _CopySeg(ctypes.Structure):      # mnv_copy_seg_t (include/mnv.h)
        _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("inner", ctypes.c_size_t), ("src_stride", ctypes.c_size_t),
                    ("dst_stride", ctypes.c_size_t)]

    @staticmethod
    def _copy_segs(dev, segs, outer):
        """Strided block copies sharing `outer`, eight per launch (mnv_copy_strided_n)."""
        for i in range(0, len(segs), 8):
            chunk = segs[i:i + 8]
            if len(chunk) == 1:
                src, dst, inner, ss, ds = chunk[0]
                NArray._call("mnv_copy_strided", dev, src, dst, inner, outer, ss, ds)
                continue
            tab = (NArray._CopySeg * len(chunk))(*[NArray._CopySeg(*c) for c in chunk])
            NArray._call("mnv_copy_strided_n", dev, tab, len(chunk), outer, prof_args=(len(chunk), sum(c[2] for c in chunk) * outer))

    @staticmethod
    def split(src, dim, counts):
        """The consecutive slices of `src` along `dim` with the given extents (the pieces a Concat put together), in one launch."""
        nd = len(src._shape)
        _check(nd - dim <= 2, "Currently only support slice on the last two dims!")
        _check(sum(counts) == src._shape[dim], "split extents must cover the dimension")
        dev = _rt.current_device()
        inner_unit = _prod(src._shape[:dim])
        outer = _prod(src._shape[dim + 1:])
        outs, segs, st = [], [], 0
        base = src._on(dev).data_ptr()
        for cnt in counts:
            oshape = list(src._shape)
            oshape[dim] = cnt
            o = NArray._new(oshape, dev)
            segs.append((base + 4 * inner_unit * st, o._t.data_ptr(), inner_unit * cnt, inner_unit * src._shape[dim], inner_unit * cnt))
            outs.append(o)
            st += cnt
        NArray._copy_segs(dev, segs, outer)
        return outs

    @staticmethod
    def slice(src, slice_dim, st_off, slice_count):
        nd = len(src._shape)
        _check(nd - slice_dim <= 2, "Currently only support slice on the last two dims!")
        dev = _rt.current_device()
        oshape = list(src._shape)
        oshape[slice_dim] = slice_count
        out = NArray._new(oshape, dev)
        inner_unit = _prod(oshape[:slice_dim])
        outer = _prod(oshape[slice_dim + 1:])
        NArray._call("mnv_copy_strided", dev, src._on(dev).data_ptr() + 4 * inner_unit * st_off, out._t.data_ptr(),
                     inner_unit * slice_count, outer, inner_unit * src._shape[slice_dim], inner_unit * slice_count)
        return out

    @staticmethod
    def from_numpy(n):
        dev = _rt.current_device()
        host = torch.from_numpy(np.ascontiguousarray(n, dtype=np.float32).reshape(-1))
        t = torch.empty(host.numel(), dtype=torch.float32, device=dev.device)
        t.copy_(host, non_blocking=True)
        return NArray(t, list(reversed(n.shape)), dev)

    def to_numpy(self):
        """Blocking device->host read (NArray::Get, narray.cpp:221-224), shape reversed."""
        self._dev.stream.synchronize()
        torch.cuda.current_stream(self._dev.device).synchronize()
        return self._t.cpu().numpy().reshape(tuple(reversed(self._shape)))
