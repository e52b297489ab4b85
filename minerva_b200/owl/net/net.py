"""Compute units and the Net DAG (reference: owl/owl/net/net.py:24-1139)."""
import zlib

import numpy as np


def _default_backend():
    import minerva_b200.owl as owl
    import minerva_b200.owl.conv as co
    import minerva_b200.owl.elewise as ele

    class _B:
        pass
    b = _B()
    b.owl, b.co, b.ele = owl, co, ele
    return b


class ComputeUnit(object):
    """net.py:24-80: named bottoms/tops, forward(from_btm, to_top, phase), backward(from_top, to_btm, phase)."""

    def __init__(self, name, btm_names, top_names):
        self.name, self.btm_names, self.top_names = name, list(btm_names), list(top_names)
        self.out = None
        self.B = None

    def forward(self, from_btm, to_top, phase):
        pass

    def backward(self, from_top, to_btm, phase):
        pass

    def weight_update(self, base_lr, base_weight_decay, momentum, batch_size):
        pass


class ComputeUnitSimple(ComputeUnit):
    """net.py:82-110: one bottom, one top, ff/bp."""

    def forward(self, from_btm, to_top, phase):
        to_top[self.top_names[0]] = self.ff(from_btm[self.btm_names[0]], phase)
        self.out = to_top[self.top_names[0]]

    def backward(self, from_top, to_btm, phase):
        to_btm[self.btm_names[0]] = self.bp(from_top[self.top_names[0]], phase)


class WeightedComputeUnit(ComputeUnitSimple):
    """net.py:112-266: weight/bias, their grads and momentum buffers, fillers, SGD update."""

    def __init__(self, name, btm, top, lr_mult=(1.0, 2.0), decay_mult=(1.0, 0.0), weight_std=0.01, bias_value=0.0,
                 weight_filler="gaussian"):
        super().__init__(name, [btm], [top])
        self.lr_mult_w, self.lr_mult_b = lr_mult
        self.decay_mult_w, self.decay_mult_b = decay_mult
        self.weight_std, self.bias_value, self.weight_filler = weight_std, bias_value, weight_filler
        self.weight = self.bias = None
        self.weightdelta = self.biasdelta = None
        self.weightgrad = self.biasgrad = None
        self.wshape = self.bshape = None
        self.fan_in = self.fan_out = 1
        self.need_bp = True     # first layer: the data gradient is not needed (owl.net computes it anyway)

    grad_out = None      # (flat fp32 buffer for weightgrad, for biasgrad): set by the data-parallel trainer's peer merge

    def _place_grads(self):
        """The two results computed next (weight gradient, then bias gradient) are written into `grad_out`."""
        if self.grad_out is not None:
            self.B.owl.NArray.place_next(*self.grad_out)

    def init_weights_with_filler(self):
        """net.py:196-238.  Gaussian fillers are drawn on the device; xavier (uniform) on the host with
        numpy, as the reference does, then uploaded."""
        owl = self.B.owl
        if self.weight_filler == "xavier":       # net.py:222-224
            scale = float(np.sqrt(3.0 / self.fan_in))
            # seeded by a stable hash of the FULL unit name: same-length names (inception_4b/5x5_reduce vs
            # inception_4c/5x5_reduce) must not draw identical weights; identical on every data-parallel rank
            rs = np.random.RandomState((zlib.crc32(self.name.encode()) ^ 1234) & 0x7FFFFFFF)
            self.weight = owl.from_numpy(rs.uniform(-scale, scale, list(reversed(self.wshape))).astype(np.float32))
        else:
            self.weight = owl.randn(self.wshape, 0.0, self.weight_std)
        self.bias = owl.zeros(self.bshape)
        if self.bias_value:
            self.bias = self.bias + self.bias_value
        self.weightdelta = owl.zeros(self.wshape)
        self.biasdelta = owl.zeros(self.bshape)

    def weight_update(self, base_lr, base_weight_decay, momentum, batch_size):
        """net.py:240-266, the reference's ten-op chain per tensor."""
        self.weightdelta = momentum * self.weightdelta \
            - (base_lr * self.lr_mult_w / batch_size) * self.weightgrad \
            - (base_lr * self.lr_mult_w * base_weight_decay * self.decay_mult_w) * self.weight
        self.weight = self.weight + self.weightdelta
        self.weightgrad = None
        self.biasdelta = momentum * self.biasdelta \
            - (base_lr * self.lr_mult_b / batch_size) * self.biasgrad \
            - (base_lr * self.lr_mult_b * base_weight_decay * self.decay_mult_b) * self.bias
        self.bias = self.bias + self.biasdelta
        self.biasgrad = None


class ReluUnit(ComputeUnitSimple):
    def __init__(self, name, btm, top):
        super().__init__(name, [btm], [top])

    fused = False     # set by Net._plan_fusion: the producing convolution already rectified x in its epilogue
    bp_fused = False  # set by Net._plan_fusion: the consuming LRN / max-pooling unit's backward kernel applies the mask

    def ff(self, x, phase):
        self.ff_y = x if self.fused else self.B.ele.relu(x)
        return self.ff_y

    twin_for = None   # set by Net._plan_fusion: the ConvConnection whose output this unit alone reads

    def bp(self, y, phase):
        if self.bp_fused:
            return y
        geo = getattr(self.twin_for, "geo", None)
        if geo is not None and len(y.shape) == 4:
            # the result feeds that convolution's backward calls only: leave its channels-last twin with it
            return self.B.owl.NArray.relu_back_tw(y, self.ff_y, geo)
        return self.B.ele.relu_back(y, self.ff_y)


class SigmoidUnit(ComputeUnitSimple):
    def __init__(self, name, btm, top):
        super().__init__(name, [btm], [top])

    def ff(self, x, phase):
        self.ff_y = self.B.ele.sigm(x)
        return self.ff_y

    def bp(self, y, phase):
        return self.B.owl.NArray.sigm_back(y, self.ff_y, self.ff_y)


class TanhUnit(ComputeUnitSimple):
    def __init__(self, name, btm, top):
        super().__init__(name, [btm], [top])

    def ff(self, x, phase):
        self.ff_y = self.B.ele.tanh(x)
        return self.ff_y

    def bp(self, y, phase):
        return self.B.owl.NArray.tanh_back(y, self.ff_y, self.ff_y)


class PoolingUnit(ComputeUnitSimple):
    def __init__(self, name, btm, top, kernel, stride, pad=0, pool="max"):
        super().__init__(name, [btm], [top])
        self.geom = (kernel, kernel, stride, stride, pad, pad)
        self.pool = pool

    use_idx = True    # Net.fuse_pool_index: 3x3/2 max pooling remembers its arg-max bytes for the backward pass

    def ff(self, x, phase):
        if not hasattr(self, "pooler"):
            co = self.B.co
            self.pooler = co.Pooler(*self.geom, op=co.pool_op.max if self.pool == "max" else co.pool_op.avg)
        self.ff_x = x
        self.ff_idx = None
        if self.use_idx and self.pool == "max" and hasattr(self.pooler, "ff_idx") and self.pooler.idx_ok(x.shape):
            # backward reads 5 B per pooled element instead of re-reading the bottom (mnv_max_pooling_*_idx)
            self.ff_y, self.ff_idx = self.pooler.ff_idx(x)
            return self.ff_y
        self.ff_y = self.pooler.ff(x)
        return self.ff_y

    relu_bp = False   # set by Net._plan_fusion: ff_x is a ReLU output whose backward mask this unit applies

    def bp(self, y, phase):
        if self.ff_idx is not None:
            return self.pooler.bp_idx(y, self.ff_idx, self.ff_y, self.ff_x.shape, relu=self.relu_bp)
        if self.relu_bp:
            return self.pooler.bp(y, self.ff_y, self.ff_x, relu=True)
        return self.pooler.bp(y, self.ff_y, self.ff_x)


class DropoutUnit(ComputeUnitSimple):
    """net.py:355-383: mask = randb(shape, keep); y = x o mask * 1/(1-ratio)."""

    def __init__(self, name, btm, top, dropout_ratio):
        super().__init__(name, [btm], [top])
        self.scale = 1.0 / (1.0 - dropout_ratio)
        self.keep_ratio = 1 - dropout_ratio

    def ff(self, x, phase):
        if phase == "TRAIN":
            self.dropmask = self.B.owl.randb(x.shape, self.keep_ratio)
            return self.B.ele.mult(x, self.dropmask) * self.scale
        return x

    def bp(self, y, phase):
        if phase == "TRAIN":
            return self.B.ele.mult(y, self.dropmask) * self.scale
        return y


class LRNUnit(ComputeUnitSimple):
    """net.py:486-502"""

    def __init__(self, name, btm, top, local_size=5, alpha=1e-4, beta=0.75):
        super().__init__(name, [btm], [top])
        self.args = (local_size, alpha, beta)

    lite = True       # Net.fuse_lrn_recompute: drop the `scale` array, recompute it in the backward kernel

    def ff(self, x, phase):
        if not hasattr(self, "lrner"):
            self.lrner = self.B.co.Lrner(*self.args)
        self.ff_x = x
        if self.lite and hasattr(self.lrner, "ff_lite") and self.lrner.lite_ok(x.shape):
            # 8 + 12 B per element for the two passes instead of 12 + 20 (mnv_lrn_forward_lite / _backward_lite)
            self.scale = None
            self.ff_y = self.lrner.ff_lite(x)
            return self.ff_y
        # the reference allocates `scale` with owl.zeros (net.py:498); LRNForward overwrites every element, so the
        # device backend skips that fill (182 + 119 MB of writes per AlexNet step)
        self.scale = getattr(self.B.owl, "_uninit", self.B.owl.zeros)(x.shape)
        self.ff_y = self.lrner.ff(x, self.scale)
        return self.ff_y

    relu_bp = False   # set by Net._plan_fusion: ff_x is a ReLU output whose backward mask this unit applies

    def bp(self, y, phase):
        if self.scale is None:
            return self.lrner.bp_lite(self.ff_x, y, relu=self.relu_bp)
        if self.relu_bp:
            return self.lrner.bp(self.ff_x, self.ff_y, self.scale, y, relu=True)
        return self.lrner.bp(self.ff_x, self.ff_y, self.scale, y)


class ConcatUnit(ComputeUnit):
    """net.py:504-560: concat on the channel dim (Caffe concat_dim 1 == owl dim 2), slice on the way back."""

    def __init__(self, name, btms, top, concat_dim_caffe=1):
        super().__init__(name, btms, [top])
        self.concat_dim_caffe = concat_dim_caffe

    def forward(self, from_btm, to_top, phase):
        narrays = [from_btm[b] for b in self.btm_names]
        self.dim = len(narrays[0].shape) - 1 - self.concat_dim_caffe
        self.slice_count = [a.shape[self.dim] for a in narrays]
        to_top[self.top_names[0]] = self.B.owl.concat(narrays, self.dim)
        self.out = to_top[self.top_names[0]]

    def backward(self, from_top, to_btm, phase):
        top = from_top[self.top_names[0]]
        split = getattr(getattr(self.B.owl, "NArray", None), "split", None)
        if split is not None:            # every slice in one launch (mnv_copy_strided_n)
            for b, piece in zip(self.btm_names, split(top, self.dim, self.slice_count)):
                to_btm[b] = piece
            return
        st = 0
        for b, cnt in zip(self.btm_names, self.slice_count):
            to_btm[b] = self.B.owl.slice(top, self.dim, st, cnt)
            st += cnt


class FullyConnection(WeightedComputeUnit):
    """net.py:563-619: y = W * x + b with x reshaped to {features, N}."""

    def __init__(self, name, btm, top, num_output, **kw):
        super().__init__(name, btm, top, **kw)
        self.num_output = num_output

    def ff(self, act, phase):
        shp = act.shape
        a = act.reshape([int(np.prod(shp[0:-1])), shp[-1]]) if len(shp) > 2 else act
        self.ff_act, self.ff_a2d = act, a
        if self.weight is None:
            self.fan_in, self.fan_out = a.shape[0], self.num_output
            self.wshape, self.bshape = [self.num_output, a.shape[0]], [self.num_output, 1]
            self.init_weights_with_filler()
        return self.weight * a + self.bias

    def bp(self, sen, phase):
        shp = self.ff_act.shape
        self._place_grads()
        self.weightgrad = sen * self.ff_a2d.trans()
        self.biasgrad = sen.sum(1)
        if not self.need_bp:
            return None
        s = self.weight.trans() * sen
        return s.reshape(shp) if len(shp) > 2 else s


class ConvConnection(WeightedComputeUnit):
    """net.py:621-716 (group == 1 only, as the reference asserts at :690-697)."""

    def __init__(self, name, btm, top, num_output, kernel_size, stride=1, pad=0, **kw):
        super().__init__(name, btm, top, **kw)
        self.num_output, self.kernel_size, self.stride, self.pad = num_output, kernel_size, stride, pad
        self.fuse_relu = False

    def ff(self, act, phase):
        if not hasattr(self, "convolver"):
            self.convolver = self.B.co.Convolver(self.pad, self.pad, self.stride, self.stride)
        self.ff_act = act
        if self.weight is None:
            ci = act.shape[2]
            self.fan_in = self.kernel_size * self.kernel_size * ci
            self.wshape, self.bshape = [self.kernel_size, self.kernel_size, ci, self.num_output], [self.num_output]
            self.init_weights_with_filler()
        if self.geo is None and hasattr(self.convolver, "geo"):
            self.geo = self.convolver.geo(act, self.weight)
        if self.fuse_relu:
            return self.convolver.ff(act, self.weight, self.bias, relu=True)
        return self.convolver.ff(act, self.weight, self.bias)

    geo = None           # the C-ABI geometry tuple of this layer (backends with channels-last twins only)
    fuse_grads = True    # Net.fuse_conv_grads: weight and bias gradient from one call (mnv_conv_backward_filter_bias)

    def bp(self, sen, phase):
        self._place_grads()
        if self.fuse_grads and hasattr(self.convolver, "weight_bias_grad"):
            self.weightgrad, self.biasgrad = self.convolver.weight_bias_grad(sen, self.ff_act, self.weight)
        else:
            self.weightgrad = self.convolver.weight_grad(sen, self.ff_act, self.weight)
            self.biasgrad = self.convolver.bias_grad(sen)
        if not self.need_bp:
            return None
        return self.convolver.bp(sen, self.ff_act, self.weight)


class SoftmaxUnit(ComputeUnit):
    """net.py:385-440: softmax + cross-entropy; backward is (y - label) * loss_weight."""

    def __init__(self, name, btm, label, top, loss_weight=1.0):
        super().__init__(name, [btm, label], [top])
        self.loss_weight = loss_weight

    def forward(self, from_btm, to_top, phase):
        co = self.B.co
        self.ff_y = co.softmax(from_btm[self.btm_names[0]], co.soft_op.instance)
        self.y = from_btm[self.btm_names[1]]      # one-hot {classes, N}
        to_top[self.top_names[0]] = self.ff_y
        self.out = self.ff_y

    def backward(self, from_top, to_btm, phase):
        d = self.ff_y - self.y
        to_btm[self.btm_names[0]] = d * self.loss_weight if self.loss_weight != 1.0 else d

    def getloss(self):
        lossmat = self.B.ele.mult(self.B.ele.ln(self.ff_y), self.y)
        res = lossmat.sum(0).sum(1).to_numpy()
        return -float(res.reshape(-1)[0]) / lossmat.shape[1]

    def getloss_device(self):
        """The same reduction left on the device: -> (1-element NArray holding sum(ln(y) o label), batch size).  The
        caller reads it back when it chooses to (a training loop that logs the loss one step late never drains the
        launch queue; the reference's trainer prints losses from asynchronously evaluated NArrays the same way,
        owl/owl/net/trainer.py:139-146)."""
        lossmat = self.B.ele.mult(self.B.ele.ln(self.ff_y), self.y)
        return lossmat.sum(0).sum(1), lossmat.shape[1]


class AccuracyUnit(ComputeUnit):
    """net.py:436-487: top-1 accuracy of the scores against one-hot labels.  Non-lazy like the reference's (count_zero
    blocks), so it only runs in the TEST phase or when asked (`getacc`)."""

    def __init__(self, name, btm, label, top, top_k=1):
        super().__init__(name, [btm, label], [top])
        self.top_k, self.acc, self.batch_size = top_k, 0.0, 0

    def forward(self, from_btm, to_top, phase):
        self._scores, self._labels = from_btm[self.btm_names[0]], from_btm[self.btm_names[1]]
        if phase == "TEST":
            self.getacc()

    def getacc(self):
        predict = self._scores.max_index(0)
        truth = self._labels.max_index(0)
        self.batch_size = self._scores.shape[-1]
        self.acc = (predict - truth).count_zero() * 1.0 / self.batch_size
        return self.acc


class DataUnit(ComputeUnit):
    """Synthetic data layer: the caller sets `.data` / `.label` (owl NArrays) before forward."""

    def __init__(self, name, tops):
        super().__init__(name, [], tops)
        self.data = self.label = None

    def forward(self, from_btm, to_top, phase):
        to_top[self.top_names[0]] = self.data
        if len(self.top_names) > 1:
            to_top[self.top_names[1]] = self.label


class Net(object):
    """The unit DAG (net.py:880-1139): topological forward, reverse backward with multi-consumer
    sensitivities summed by `+` (net.py:1102-1114), per-unit update."""

    def __init__(self, backend=None):
        self.B = backend or _default_backend()
        self.units = []
        self.name_to_uid = {}
        self.base_lr = self.current_lr = 0.01
        self.momentum = 0.9
        self.base_weight_decay = 5e-4
        self.batch_size = 0          # GLOBAL batch: the update divisor (net.py:252-254,1121-1124)
        self.on_weight_grad = None   # hook(unit) fired as soon as a unit's gradients exist
        self.fuse_conv_relu = True   # False: run conv and ReLU as the reference's two ops
        self.fuse_relu_backward = True   # False: ReLU backward stays its own pass in front of LRN / max-pooling backward
        self.fuse_lrn_recompute = True   # False: LRN keeps the reference's (bottom, top, scale) three-array form
        self.fuse_pool_index = True      # False: max pooling backward recomputes the arg-max from the bottom
        self.fuse_conv_grads = True      # False: ConvBackwardFilter and ConvBackwardBias stay two ops
        self.fuse_conv_twins = True      # False: every convolution call makes its own channels-last workspace copy

    def add_unit(self, unit):
        unit.B = self.B
        self.name_to_uid[unit.name] = len(self.units)
        self.units.append(unit)
        return unit

    def get_weighted_unit_ids(self):
        return [i for i, u in enumerate(self.units) if isinstance(u, WeightedComputeUnit)]

    def get_loss_units(self):
        return [u for u in self.units if isinstance(u, SoftmaxUnit)]

    def get_data_unit(self):
        return [u for u in self.units if isinstance(u, DataUnit)][0]

    def _plan_fusion(self):
        """SURVEY 8f fusion: a ConvConnection whose output is read by exactly one unit, a ReluUnit, rectifies in its
        own epilogue (mnv_conv_forward_relu) and the ReluUnit passes the blob through; results are bit-identical to
        the two-op sequence (owl/owl/net/net.py:281-296 after :621-716).  Only backends that advertise the fused
        entry point take part (the CPU twin used by the parity tests does not)."""
        self._fusion_planned = True
        if hasattr(self.B.owl, "NArray") and hasattr(self.B.owl.NArray, "use_twins"):
            self.B.owl.NArray.use_twins = bool(self.fuse_conv_twins)
            self.B.owl.NArray._twin_wanted = {}
        for u in self.units:
            if isinstance(u, LRNUnit):
                u.lite = bool(self.fuse_lrn_recompute)
            if isinstance(u, PoolingUnit):
                u.use_idx = bool(self.fuse_pool_index)
            if isinstance(u, ConvConnection):
                u.fuse_grads = bool(self.fuse_conv_grads)

        def readers_of(i, top):
            out = []
            for v in self.units[i + 1:]:
                if top in v.btm_names:
                    out.append(v)
                if top in v.top_names:      # an in-place unit rewrites the name: later readers see its output
                    break
            return out
        if getattr(self.B.co, "FUSED_CONV_RELU", False) and self.fuse_conv_relu:
            for i, u in enumerate(self.units):
                if isinstance(u, ConvConnection):
                    readers = readers_of(i, u.top_names[0])
                    if len(readers) == 1 and isinstance(readers[0], ReluUnit):
                        u.fuse_relu, readers[0].fused = True, True
        # a ReluUnit that alone reads a convolution's output produces, in backward, exactly that convolution's top_diff:
        # its kernel leaves the channels-last twin with the result (mnv_relu_backward_tw)
        if self.fuse_conv_twins and hasattr(getattr(self.B.owl, "NArray", None), "relu_back_tw"):
            for i, u in enumerate(self.units):
                if isinstance(u, ConvConnection):
                    readers = readers_of(i, u.top_names[0])
                    if len(readers) == 1 and isinstance(readers[0], ReluUnit):
                        readers[0].twin_for = u
        # backward: a ReluUnit read only by an LRN or max-pooling unit hands its mask to that unit's backward kernel
        # (mnv_lrn_backward_relu / mnv_max_pooling_backward_relu; both read the ReLU output anyway)
        if getattr(self.B.co, "FUSED_RELU_BACKWARD", False) and self.fuse_relu_backward:
            for i, u in enumerate(self.units):
                if isinstance(u, ReluUnit):
                    readers = readers_of(i, u.top_names[0])
                    if len(readers) == 1 and (isinstance(readers[0], LRNUnit) or
                                              (isinstance(readers[0], PoolingUnit) and readers[0].pool == "max")):
                        u.bp_fused, readers[0].relu_bp = True, True

    def forward(self, phase="TRAIN"):
        if not getattr(self, "_fusion_planned", False):
            self._plan_fusion()
        blobs = {}
        for u in self.units:      # builders append units in topological order
            u.forward(blobs, blobs, phase)
        self._blobs = blobs

    def backward(self, phase="TRAIN"):
        # `sens[name]` is the sensitivity of the CURRENT version of blob `name`: a unit takes (pops) its tops'
        # sensitivities before it writes its bottoms', so an in-place unit (top name == bottom name, the Caffe
        # convention for relu / dropout) replaces the entry instead of adding to it, and only contributions of
        # different consumers of one blob version are summed -- what the reference's per-unit dicts do
        # (owl/owl/net/net.py:1102-1114).
        sens = {}        # blob name -> list of contributions, summed left to right when the producer takes them
        add_n = getattr(getattr(self.B.owl, "NArray", None), "add_n", None)

        def take(name):
            parts = sens.pop(name, None)
            if not parts:
                return None
            if len(parts) == 1:
                return parts[0]
            if add_n is not None:          # one pass, the bits of the chained `+` (mnv_add_n)
                return add_n(parts)
            total = parts[0]
            for v in parts[1:]:
                total = total + v
            return total
        for u in reversed(self.units):
            if isinstance(u, DataUnit):
                continue
            top_sens = {t: take(t) for t in u.top_names}
            if not isinstance(u, SoftmaxUnit) and any(v is None for v in top_sens.values()):
                continue
            out = {}
            u.backward(top_sens, out, phase)
            for k, v in out.items():
                if v is None:
                    continue
                sens.setdefault(k, []).append(v)
            if self.on_weight_grad is not None and isinstance(u, WeightedComputeUnit):
                self.on_weight_grad(u)

    def update(self, uid):
        self.units[uid].weight_update(self.current_lr, self.base_weight_decay, self.momentum, self.batch_size)

    def weight_update(self):
        for uid in self.get_weighted_unit_ids():
            self.update(uid)
