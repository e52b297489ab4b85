"""owl.conv -- convolution, pooling, LRN and softmax wrappers (reference: owl/owl/conv.py:13-182)."""
from .narray import NArray, ConvInfo, PoolingInfo, pooling_algo, softmax_algo

soft_op = softmax_algo
pool_op = pooling_algo
FUSED_CONV_RELU = True     # Convolver.ff(..., relu=True) exists: owl.net folds single-consumer ReLU units into it
FUSED_RELU_BACKWARD = True  # Lrner.bp / max Pooler.bp(..., relu=True) exist: owl.net folds the preceding ReLU unit's backward into them


def softmax(x, op=soft_op.instance):
    """owl/owl/conv.py:13-33: non-4D inputs are padded to {d0,1,1,N} first."""
    if len(x.shape) == 4:
        return NArray.softmax_forward(x, op)
    ori_shape = list(x.shape)
    soft_shape = x.shape[0:-1] + [1 for _ in range(4 - len(ori_shape))] + [x.shape[-1]]
    return NArray.softmax_forward(x.reshape(soft_shape), op).reshape(ori_shape)


class Lrner:
    def __init__(self, local_size, alpha, beta):
        self.local_size, self.alpha, self.beta = local_size, alpha, beta

    def ff(self, x, scale):
        return NArray.lrn_forward(x, scale, self.local_size, self.alpha, self.beta)

    def bp(self, bottom_data, top_data, scale, top_diff, relu=False):
        return NArray.lrn_backward(bottom_data, top_data, scale, top_diff, self.local_size, self.alpha, self.beta, relu)

    # scale-less pair (not in the reference API): forward stores only its output, backward recomputes scale / top
    def lite_ok(self, shape):
        return NArray.lrn_lite_ok(shape, self.local_size)

    def ff_lite(self, x):
        return NArray.lrn_forward_lite(x, self.local_size, self.alpha, self.beta)

    def bp_lite(self, bottom_data, top_diff, relu=False):
        return NArray.lrn_backward_lite(bottom_data, top_diff, self.local_size, self.alpha, self.beta, relu)


class Convolver:
    def __init__(self, pad_h, pad_w, stride_v, stride_h):
        self.param = ConvInfo(pad_h, pad_w, stride_v, stride_h)

    def ff(self, x, w, b, relu=False):
        """relu=True (not in the reference signature): rectify in the convolution's epilogue (mnv_conv_forward_relu)."""
        return NArray.conv_forward(x, w, b, self.param, relu)

    def bp(self, y, x, w):
        return NArray.conv_backward_data(y, x, w, self.param)

    def geo(self, x, w):
        """The (N, Ci, Co, H, W, pads, strides, fh, fw) tuple the mnv_conv_* entries take for this layer (not in the reference API)."""
        return NArray.conv_geo(x, w, self.param)

    def weight_grad(self, y, x, w):
        return NArray.conv_backward_filter(y, x, w, self.param)

    def weight_bias_grad(self, y, x, w):
        """(weight_grad, bias_grad) from one call (not in the reference API): the bias sums ride on a pre-pass of weight_grad."""
        return NArray.conv_backward_filter_bias(y, x, w, self.param)

    def bias_grad(self, y):
        return NArray.conv_backward_bias(y)


class Pooler:
    def __init__(self, h, w, stride_v, stride_h, pad_h=0, pad_w=0, op=pool_op.max):
        self.param = PoolingInfo(op, h, w, stride_v, stride_h, pad_h, pad_w)

    def ff(self, x):
        return NArray.pooling_forward(x, self.param)

    def bp(self, y, ff_y, ff_x, relu=False):
        return NArray.pooling_backward(y, ff_y, ff_x, self.param, relu)

    # arg-max remembering pair for 3x3/2 max pooling (not in the reference API): see mnv_max_pooling_*_idx
    def idx_ok(self, shape):
        return NArray.pooling_idx_ok(self.param, shape)

    def ff_idx(self, x):
        return NArray.pooling_forward_idx(x, self.param)

    def bp_idx(self, y, idx, ff_y, bottom_shape, relu=False):
        return NArray.pooling_backward_idx(y, idx, bottom_shape, self.param, ff_y if relu else None)
