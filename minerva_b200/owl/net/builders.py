"""Programmatic builders for the BASELINE.json configs (the reference builds these graphs from Caffe
prototxt files it does not vendor).  Shapes, fillers and hyper-parameters follow
SURVEY.md 8(d): bvlc_alexnet train_val without groups (README.md:100), apps/mnist_common.h for the
MNIST nets, bvlc_googlenet train_val."""
from .net import (Net, DataUnit, ConvConnection, FullyConnection, ReluUnit, LRNUnit, PoolingUnit, DropoutUnit,
                  SoftmaxUnit, ConcatUnit)


def _conv_relu(net, name, btm, num_output, k, stride=1, pad=0, std=0.01, bias=0.0, filler="gaussian"):
    net.add_unit(ConvConnection(name, btm, name, num_output, k, stride, pad, weight_std=std, bias_value=bias,
                                weight_filler=filler))
    net.add_unit(ReluUnit("relu_" + name, name, name + "_r"))
    return name + "_r"


def build_alexnet(backend=None, num_classes=1000):
    """data{227,227,3,N} -> conv 11x11/4 x96 -> relu -> LRN -> pool 3x3/2 -> conv 5x5 pad2 x256 -> relu ->
    LRN -> pool -> conv 3x3 pad1 x384 -> relu -> x384 -> relu -> x256 -> relu -> pool -> fc4096 -> relu ->
    dropout .5 -> fc4096 -> relu -> dropout .5 -> fc1000 -> softmax-loss."""
    net = Net(backend)
    net.add_unit(DataUnit("data", ["data", "label"]))
    t = _conv_relu(net, "conv1", "data", 96, 11, 4, 0, std=0.01, bias=0.0)
    net.units[1].need_bp = False     # owl.net computes conv1's data gradient (net.py:709); it is unused
    net.add_unit(LRNUnit("norm1", t, "norm1", 5, 1e-4, 0.75))
    net.add_unit(PoolingUnit("pool1", "norm1", "pool1", 3, 2))
    t = _conv_relu(net, "conv2", "pool1", 256, 5, 1, 2, std=0.01, bias=0.1)
    net.add_unit(LRNUnit("norm2", t, "norm2", 5, 1e-4, 0.75))
    net.add_unit(PoolingUnit("pool2", "norm2", "pool2", 3, 2))
    t = _conv_relu(net, "conv3", "pool2", 384, 3, 1, 1, std=0.01, bias=0.0)
    t = _conv_relu(net, "conv4", t, 384, 3, 1, 1, std=0.01, bias=0.1)
    t = _conv_relu(net, "conv5", t, 256, 3, 1, 1, std=0.01, bias=0.1)
    net.add_unit(PoolingUnit("pool5", t, "pool5", 3, 2))
    net.add_unit(FullyConnection("fc6", "pool5", "fc6", 4096, weight_std=0.005, bias_value=0.1))
    net.add_unit(ReluUnit("relu6", "fc6", "fc6_r"))
    net.add_unit(DropoutUnit("drop6", "fc6_r", "fc6_d", 0.5))
    net.add_unit(FullyConnection("fc7", "fc6_d", "fc7", 4096, weight_std=0.005, bias_value=0.1))
    net.add_unit(ReluUnit("relu7", "fc7", "fc7_r"))
    net.add_unit(DropoutUnit("drop7", "fc7_r", "fc7_d", 0.5))
    net.add_unit(FullyConnection("fc8", "fc7_d", "fc8", num_classes, weight_std=0.01, bias_value=0.0))
    net.add_unit(SoftmaxUnit("loss", "fc8", "label", "prob"))
    net.base_lr = net.current_lr = 0.01
    net.momentum, net.base_weight_decay = 0.9, 5e-4
    net.input_shape, net.num_classes = [227, 227, 3], num_classes
    return net


def build_lenet(backend=None):
    """apps/mnist_common.h:123-222 (MnistCnnAlgo): conv 5x5 1->16 -> relu -> max 2x2/2 -> conv 5x5 pad2
    16->32 -> relu -> max 3x3/3 -> fc 10 -> softmax.  N(0,0.1) weights, lr 0.01, plain SGD."""
    net = Net(backend)
    net.add_unit(DataUnit("data", ["data", "label"]))
    t = _conv_relu(net, "conv1", "data", 16, 5, 1, 0, std=0.1)
    net.units[1].need_bp = False
    net.add_unit(PoolingUnit("pool1", t, "pool1", 2, 2))
    t = _conv_relu(net, "conv2", "pool1", 32, 5, 1, 2, std=0.1)
    net.add_unit(PoolingUnit("pool2", t, "pool2", 3, 3))
    net.add_unit(FullyConnection("fc", "pool2", "fc", 10, weight_std=0.1))
    net.add_unit(SoftmaxUnit("loss", "fc", "label", "prob"))
    net.base_lr = net.current_lr = 0.01
    net.momentum, net.base_weight_decay = 0.0, 0.0
    net.input_shape, net.num_classes = [28, 28, 1], 10
    return net


def build_mnist_mlp(backend=None):
    """apps/mnist_common.h:224-288 (MnistMlpAlgo): 784 -> 256 relu -> 10 softmax."""
    net = Net(backend)
    net.add_unit(DataUnit("data", ["data", "label"]))
    net.add_unit(FullyConnection("fc1", "data", "fc1", 256, weight_std=0.1))
    net.units[1].need_bp = False
    net.add_unit(ReluUnit("relu1", "fc1", "fc1_r"))
    net.add_unit(FullyConnection("fc2", "fc1_r", "fc2", 10, weight_std=0.1))
    net.add_unit(SoftmaxUnit("loss", "fc2", "label", "prob"))
    net.base_lr = net.current_lr = 0.01
    net.momentum, net.base_weight_decay = 0.0, 0.0
    net.input_shape, net.num_classes = [784], 10
    return net


def _inception(net, name, btm, c1, c3r, c3, c5r, c5, cp):
    a = _conv_relu(net, name + "/1x1", btm, c1, 1, std=0.03, bias=0.2, filler="xavier")
    b = _conv_relu(net, name + "/3x3_reduce", btm, c3r, 1, std=0.09, bias=0.2, filler="xavier")
    b = _conv_relu(net, name + "/3x3", b, c3, 3, 1, 1, std=0.03, bias=0.2, filler="xavier")
    c = _conv_relu(net, name + "/5x5_reduce", btm, c5r, 1, std=0.2, bias=0.2, filler="xavier")
    c = _conv_relu(net, name + "/5x5", c, c5, 5, 1, 2, std=0.03, bias=0.2, filler="xavier")
    net.add_unit(PoolingUnit(name + "/pool", btm, name + "/pool", 3, 1, 1))
    d = _conv_relu(net, name + "/pool_proj", name + "/pool", cp, 1, std=0.1, bias=0.2, filler="xavier")
    net.add_unit(ConcatUnit(name + "/output", [a, b, c, d], name + "/output"))
    return name + "/output"


def _aux_head(net, name, btm, num_classes):
    net.add_unit(PoolingUnit(name + "/ave_pool", btm, name + "/ave_pool", 5, 3, 0, pool="avg"))
    t = _conv_relu(net, name + "/conv", name + "/ave_pool", 128, 1, std=0.08, bias=0.2, filler="xavier")
    net.add_unit(FullyConnection(name + "/fc", t, name + "/fc", 1024, weight_std=0.02, bias_value=0.2, weight_filler="xavier"))
    net.add_unit(ReluUnit(name + "/relu_fc", name + "/fc", name + "/fc_r"))
    net.add_unit(DropoutUnit(name + "/drop_fc", name + "/fc_r", name + "/fc_d", 0.7))
    net.add_unit(FullyConnection(name + "/classifier", name + "/fc_d", name + "/classifier", num_classes,
                                 weight_std=0.0009765625, bias_value=0.0, weight_filler="xavier"))
    net.add_unit(SoftmaxUnit(name + "/loss", name + "/classifier", "label", name + "/prob", loss_weight=0.3))


def build_googlenet(backend=None, num_classes=1000):
    """bvlc_googlenet train_val: 9 inception modules, LRN x2, 2 auxiliary heads (loss weight 0.3),
    data{224,224,3,N}."""
    net = Net(backend)
    net.add_unit(DataUnit("data", ["data", "label"]))
    t = _conv_relu(net, "conv1/7x7_s2", "data", 64, 7, 2, 3, std=0.015, bias=0.2, filler="xavier")
    net.units[1].need_bp = False
    net.add_unit(PoolingUnit("pool1/3x3_s2", t, "pool1", 3, 2))
    net.add_unit(LRNUnit("pool1/norm1", "pool1", "norm1", 5, 1e-4, 0.75))
    t = _conv_relu(net, "conv2/3x3_reduce", "norm1", 64, 1, std=0.1, bias=0.2, filler="xavier")
    t = _conv_relu(net, "conv2/3x3", t, 192, 3, 1, 1, std=0.03, bias=0.2, filler="xavier")
    net.add_unit(LRNUnit("conv2/norm2", t, "norm2", 5, 1e-4, 0.75))
    net.add_unit(PoolingUnit("pool2/3x3_s2", "norm2", "pool2", 3, 2))
    t = _inception(net, "inception_3a", "pool2", 64, 96, 128, 16, 32, 32)
    t = _inception(net, "inception_3b", t, 128, 128, 192, 32, 96, 64)
    net.add_unit(PoolingUnit("pool3/3x3_s2", t, "pool3", 3, 2))
    t = _inception(net, "inception_4a", "pool3", 192, 96, 208, 16, 48, 64)
    _aux_head(net, "loss1", t, num_classes)
    t = _inception(net, "inception_4b", t, 160, 112, 224, 24, 64, 64)
    t = _inception(net, "inception_4c", t, 128, 128, 256, 24, 64, 64)
    t = _inception(net, "inception_4d", t, 112, 144, 288, 32, 64, 64)
    _aux_head(net, "loss2", t, num_classes)
    t = _inception(net, "inception_4e", t, 256, 160, 320, 32, 128, 128)
    net.add_unit(PoolingUnit("pool4/3x3_s2", t, "pool4", 3, 2))
    t = _inception(net, "inception_5a", "pool4", 256, 160, 320, 32, 128, 128)
    t = _inception(net, "inception_5b", t, 384, 192, 384, 48, 128, 128)
    net.add_unit(PoolingUnit("pool5/7x7_s1", t, "pool5", 7, 1, 0, pool="avg"))
    net.add_unit(DropoutUnit("pool5/drop", "pool5", "pool5_d", 0.4))
    net.add_unit(FullyConnection("loss3/classifier", "pool5_d", "loss3/classifier", num_classes, weight_std=0.01, weight_filler="xavier"))
    net.add_unit(SoftmaxUnit("loss3/loss3", "loss3/classifier", "label", "prob"))
    net.base_lr = net.current_lr = 0.01
    net.momentum, net.base_weight_decay = 0.9, 2e-4
    net.input_shape, net.num_classes = [224, 224, 3], num_classes
    return net
