"""owl.net -- the Caffe-style layer graph of the reference (owl/owl/net/net.py) reduced to what the
hot path needs: the compute units AlexNet / GoogLeNet / LeNet / the MNIST MLP are made of, the
forward / backward / update traversals (net.py:1082-1124), programmatic builders for the
BASELINE.json configs, the Caffe front end (net_helper.py: prototxt -> units, snapshots, .caffemodel
converter; the model files live under models/) and the data layer (data.py: pinned, double-buffered,
stream-ordered uint8 feed with the transform on the device).

Every unit speaks only the owl API (NArray operators, owl.conv, owl.elewise), exactly like the
reference's units, so the op sequence per layer is the reference's.  The backend module is
injectable (`Net(backend=...)`): the parity tests run the same graph on the CPU oracle.
"""
from .net import (Net, ComputeUnit, ConvConnection, FullyConnection, ReluUnit, SigmoidUnit, TanhUnit, LRNUnit,  # noqa: F401
                  PoolingUnit, DropoutUnit, SoftmaxUnit, ConcatUnit, DataUnit, AccuracyUnit)
from .net_helper import CaffeNetBuilder, CaffeModelLoader  # noqa: F401
from .builders import build_alexnet, build_lenet, build_mnist_mlp, build_googlenet  # noqa: F401
from .trainer import NetTrainer  # noqa: F401
