"""Caffe front end and snapshots (reference: owl/owl/net/net_helper.py:11-317).

  CaffeNetBuilder(solver_file)        solver + net prototxt -> units of an owl.net.Net (build_net), the raw per-tensor
                                      fp32 snapshot format (save_net_to_file / init_net_from_file, net_helper.py:156-270)
  CaffeModelLoader(model, dir, idx)   .caffemodel -> such a snapshot: filters rotated by 180 degrees (Minerva convolves,
                                      Caffe correlates), inner-product weights transposed (net_helper.py:272-317)

Differences from the reference, all forced by the environment: the prototxt files are parsed by
owl/net/prototxt.py instead of the generated caffe_pb2 module; Data layers become FeedDataUnits over a synthetic provider
of the declared geometry (no LMDB library in the image); in-place layers (top == bottom) keep their blob names -- Net
tracks sensitivities per blob version -- instead of being renamed after the layer.
"""
import os

import numpy as np

from . import prototxt
from . import net as N
from .caffemodel import read_caffemodel

# V1LayerParameter.LayerType names (old "layers { type: CONVOLUTION }" files) -> new-style type strings
_V1 = {"CONVOLUTION": "Convolution", "INNER_PRODUCT": "InnerProduct", "POOLING": "Pooling", "RELU": "ReLU", "LRN": "LRN",
       "DROPOUT": "Dropout", "SOFTMAX_LOSS": "SoftmaxWithLoss", "CONCAT": "Concat", "ACCURACY": "Accuracy", "DATA": "Data",
       "SIGMOID": "Sigmoid", "TANH": "TanH", "IMAGE_DATA": "ImageData", "WINDOW_DATA": "WindowData"}


def _filler(p, default_std=0.01):
    f = p.sub("weight_filler")
    kind = str(f.get("type", "constant"))
    return ("xavier", default_std) if kind == "xavier" else ("gaussian", float(f.get("std", default_std if kind == "gaussian" else 0.0)))


def _mults(layer):
    """(lr_mult_w, lr_mult_b), (decay_mult_w, decay_mult_b): `param { lr_mult decay_mult }` blocks, or V1 blobs_lr /
    weight_decay lists; defaults as net.py:183-197 of the reference."""
    ps = layer.all("param")
    if ps:
        lw, dw = float(ps[0].get("lr_mult", 1.0)), float(ps[0].get("decay_mult", 1.0))
        lb, db = (float(ps[1].get("lr_mult", 1.0)), float(ps[1].get("decay_mult", 0.0))) if len(ps) > 1 else (1.0, 0.0)
        return (lw, lb), (dw, db)
    lr, wd = layer.all("blobs_lr"), layer.all("weight_decay")
    return ((float(lr[0]) if lr else 1.0, float(lr[1]) if len(lr) > 1 else 1.0),
            (float(wd[0]) if wd else 1.0, float(wd[1]) if len(wd) > 1 else 0.0))


class CaffeNetBuilder(object):
    def __init__(self, solver_file, net_file=None):
        self.solver_file = solver_file
        with open(solver_file) as f:
            self.solverconfig = prototxt.parse(f.read())
        self.net_file = net_file or self.solverconfig.get("net")
        if not os.path.isabs(self.net_file) and not os.path.exists(self.net_file):
            self.net_file = os.path.join(os.path.dirname(os.path.abspath(solver_file)), os.path.basename(self.net_file))
        self.change_net(self.net_file)
        self.snapshot_dir = self.solverconfig.get("snapshot_prefix", "")

    def change_net(self, net_file):
        """net_helper.py:30-38: use another network file than the one the solver names."""
        self.net_file = net_file
        with open(net_file) as f:
            self.netconfig = prototxt.parse(f.read())

    # ------------------------------------------------------------------------------------------------------------
    def build_net(self, owl_net, num_gpu=1, phase="TRAIN", stored_shape=None, feed=True):
        """Fill `owl_net` with the units of the network file (layers restricted to `phase` by their include rules).
        num_gpu: the TRAIN batch of the data layer is divided by it (netio.py:282).  stored_shape (C, H, W): geometry of
        the stored images the data layer's synthetic provider draws (default: 3 x 256 x 256, or the crop size when the
        layer does not crop).  feed=False: a plain DataUnit (the caller sets .data / .label)."""
        s = self.solverconfig
        owl_net.base_lr = owl_net.current_lr = float(s.get("base_lr", 0.01))
        owl_net.base_weight_decay = float(s.get("weight_decay", 0.0))
        owl_net.momentum = float(s.get("momentum", 0.0))
        owl_net.solver, owl_net.lr_policy = s, str(s.get("lr_policy", "fixed"))
        owl_net.data_layers, owl_net.loss_uids, owl_net.accuracy_uids = [], [], []
        for l in self.netconfig.all("layer") + self.netconfig.all("layers"):
            inc = [str(r.get("phase")) for r in l.all("include") if r.has("phase")]
            if inc and phase not in inc:
                continue
            exc = [str(r.get("phase")) for r in l.all("exclude") if r.has("phase")]
            if phase in exc:
                continue
            unit = self._convert_type(l, num_gpu, owl_net, stored_shape, feed)
            if unit is None:
                continue
            owl_net.add_unit(unit)
            uid = len(owl_net.units) - 1
            if isinstance(unit, N.DataUnit):
                owl_net.data_layers.append(unit.name)
            elif isinstance(unit, N.SoftmaxUnit):
                owl_net.loss_uids.append(uid)
            elif isinstance(unit, N.AccuracyUnit):
                owl_net.accuracy_uids.append(uid)
        for u in owl_net.units:           # the first weighted layer reads the data blob: its data gradient is never used
            if isinstance(u, N.WeightedComputeUnit):
                if u.btm_names[0] in [d.top_names[0] for d in owl_net.units if isinstance(d, N.DataUnit)]:
                    u.need_bp = False
        return owl_net

    def _convert_type(self, l, num_gpu, owl_net, stored_shape, feed):
        ty = str(l.get("type"))
        ty = _V1.get(ty, ty)
        name, btm, top = str(l.get("name")), [str(b) for b in l.all("bottom")], [str(t) for t in l.all("top")]
        if ty in ("Data", "ImageData", "WindowData"):
            return self._data_unit(l, name, top, num_gpu, owl_net, stored_shape, feed)
        if ty == "Convolution":
            p = l.sub("convolution_param")
            assert int(p.get("group", 1)) == 1, "group convolution is not supported (the reference asserts the same, net.py:690-697)"
            filler, std = _filler(p)
            lr, dm = _mults(l)
            return N.ConvConnection(name, btm[0], top[0], int(p.get("num_output")), int(p.get("kernel_size")),
                                    int(p.get("stride", 1)), int(p.get("pad", 0)), lr_mult=lr, decay_mult=dm, weight_std=std,
                                    bias_value=float(p.sub("bias_filler").get("value", 0.0)), weight_filler=filler)
        if ty == "InnerProduct":
            p = l.sub("inner_product_param")
            filler, std = _filler(p)
            lr, dm = _mults(l)
            return N.FullyConnection(name, btm[0], top[0], int(p.get("num_output")), lr_mult=lr, decay_mult=dm, weight_std=std,
                                     bias_value=float(p.sub("bias_filler").get("value", 0.0)), weight_filler=filler)
        if ty == "Pooling":
            p = l.sub("pooling_param")
            pool = str(p.get("pool", "MAX"))
            return N.PoolingUnit(name, btm[0], top[0], int(p.get("kernel_size")), int(p.get("stride", 1)), int(p.get("pad", 0)),
                                 pool="max" if pool == "MAX" else "avg")
        if ty == "ReLU":
            return N.ReluUnit(name, btm[0], top[0])
        if ty == "Sigmoid":
            return N.SigmoidUnit(name, btm[0], top[0])
        if ty == "TanH":
            return N.TanhUnit(name, btm[0], top[0])
        if ty == "Dropout":
            return N.DropoutUnit(name, btm[0], top[0], float(l.sub("dropout_param").get("dropout_ratio", 0.5)))
        if ty == "LRN":
            p = l.sub("lrn_param")
            return N.LRNUnit(name, btm[0], top[0], int(p.get("local_size", 5)), float(p.get("alpha", 1.0)), float(p.get("beta", 0.75)))
        if ty == "Concat":
            p = l.sub("concat_param")
            return N.ConcatUnit(name, btm, top[0], int(p.get("axis", p.get("concat_dim", 1))))
        if ty == "SoftmaxWithLoss":
            lw = l.all("loss_weight")
            return N.SoftmaxUnit(name, btm[0], btm[1], top[0] if top else name, float(lw[0]) if lw else 1.0)
        if ty == "Accuracy":
            return N.AccuracyUnit(name, btm[0], btm[1], top[0] if top else name, int(l.sub("accuracy_param").get("top_k", 1)))
        print("Not implemented type:", ty)      # net_helper.py:153
        return None

    def _data_unit(self, l, name, top, num_gpu, owl_net, stored_shape, feed):
        from .data import FeedDataUnit, HostFeed, SyntheticImageProvider
        dp = l.sub("data_param") if l.has("data_param") else l.sub("image_data_param") if l.has("image_data_param") else l.sub("window_data_param")
        tp = l.sub("transform_param")
        batch = int(dp.get("batch_size", 1)) // max(int(num_gpu), 1)            # netio.py:282
        owl_net.batch_size = int(dp.get("batch_size", 1))                      # GLOBAL batch = the update divisor
        crop = int(tp.get("crop_size", 0))
        C, H, W = stored_shape or ((3, 256, 256) if crop else (3, 224, 224))
        ch, cw = (crop, crop) if crop else (H, W)
        owl_net.input_shape = [cw, ch, C]
        unit = FeedDataUnit(name, top) if feed else N.DataUnit(name, top)
        unit.geometry = dict(batch=batch, stored=(C, H, W), crop=(ch, cw), mirror=bool(tp.get("mirror", False)),
                             scale=float(tp.get("scale", 1.0)), mean=[float(v) for v in tp.all("mean_value")] or None,
                             mean_file=tp.get("mean_file"), source=dp.get("source"))
        if feed:
            def attach(classes, owl=None, rt=None, seed=0, geo=unit.geometry, unit=unit):
                """Bind the synthetic provider + HostFeed once the device exists (the builder itself runs without a GPU)."""
                import minerva_b200.owl as _owl
                from minerva_b200.owl import _runtime as _rt
                prov = SyntheticImageProvider(geo["batch"], geo["stored"][0], geo["stored"][1:], classes, seed=seed)
                unit.feed = HostFeed(owl or _owl, rt or _rt, provider=prov, mean=geo["mean"], scale=geo["scale"],
                                     crop=geo["crop"], mirror=geo["mirror"], seed=seed)
                return unit.feed
            unit.attach = attach
        return unit

    # ---- snapshots: <dir>/snapshot<idx>/<layer>_{weights,weightdelta,bias,biasdelta}.dat, raw fp32 -----------------
    @staticmethod
    def _path(weightpath, snapshotidx, unit, what):
        return os.path.join(weightpath, "snapshot%d" % snapshotidx, "%s_%s.dat" % (unit.name.replace("/", "_"), what))

    def save_net_to_file(self, owl_net, weightpath, snapshotidx):
        """net_helper.py:236-270."""
        os.makedirs(os.path.join(weightpath, "snapshot%d" % snapshotidx), exist_ok=True)
        for u in owl_net.units:
            if isinstance(u, N.WeightedComputeUnit) and u.weight is not None:
                for what, arr in (("weights", u.weight), ("weightdelta", u.weightdelta), ("bias", u.bias), ("biasdelta", u.biasdelta)):
                    arr.to_numpy().astype(np.float32).reshape(-1).tofile(self._path(weightpath, snapshotidx, u, what))

    def init_net_from_file(self, owl_net, weightpath, snapshotidx):
        """net_helper.py:156-234: load what exists and fits; a missing file or a size mismatch leaves the filler-initialised
        tensor in place ("Weight Need Reinit").  The units must know their shapes (one forward pass, or set wshape/bshape)."""
        owl = owl_net.B.owl
        reinit = []
        for u in owl_net.units:
            if not isinstance(u, N.WeightedComputeUnit):
                continue
            if u.wshape is None:
                raise RuntimeError("init_net_from_file: run one forward pass first so that %s knows its shape" % u.name)
            for what, attr, shape in (("weights", "weight", u.wshape), ("weightdelta", "weightdelta", u.wshape),
                                      ("bias", "bias", u.bshape), ("biasdelta", "biasdelta", u.bshape)):
                path = self._path(weightpath, snapshotidx, u, what)
                if not os.path.isfile(path):
                    if what in ("weights", "bias"):
                        reinit.append((u.name, what))
                    continue
                a = np.fromfile(path, dtype=np.float32)
                if a.size != int(np.prod(shape)):
                    reinit.append((u.name, what))
                    continue
                setattr(u, attr, owl.from_numpy(a.reshape(list(reversed(shape)))))
        for name, what in reinit:
            print("%s Need Reinit %s" % ("Weight" if what == "weights" else "Bias", name))
        return reinit


class CaffeModelLoader(object):
    """net_helper.py:272-317 / scripts/modelconvertor/caffe2minerva.py: write a snapshot from a .caffemodel.
    Convolution filters [Co, Ci, kh, kw] are rotated by 180 degrees per (co, ci) plane; inner-product weights
    [num_output, input_dim] are transposed (owl's {num_output, input_dim} matrix is column-major)."""

    def __init__(self, model_file, weightdir, snapshot):
        out = os.path.join(weightdir, "snapshot%d" % snapshot)
        os.makedirs(out, exist_ok=True)
        self.converted = []
        for l in read_caffemodel(model_file):
            if len(l["blobs"]) != 2:
                continue
            w, b = l["blobs"]
            layername = l["name"].replace("/", "_")
            if l["type"] == "Convolution":
                shape = [d for d in w["shape"]]
                co, ci, kh, kw = ([1] * (4 - len(shape)) + shape)[-4:]
                filt = w["data"].reshape(co, ci, kh, kw)[:, :, ::-1, ::-1]
                np.ascontiguousarray(filt).reshape(-1).tofile(os.path.join(out, layername + "_weights.dat"))
            else:
                num_output = b["data"].size
                mat = w["data"].reshape(num_output, -1)
                np.ascontiguousarray(mat.T).reshape(-1).tofile(os.path.join(out, layername + "_weights.dat"))
            b["data"].tofile(os.path.join(out, layername + "_bias.dat"))
            self.converted.append(l["name"])


def net_to_prototxt(net, name, batch, crop, mirror=True, mean_values=(104.0, 117.0, 123.0), scale=1.0, inplace=True):
    """Write an owl.net.Net (as the programmatic builders make it) as a Caffe train_val NetParameter -- how the model files
    under models/ were produced (the reference reads BVLC's files, which it does not vendor).  inplace=True uses
    Caffe's convention of ReLU / Dropout layers whose top is their bottom."""
    M = prototxt.Msg
    root = M().add("name", name)
    alias = {}                       # blob name in `net` -> blob name in the file (in-place layers collapse names)

    def nm(b):
        return alias.get(b, b)
    for u in net.units:
        l = M().add("name", u.name)
        if isinstance(u, N.DataUnit):
            l.add("type", "Data")
            for t in u.top_names:
                l.add("top", t)
            l.add("include", M().add("phase", prototxt.Enum("TRAIN")))
            tp = M().add("mirror", bool(mirror))
            if crop:
                tp.add("crop_size", int(crop))
            for v in (mean_values or ()):
                tp.add("mean_value", float(v))
            if scale != 1.0:
                tp.add("scale", float(scale))
            l.add("transform_param", tp)
            l.add("data_param", M().add("source", "synthetic").add("batch_size", int(batch)).add("backend", prototxt.Enum("LMDB")))
            root.add("layer", l)
            continue
        for b in u.btm_names:
            l.add("bottom", nm(b))
        if inplace and isinstance(u, (N.ReluUnit, N.DropoutUnit)):
            alias[u.top_names[0]] = nm(u.btm_names[0])
        for t in u.top_names:
            l.add("top", nm(t))
        if isinstance(u, N.WeightedComputeUnit):
            l.order.insert(1, ("type", "Convolution" if isinstance(u, N.ConvConnection) else "InnerProduct"))
            l.fields["type"] = [l.order[1][1]]
            l.add("param", M().add("lr_mult", float(u.lr_mult_w)).add("decay_mult", float(u.decay_mult_w)))
            l.add("param", M().add("lr_mult", float(u.lr_mult_b)).add("decay_mult", float(u.decay_mult_b)))
            p = M().add("num_output", int(u.num_output))
            if isinstance(u, N.ConvConnection):
                if u.pad:
                    p.add("pad", int(u.pad))
                p.add("kernel_size", int(u.kernel_size))
                if u.stride != 1:
                    p.add("stride", int(u.stride))
            wf = M().add("type", u.weight_filler)
            if u.weight_filler == "gaussian":
                wf.add("std", float(u.weight_std))
            p.add("weight_filler", wf).add("bias_filler", M().add("type", "constant").add("value", float(u.bias_value)))
            l.add("convolution_param" if isinstance(u, N.ConvConnection) else "inner_product_param", p)
        elif isinstance(u, N.ReluUnit):
            l.order.insert(1, ("type", "ReLU")); l.fields["type"] = ["ReLU"]
        elif isinstance(u, N.SigmoidUnit):
            l.order.insert(1, ("type", "Sigmoid")); l.fields["type"] = ["Sigmoid"]
        elif isinstance(u, N.TanhUnit):
            l.order.insert(1, ("type", "TanH")); l.fields["type"] = ["TanH"]
        elif isinstance(u, N.DropoutUnit):
            l.order.insert(1, ("type", "Dropout")); l.fields["type"] = ["Dropout"]
            l.add("dropout_param", M().add("dropout_ratio", float(1.0 - u.keep_ratio)))
        elif isinstance(u, N.LRNUnit):
            l.order.insert(1, ("type", "LRN")); l.fields["type"] = ["LRN"]
            l.add("lrn_param", M().add("local_size", int(u.args[0])).add("alpha", float(u.args[1])).add("beta", float(u.args[2])))
        elif isinstance(u, N.PoolingUnit):
            l.order.insert(1, ("type", "Pooling")); l.fields["type"] = ["Pooling"]
            k, _, s, _, pd, _ = u.geom
            p = M().add("pool", prototxt.Enum("MAX" if u.pool == "max" else "AVE")).add("kernel_size", int(k)).add("stride", int(s))
            if pd:
                p.add("pad", int(pd))
            l.add("pooling_param", p)
        elif isinstance(u, N.ConcatUnit):
            l.order.insert(1, ("type", "Concat")); l.fields["type"] = ["Concat"]
        elif isinstance(u, N.SoftmaxUnit):
            l.order.insert(1, ("type", "SoftmaxWithLoss")); l.fields["type"] = ["SoftmaxWithLoss"]
            if u.loss_weight != 1.0:
                l.add("loss_weight", float(u.loss_weight))
        elif isinstance(u, N.AccuracyUnit):
            l.order.insert(1, ("type", "Accuracy")); l.fields["type"] = ["Accuracy"]
        root.add("layer", l)
    return prototxt.dump(root)
