#!/usr/bin/env python
"""bench.py -- AlexNet (bvlc_alexnet without groups, batch 256 per GPU) training images/s on N B200s,
the metric BASELINE.json names, measured through the owl API over the sm_100a kernel library.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's CPU path on the host cores)

One JSON line on stdout (rank 0).  `value` times K training steps (forward, backward, gradient merge over
NVLink overlapped with backward, momentum-SGD update) with inputs resident in HBM, CUDA events on the
compute stream, max over ranks.  `e2e` is the same step fed from pinned host memory every step (the uint8
image batch the reference's data layer reads, converted on the device; the fp32 feed is reported beside it)
with the loss read back every step.  `roofline` describes the dominant kernel (the tcgen05 GEMM /
implicit-GEMM conv kernel) from per-call CUDA events; `op_table` carries every C-ABI call's algorithmic
bytes / flops, achieved GB/s / TFLOP/s and roofline fraction (BASELINE metric: "per-op HBM GB/s and
tensor-pipe %"); `other_configs` are the remaining BASELINE.json configs measured after the headline;
`cpu_baseline` / `cpu_reference_ops` time the CPU oracle / the reference's own basic:: functions on the
host; at N>1 `merge_check` proves the data-parallel gradient merge before anything is timed.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


WORKLOADS = {
    "alexnet": dict(builder="build_alexnet", batch=256, classes=1000, uniform=False,
                    name="owl AlexNet (bvlc_alexnet train_val, no groups), batch 256/GPU, random-init weights"),
    "lenet": dict(builder="build_lenet", batch=256, classes=10, uniform=True, name="apps/mnist_cnn LeNet-style CNN, batch 256"),
    "mlp": dict(builder="build_mnist_mlp", batch=256, classes=10, uniform=True, name="apps/mnist_mlp 784-256-10 MLP, batch 256"),
    "googlenet": dict(builder="build_googlenet", batch=120, classes=1000, uniform=False,
                      name="owl GoogLeNet (bvlc_googlenet train_val), batch 120/GPU, random-init weights"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="alexnet", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--unfused-update", action="store_true", help="use the reference's ten-op SGD chain")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the GoogLeNet / LeNet / MLP lines after the headline")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from a CUDA graph (NetTrainer(graph=True)); auto = on at 1 GPU")
    ap.add_argument("--no-pdl", action="store_true", help="A/B: plain stream-ordered launches (mnv_set_dependent_launch(0))")
    ap.add_argument("--merge", default="auto", choices=["auto", "peer", "nccl", "off"], help="N>1 gradient merge (owl/net/merge.py)")
    ap.add_argument("--nccl-ctas", type=int, default=0,
                    help="N>1: SMs left to NCCL (NCCL_MAX_CTAS) and kept out of the persistent tensor-core kernel's grid; 0 = do not manage")
    ap.add_argument("--mnv-opt", action="append", default=[], metavar="KEY=INT",
                    help="tuning: switch to the tuning build of the library and set a mnv_debug_set_option key (recorded in config.tuning)")
    return ap.parse_args()


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the tensor-core kernel, averaged over the launches of one
    ncu launch-list capture of this command (profiles/r0N_launch_list_summary.json, tools/launch_summary.py); null for
    workloads that have no committed capture."""
    if workload != "alexnet":
        return None
    for name in ("r02_launch_list_summary.json", "r01_launch_list_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)["umma_gemm_kernel"]["dram_bytes_per_launch"])
        except Exception:
            continue
    return None


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        p.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
        p["source"] = "measured"
    except Exception:
        pass
    return p


# ------------------------------------------------------------------------------------------------
# synthetic data of the named shape
# ------------------------------------------------------------------------------------------------
def host_batch(wl, shape, batch, seed):
    import numpy as np
    rs = np.random.RandomState(seed)
    dims = [batch] + list(reversed(shape))
    x = (rs.uniform(0, 1, dims) if wl.get("uniform") else rs.standard_normal(dims)).astype(np.float32)
    lab = rs.randint(0, wl["classes"], batch)
    onehot = np.zeros((batch, wl["classes"]), np.float32)
    onehot[np.arange(batch), lab] = 1
    return x, onehot


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML)
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.03)      # a few samples per timed region: every NVML query takes driver locks the launching thread also wants

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# algorithmic work of one C-ABI call: ("tensor", flops) or ("hbm", bytes)
# ------------------------------------------------------------------------------------------------
def _prod(v):
    p = 1
    for x in v:
        p *= int(x)
    return p


def _pooled(x, pad, win, stride):
    p = (x + 2 * pad - win + stride - 1) // stride + 1
    return p - 1 if (p - 1) * stride >= x + pad else p


def op_work(name, a):
    if name in ("mnv_matmult", "mnv_matmult_ex"):
        return "tensor", 2.0 * a[3] * a[4] * a[5]
    if name in ("mnv_conv_forward", "mnv_conv_forward_relu", "mnv_conv_backward_data", "mnv_conv_backward_filter",
                "mnv_conv_backward_filter_bias", "mnv_conv_forward_tw", "mnv_conv_backward_data_tw", "mnv_conv_backward_filter_tw"):
        off = 4 if name.startswith("mnv_conv_forward") or name.endswith("_bias") or name == "mnv_conv_backward_filter_tw" else 3
        N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw = a[off:off + 11]
        Ho, Wo = (H + 2 * ph - fh) // sv + 1, (W + 2 * pw - fw) // sh + 1
        return "tensor", 2.0 * N * Ho * Wo * Co * Ci * fh * fw
    if name in ("mnv_add", "mnv_sub", "mnv_dot_mult", "mnv_dot_div"):
        return "hbm", 12.0 * a[3]
    if name == "mnv_accumulate":
        return "hbm", 12.0 * a[2]
    if name == "mnv_copy_strided_n":          # prof_args = (count, elements)
        return "hbm", 8.0 * a[1]
    if name == "mnv_add_n":                   # prof_args = (count, n)
        return "hbm", 4.0 * (a[0] + 1) * a[1]
    if name == "mnv_relu_backward_tw":        # top, top_diff, bottom_diff, N, C, H, W, ...: 8 B read + 8 B written (NCHW result + its channels-last twin)
        return "hbm", 16.0 * _prod(a[3:7])
    if name == "mnv_scale":
        return "hbm", 8.0 * a[2]
    if name in ("mnv_const_add", "mnv_left_const_sub", "mnv_left_const_div", "mnv_const_div"):
        return "hbm", 8.0 * a[3]
    if name in ("mnv_elewise_exp", "mnv_elewise_ln", "mnv_elewise_negative", "mnv_copy"):
        return "hbm", 8.0 * a[2]
    if name == "mnv_reshape":
        return "hbm", 2.0 * a[2]
    if name in ("mnv_sigmoid_forward", "mnv_relu_forward", "mnv_tanh_forward"):
        return "hbm", 8.0 * _prod(a[2:6])
    if name in ("mnv_sigmoid_backward", "mnv_relu_backward", "mnv_tanh_backward"):
        return "hbm", 12.0 * _prod(a[4:8])
    if name.startswith("mnv_norm_"):
        return "hbm", 8.0 * a[3] * a[4] + 4.0 * max(a[3], a[4])
    if name.startswith("mnv_reduction_") or name.startswith("mnv_max_index_"):
        m, n = a[2], a[3]
        return "hbm", 4.0 * (m * n + (n if name.endswith("_col") else m))
    if name == "mnv_transpose":
        return "hbm", 8.0 * a[2] * a[3]
    if name == "mnv_copy_strided":
        return "hbm", 8.0 * a[2] * a[3]
    if name == "mnv_conv_backward_bias":
        return "hbm", 4.0 * (_prod(a[2:6]) + a[3])
    if name.endswith("_softmax_forward"):
        return "hbm", 8.0 * _prod(a[2:6])
    if name.endswith("_softmax_backward"):
        return "hbm", 12.0 * _prod(a[3:7])
    if name in ("mnv_max_pooling_forward", "mnv_average_pooling_forward", "mnv_max_pooling_forward_idx"):
        off = 3 if name.endswith("_idx") else 2
        N, C, H, W, sv, sh, wh, ww, ph, pw = a[off:off + 10]
        ein, eout = N * C * H * W, N * C * _pooled(H, ph, wh, sv) * _pooled(W, pw, ww, sh)
        return "hbm", 4.0 * ein + (5.0 if name.endswith("_idx") else 4.0) * eout
    if name in ("mnv_max_pooling_backward", "mnv_max_pooling_backward_relu", "mnv_average_pooling_backward", "mnv_max_pooling_backward_idx"):
        N, C, H, W, sv, sh, wh, ww, ph, pw = a[4:14]
        ein, eout = N * C * H * W, N * C * _pooled(H, ph, wh, sv) * _pooled(W, pw, ww, sh)
        if name.endswith("_idx"):
            return "hbm", 5.0 * eout + 4.0 * ein + (4.0 * eout if a[2] else 0.0)
        if name.startswith("mnv_average"):
            return "hbm", 4.0 * (ein + eout)
        return "hbm", 4.0 * (2 * ein + 2 * eout)
    if name == "mnv_lrn_forward":
        return "hbm", 12.0 * _prod(a[6:10])
    if name in ("mnv_lrn_backward", "mnv_lrn_backward_relu"):
        return "hbm", 20.0 * _prod(a[8:12])
    if name == "mnv_lrn_forward_lite":
        return "hbm", 8.0 * _prod(a[5:9])
    if name == "mnv_lrn_backward_lite":
        return "hbm", 12.0 * _prod(a[6:10])
    if name in ("mnv_fill", "mnv_randn", "mnv_rand_bernoulli"):
        return "hbm", 4.0 * a[1]
    if name == "mnv_sgd_momentum_update":
        return "hbm", 20.0 * a[3]
    if name == "mnv_sgd_momentum_update_multi":
        return "hbm", 20.0 * a[3]          # the owl binding reports (.., .., .., total parameters) to the profiler
    if name == "mnv_image_transform_u8":
        return "hbm", 5.0 * a[4] * a[5] * a[8] * a[9]        # 1 B in + 4 B out per element (the mean image stays in L2)
    return None, 0.0


def op_table_from(table, steps, pk):
    """table: [(name, args, ms)] over `steps` instrumented steps -> ({name: row}, totals of the tensor / HBM-bound calls)."""
    per = {}
    for name, a, ms in table:
        kind, work = op_work(name, a)
        d = per.setdefault(name, {"n": 0, "ms": 0.0, "work": 0.0, "kind": kind})
        d["n"] += 1
        d["ms"] += ms
        d["work"] += work
    total_ms = sum(v["ms"] for v in per.values())
    tf32_burst, hbm = pk["bf16_tflops"] / 2.0, pk["hbm_gbs"]
    out = {}
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"calls_per_step": v["n"] / steps, "ms_per_step": v["ms"] / steps, "share": v["ms"] / total_ms if total_ms else 0.0}
        sec = v["ms"] * 1e-3
        if v["kind"] == "tensor" and sec > 0:
            row.update(bound="tensor", flops_per_step=v["work"] / steps, tflops=v["work"] / sec / 1e12,
                       frac=v["work"] / sec / 1e12 / tf32_burst)
        elif v["kind"] == "hbm" and sec > 0:
            row.update(bound="hbm", bytes_algorithmic_per_step=v["work"] / steps, gbs=v["work"] / sec / 1e9,
                       frac=v["work"] / sec / 1e9 / hbm)
        out[k] = row
    g = [v for v in per.values() if v["kind"] == "tensor"]
    g_flops, g_ms, g_n = sum(v["work"] for v in g), sum(v["ms"] for v in g), sum(v["n"] for v in g)
    h = [v for v in per.values() if v["kind"] == "hbm"]
    h_bytes, h_ms = sum(v["work"] for v in h), sum(v["ms"] for v in h)
    return out, (g_flops, g_ms, g_n, total_ms, h_bytes, h_ms)


# ------------------------------------------------------------------------------------------------
# CPU legs: the same owl.net graph on the oracle backend (test infrastructure used as a baseline only)
# ------------------------------------------------------------------------------------------------
def cpu_step_rate(wl, sample_batch, steps, warmup, use_ref=True):
    from oracle import owl_cpu
    import minerva_b200.owl.net as onet
    used_ref = owl_cpu.use_reference(use_ref)
    B = owl_cpu.Backend()
    owl_cpu.set_seed(1)
    net = getattr(onet, wl["builder"])(B)
    x, onehot = host_batch(wl, net.input_shape, sample_batch, 0)
    du = net.get_data_unit()
    du.data, du.label = B.owl.from_numpy(x), B.owl.from_numpy(onehot)
    net.batch_size = sample_batch
    tr = onet.NetTrainer(net, None, fused_update=False)
    for _ in range(warmup):
        tr.step()
    t0 = time.time()
    for _ in range(steps):
        tr.step()
    dt = time.time() - t0
    return sample_batch * steps / dt, dt, used_ref


def cpu_reference_ops(budget_s=20.0):
    """the reference's OWN CPU functions (minerva/op/impl/basic.cpp compiled into oracle/_ref)
    timed one op at a time, single-threaded -- that is how the reference's CpuDevice runs an op -- at the shapes the
    AlexNet step issues (MatMult is the naive triple loop of basic.cpp:164-171: the fc8 shape only, fc6 would take minutes).
    -> {op: {shape, ms, GB/s | GFLOP/s}} or {"unavailable": why}."""
    import numpy as np
    from oracle import pyoracle as orc
    if not orc.have_ref():
        return {"unavailable": "oracle/_ref/libminerva_ref.so not built (needs /root/reference at build time)"}
    R = orc.Ref
    rs = np.random.RandomState(5)
    out = {"how": "reference basic:: functions (oracle/_ref), one thread per op as on the reference's CpuDevice, best of 2"}
    t_start = time.time()

    def timed(fn):
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best

    def add(name, shape, fn, bytes_=None, flops=None):
        if time.time() - t_start > budget_s:
            return
        dt = timed(fn)
        row = {"shape": shape, "ms": dt * 1e3}
        if bytes_ is not None:
            row["gbs"] = bytes_ / dt / 1e9
        if flops is not None:
            row["gflops"] = flops / dt / 1e9
        out[name] = row

    E = 290400 * 256 // 4            # a quarter of the conv1 activations (18.6 M): bounded sample, same access pattern
    a, b = rs.standard_normal(E).astype(np.float32), rs.standard_normal(E).astype(np.float32)
    add("Arithmetic add", "18.6 M (conv1 activations / 4)", lambda: R.arithmetic("add", a, b), bytes_=12.0 * E)
    add("ArithmeticConst mult", "18.6 M", lambda: R.arithmetic_const("mult", 0, 0.5, a), bytes_=8.0 * E)
    add("ReluForward", "18.6 M", lambda: R.activation("relu", a), bytes_=8.0 * E)
    s = a[:4096 * 256 * 4]
    add("SigmoidForward", "4.2 M", lambda: R.activation("sigmoid", s), bytes_=8.0 * s.size)
    add("TanhForward", "4.2 M", lambda: R.activation("tanh", s), bytes_=8.0 * s.size)
    add("Elewise exp", "4.2 M", lambda: R.elewise("exp", s), bytes_=8.0 * s.size)
    m = a[:4096 * 256]
    v = b[:4096]
    add("NormArithmetic add (fc bias)", "{4096,256} + {4096,1}", lambda: R.norm_arithmetic("add", 1, m, v, 4096, 256), bytes_=8.0 * m.size)
    add("Reduction sum (fc bias grad)", "{4096,256} -> {4096,1}", lambda: R.reduction("sum", 1, m, 4096, 256), bytes_=4.0 * (m.size + 4096))
    x = a[:1000 * 256]
    add("MaxIndex", "{1000,256} -> {1,256}", lambda: R.max_index(0, x, 1000, 256), bytes_=4.0 * (x.size + 256))
    add("SoftmaxForward", "{1000,1,1,256}", lambda: R.softmax_forward(x, 1000, 1, 1, 256), bytes_=8.0 * x.size)
    t = a[:4096 * 4096]
    add("Transpose (fc7 weight)", "{4096,4096}", lambda: R.transpose(t, 4096, 4096), bytes_=8.0 * t.size)
    w8, a8 = a[:1000 * 4096], b[:4096 * 256]
    add("MatMult (fc8 forward)", "1000 x 256 x 4096", lambda: R.matmult(w8, a8, 1000, 256, 4096), flops=2.0 * 1000 * 256 * 4096)
    w1, a1 = a[:256 * 784], b[:784 * 256]
    add("MatMult (mnist_mlp fc1)", "256 x 256 x 784", lambda: R.matmult(w1, a1, 256, 256, 784), flops=2.0 * 256 * 256 * 784)
    return out


def run_reference(args):
    """The reference's own CPU implementation of the path.  Minerva has NO CPU convolution / pooling /
    LRN / backward ops (minerva/op/impl/bundle.h:29-44,51-54), so the AlexNet step runs on the oracle
    restatement, with the reference's compiled basic:: functions (oracle/_ref) for the ops it does
    implement (arithmetic, MatMult, transpose, reduction, relu forward)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    if args.workload == "mlp":
        # configs[0] is the one config the reference can run on its own CPU stack: NArray -> DagScheduler ->
        # CpuDevice -> basic::, compiled from /root/reference into oracle/_ref, ReluBackward supplied as a user ComputeFn
        from oracle import pyoracle as orc
        r = orc.run_reference_mlp(wl["batch"], max(args.steps, 1), max(args.warmup, 0))
        if r is not None:
            emit({"impl": "reference", "metric": "mlp train images/s", "value": r["images_per_s"], "unit": "images/s", "n_gpus": 0,
                  "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                  "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": wl["name"], "app": r["app"]},
                  "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": 4, "kind": "reference",
                                   "sample": "%d steps of batch %d through the reference's own stack (4 CpuDevice worker threads, one "
                                             "thread per op); ReluBackward has no CPU implementation in the reference (bundle.h:32) and is "
                                             "supplied as a user ComputeFn" % (r["steps"], r["mb"])},
                  "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            return
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    rate1, dt1, used_ref = cpu_step_rate(wl, 1, 1, 0)           # calibration: one image
    budget = 150.0
    per_img = 1.0 / rate1
    sample = int(max(1, min(8, budget / max(1e-9, per_img * (args.steps + args.warmup)))))
    steps = args.steps
    while sample * per_img * (steps + args.warmup) > budget and steps > 1:
        steps -= 1
    rate, dt, _ = cpu_step_rate(wl, sample, steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "AlexNet train images/s" if args.workload == "alexnet" else args.workload + " train images/s",
        "value": rate, "unit": "images/s", "n_gpus": 0, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "sample_batch": sample},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d-image minibatch x %d steps of the same training step on the CPU oracle "
                                   "(OpenMP over independent outputs); reference basic:: functions used for the ops "
                                   "Minerva implements on CPU: %s" % (sample, steps, used_ref)},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def merge_check(C, net, trainer):
    """N>1, before anything is timed: (i) the product gradient merge (peer exchange or per-unit NCCL all-reduce) gives,
    on every rank, the NCCL all-reduce of the ranks' local gradients to <= 1e-5 relative per tensor; (ii) after a full
    step every rank holds bit-identical weights (MIN == MAX of an exact integer checksum).  Same inputs, weights and
    dropout seeds in both passes, so the local gradients are the same bits."""
    torch, dist, owl = C.torch, C.dist, C.owl
    wids = net.get_weighted_unit_ids()
    hook = net.on_weight_grad

    def one_pass(with_merge):
        owl.set_seed(4242)
        net.on_weight_grad = hook if with_merge else None
        if trainer.peer is not None:
            trainer.peer.begin_step()
        net.forward("TRAIN")
        net.backward("TRAIN")
        if with_merge:
            trainer._wait_merge()
        return [g for uid in wids for g in (net.units[uid].weightgrad, net.units[uid].biasgrad)]

    ref = []
    for g in one_pass(False):
        t = g.as_torch().clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ref.append(t)
    merged = one_pass(True)
    worst = torch.zeros(1, dtype=torch.float64, device="cuda")
    for g, r in zip(merged, ref):
        err = (g.as_torch().double() - r.double()).abs().max() / r.double().abs().max().clamp_min(1e-30)
        worst = torch.maximum(worst, err.reshape(1))
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    net.on_weight_grad = hook
    owl.set_seed(1234 + 77)
    trainer.step()                                   # a full step incl. the update
    sums = torch.stack([net.units[uid].weight.as_torch().view(torch.int32).to(torch.int64).sum() for uid in wids] +
                       [net.units[uid].bias.as_torch().view(torch.int32).to(torch.int64).sum() for uid in wids])
    lo, hi = sums.clone(), sums.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    identical = bool(torch.equal(lo, hi))
    res = {"grad_max_rel_err_vs_nccl_allreduce": float(worst.item()), "tolerance": 1e-5, "tensors": len(ref),
           "weights_bit_identical_across_ranks": identical, "ranks": C.world, "merge": trainer.merge_kind.split(",")[0]}
    res["ok"] = bool(res["grad_max_rel_err_vs_nccl_allreduce"] <= 1e-5 and identical)
    return res


def measure(C, args, wl_key, headline):
    """Build the net of one BASELINE config, warm up, time `steps` training steps with resident inputs (device-timed, max
    over ranks), then the end-to-end loop.  -> result dict (+ net / trainer for the headline's extra sections)."""
    torch, owl, onet, rt, lib = C.torch, C.owl, C.onet, C.rt, C.lib
    wl = WORKLOADS[wl_key]
    world, rank = C.world, C.rank
    owl.set_seed(1234)                      # identical initial weights on every rank
    batch = (args.batch if headline and args.batch else 0) or wl["batch"]
    net = getattr(onet, wl["builder"])()
    net.batch_size = batch * world          # the update divisor is the global batch
    x, onehot = host_batch(wl, net.input_shape, batch, 100 + rank)
    du = net.get_data_unit()
    du.data, du.label = owl.from_numpy(x), owl.from_numpy(onehot)
    use_graph = world == 1 and not args.unfused_update and args.graph != "off"
    trainer = onet.NetTrainer(net, C.dist if world > 1 else None, fused_update=not args.unfused_update, merge=args.merge, graph=use_graph)
    gdev = rt.current_device()
    res = {"workload": wl["name"], "per_gpu_batch": batch, "global_batch": batch * world, "cuda_graph": use_graph}

    def timed_steps(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0, r0 = lib.mnv_launch_count(), trainer.graph_replays
        e0.record(gdev.stream)
        t_host = time.perf_counter()
        for _ in range(n):
            trainer.step()
        e1.record(gdev.stream)
        t_host = time.perf_counter() - t_host      # host time to ENQUEUE the steps (no sync inside)
        C.sync_all()
        launches = lib.mnv_launch_count() - l0 + (trainer.graph_replays - r0) * trainer.graph_launches_per_step
        return C.max_over_ranks(e0.elapsed_time(e1)), t_host, int(launches)

    try:
        for _ in range(max(3, args.warmup)):
            trainer.step()
        C.sync_all()
    except Exception as ex:
        if not use_graph:
            raise
        # the recording is an optimisation of the launch path, not the product: report why it failed and time the
        # call-by-call step (same kernels, same results)
        print("bench.py: CUDA-graph recording failed (%r); timing the call-by-call step" % (ex,), file=sys.stderr)
        res["cuda_graph"] = "failed: %r" % (ex,)
        use_graph = False
        trainer.graph = False
        torch.cuda.synchronize()
        for _ in range(max(3, args.warmup)):
            trainer.step()
        C.sync_all()
    if world > 1:
        res["merge_check"] = merge_check(C, net, trainer)
        C.sync_all()
        if not res["merge_check"]["ok"]:
            return res, net, trainer, du, (x, onehot)
        for _ in range(3):               # back to the steady state of the training loop (the check re-seeds, syncs and reads back)
            trainer.step()
        C.sync_all()
    sampler = ClockSampler(C.local) if headline else None
    if sampler:
        sampler.start()
    ms_total, t_host, launches = timed_steps(args.steps)
    if sampler:
        sampler.stop_flag = True
        res["clocks"] = sampler.summary()
    if use_graph:
        # the same steps launched one kernel at a time from Python (what every N > 1 run does): reported beside the headline
        res["launches_per_graph_replay"] = trainer.graph_launches_per_step
        trainer.graph = False
        for _ in range(3):
            trainer.step()
        C.sync_all()
        ms_e, t_e, _ = timed_steps(args.steps)
        res["eager"] = {"value": batch * world * args.steps / (ms_e * 1e-3), "ms_per_step": ms_e / args.steps,
                        "host_enqueue_ms_per_step": t_e * 1e3 / args.steps}
        trainer.graph = True
    res.update(value=batch * world * args.steps / (ms_total * 1e-3), ms_per_step=ms_total / args.steps,
               gpu_launches=int(launches), host_enqueue_ms_per_step=t_host * 1e3 / args.steps, loss=float(net.get_loss_units()[-1].getloss()),
               gradient_merge=trainer.merge_kind)
    if hasattr(trainer, "merge_note"):
        res["gradient_merge_note"] = trainer.merge_note
    if not args.no_e2e:
        res["e2e"] = run_e2e(C, args, net, trainer, du, x, onehot, batch)
    return res, net, trainer, du, (x, onehot)


def run_e2e(C, args, net, trainer, du, x, onehot, batch):
    """The same step fed from pinned host memory EVERY step through the data unit (owl/net/data.py: double-buffered,
    copy stream, stream-ordered hand-over), loss read back every step (one step late, like the reference's asynchronous
    trainer prints).  Two feeds: "u8" -- the uint8 image batch a real data layer reads from disk (owl/owl/net/netio.py:
    259-339 converts and mean-subtracts on the HOST; here the bytes are uploaded and converted by one kernel on the device)
    -- and "f32" -- the already converted fp32 batch (4x the bytes)."""
    import numpy as np
    torch, owl, rt = C.torch, C.owl, C.rt
    from minerva_b200.owl.net.data import HostFeed
    gdev = rt.current_device()
    steps, world = args.steps, C.world
    host_loss = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    lu = net.get_loss_units()[-1]

    def loop(feed, nsteps, pipelined):
        loss_val, pending = None, None
        feed.start()
        for it in range(nsteps):
            cur = it % 2
            du.data, du.label = feed.next()              # waits (on the stream, not the host) for this step's upload
            trainer.step()
            feed.done()                                  # the buffers may be overwritten once this step has read them
            if not pipelined:
                loss_val = lu.getloss()
                continue
            dsum, n = trainer.loss_device if trainer.graph else lu.getloss_device()     # graph: reduced inside the recording
            host_loss[cur].copy_(dsum.as_torch(), non_blocking=True)
            loss_ready[cur].record(gdev.stream)
            if pending is not None:
                loss_ready[pending[0]].synchronize()
                loss_val = -float(host_loss[pending[0]][0]) / pending[1]
            pending = (cur, n)
        if pending is not None:
            loss_ready[pending[0]].synchronize()
            loss_val = -float(host_loss[pending[0]][0]) / pending[1]
        feed.stop()
        return loss_val

    def timed(feed, pipelined=True):
        loop(feed, 4, pipelined)       # both device slots of the feed are seen twice: a graph-mode trainer has bound a recording to each
        C.sync_all()
        t0 = time.perf_counter()
        loss = loop(feed, steps, pipelined)
        C.sync_all()
        return C.max_over_ranks(time.perf_counter() - t0), loss

    # uint8 feed: stored images as the reference's LMDB holds them (256 x 256 for the ImageNet nets, cropped to the net's
    # input and mirrored per image, netio.py:303-311; 28 x 28 for MNIST), synthetic bytes, scaled to about unit variance
    rs = np.random.RandomState(7 + C.rank)
    if x.ndim == 4 and x.shape[1] == 3:
        stored = rs.randint(0, 256, (batch, 3, 256, 256), dtype=np.uint8)
        feed_u8 = HostFeed(owl, rt, data_u8=stored, mean=np.full((3, 256, 256), 127.5, np.float32), scale=1.0 / 73.9,
                           crop=(x.shape[2], x.shape[3]), mirror=True, label=onehot, seed=C.rank)
    else:
        stored = rs.randint(0, 256, x.shape, dtype=np.uint8)
        feed_u8 = HostFeed(owl, rt, data_u8=stored, scale=1.0 / 255.0, label=onehot)
    dt, loss = timed(feed_u8)
    out = {"value": batch * world * steps / dt, "unit": "images/s",
           "feed": "stored uint8 images %s -> device-side mean-subtract / random crop / mirror / scale to fp32 (mnv_image_transform_u8)" % (list(stored.shape[1:]),),
           "h2d_bytes_per_step": feed_u8.bytes_per_step(), "d2h_bytes_per_step": 4, "last_loss": float(loss),
           "loss_read": "every step's loss is reduced on the device, copied to pinned host memory and read one step later, "
                        "inside the timed region"}
    feed_f32 = HostFeed(owl, rt, data_f32=x, label=onehot)
    dt32, loss32 = timed(feed_f32)
    out["f32_feed"] = {"value": batch * world * steps / dt32, "h2d_bytes_per_step": int(x.size * 4 + onehot.size * 4), "last_loss": float(loss32)}
    dtb, _ = timed(feed_f32, pipelined=False)
    out["blocking_read_value"] = batch * world * steps / dtb
    out["blocking_read_note"] = "fp32 feed with a blocking loss read after every step (drains the launch queue)"
    du.data, du.label = feed_f32.resident()
    return out


def main():
    args = parse()
    # a wedged collective must not burn the GPU lease: hard exit after 15 minutes
    watchdog = threading.Timer(900.0, lambda: os._exit(3))
    watchdog.daemon = True
    watchdog.start()
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner there) are
    # pointed at stderr for the duration of the run
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    C = Ctx()
    C.torch, C.dist = torch, dist
    C.rank = rank = int(os.environ.get("RANK", "0"))
    C.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    C.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if args.nccl_ctas > 0:
            # Experiment (off by default): the conv/GEMM kernel is persistent, one CTA per SM, so an overlapped all-reduce that
            # holds a few SMs pushes its last CTAs into a second wave.  Giving NCCL a fixed handful of CTAs and sizing the
            # tensor-core grid to the rest measured the same or worse at 2 GPUs (7.54 ms unmanaged; 2 CTAs 9.60, 4 CTAs 7.73,
            # 8 CTAs 7.54): fewer NCCL CTAs expose the 151 MB fc6 all-reduce.
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.nccl_ctas))
            os.environ.setdefault("NCCL_MIN_CTAS", "1")
            args.mnv_opt.append("sm_budget=%d" % (148 - int(os.environ["NCCL_MAX_CTAS"])))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import minerva_b200.owl as owl
    import minerva_b200.owl.net as onet
    from minerva_b200.owl import _runtime as rt
    from minerva_b200 import _lib
    # --mnv-opt (experiments only) switches the whole process to the tuning build of the library; the default run
    # uses the product library, in which every option is a compile-time constant
    lib = _lib.use_tuning() if args.mnv_opt else _lib.load()
    for kv in args.mnv_opt:
        k, v = kv.split("=")
        lib.mnv_debug_set_option(k.encode(), int(v))
    if args.no_pdl:
        lib.mnv_set_dependent_launch(0)
    C.owl, C.onet, C.rt, C.lib = owl, onet, rt, lib
    owl.set_device(owl.create_gpu_device(local))

    def sync_all():
        owl.wait_for_all()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    C.sync_all, C.max_over_ranks = sync_all, max_over_ranks

    wl = WORKLOADS[args.workload]
    head, net, trainer, du, _ = measure(C, args, args.workload, True)
    if world > 1 and not head["merge_check"]["ok"]:
        if rank == 0:
            emit({"metric": "AlexNet train images/s", "value": None, "unit": "images/s", "n_gpus": world, "error": "merge_check failed",
                  "merge_check": head["merge_check"]})
        dist.destroy_process_group()
        sys.exit(1)

    # ---- per-call device times (every rank runs the two instrumented steps: they contain the gradient merge) ----------
    roofline, optable = None, None
    pk = peaks()
    was_graph, trainer.graph = trainer.graph, False      # per-call events need the calls: two eager steps
    if rank == 0:
        rt.profiler = rt.EventProfiler()
    for _ in range(2):
        trainer.step()
    sync_all()
    trainer.graph = was_graph
    if rank == 0:
        table = rt.profiler.table()
        rt.profiler = None
        optable, (g_flops, g_ms, g_n, total_ms, h_bytes, h_ms) = op_table_from(table, 2, pk)
        burst, sustained = pk["bf16_tflops"] / 2.0, pk["bf16_tflops_sustained"] / 2.0
        achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms else 0.0
        clocks = head.get("clocks") or {}
        capped = "sw_power_cap" in (clocks.get("reasons") or [])
        roofline = {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                    "frac_burst": achieved / burst, "frac_sustained": achieved / sustained,
                    "traffic": ncu_traffic(args.workload),
                    "kernel": "mnv::umma_gemm_kernel + shift-GEMM conv kernels (tcgen05 TF32: MatMult + conv fwd/bwd-data/bwd-filter, "
                              "every pre/post pass of those calls included)",
                    "launches_timed": g_n, "avg_launch_ms": g_ms / max(g_n, 1),
                    "share_of_step": g_ms / total_ms if total_ms else None,
                    "peak_source": "%s bf16_tflops/2 = burst dense TF32 (the timed region is ~0.1 s at %s MHz, power cap seen: %s); "
                                   "frac_sustained divides by bf16_tflops_sustained/2" % (pk["source"], clocks.get("sm_mhz"), capped),
                    "hbm_bound_calls": {"achieved_gbs": h_bytes / (h_ms * 1e-3) / 1e9 if h_ms else None, "peak_gbs": pk["hbm_gbs"],
                                        "frac": h_bytes / (h_ms * 1e-3) / 1e9 / pk["hbm_gbs"] if h_ms else None,
                                        "share_of_step": h_ms / total_ms if total_ms else None}}
    if world > 1:
        dist.barrier()

    # ---- the other BASELINE configs (same warm-up / step rule, measured after the headline) ---------------------------
    others = {}
    if not args.no_other_configs and args.workload == "alexnet" and not args.batch:
        del net, trainer, du
        torch.cuda.empty_cache()
        for key in ("googlenet", "lenet", "mlp"):
            try:
                r, n2, t2, d2, _ = measure(C, args, key, False)
                del n2, t2, d2
                torch.cuda.empty_cache()
                others["%s_b%d" % (key, r["per_gpu_batch"])] = r
            except Exception as ex:     # reported, never load-bearing for the headline
                others[key] = {"error": repr(ex)}
                if world > 1:
                    break                # a rank-local failure would desynchronise the collectives of the next config

    # ---- CPU baselines on bounded samples (rank 0, N == 1) --------------------------------------------------------------
    cpu_baseline, cpu_ops = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            os.environ.setdefault("OMP_NUM_THREADS", str(cores))
            rate, dt, used_ref = cpu_step_rate(wl, 2, 1, 0)
            cpu_baseline = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                            "sample": "one training step on a 2-image minibatch (%.1f s) on the CPU oracle; Minerva has "
                                      "no CPU conv/pool/LRN/backward (bundle.h:29-44), reference basic:: used where it "
                                      "exists: %s" % (dt, used_ref)}
        except Exception as ex:   # the baseline is reported, never load-bearing
            cpu_baseline = {"value": None, "unit": "images/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
        try:
            cpu_ops = cpu_reference_ops()
        except Exception as ex:
            cpu_ops = {"unavailable": repr(ex)}
        try:       # configs[0] end to end through the reference's own CPU stack, next to our mlp_b256 line
            from oracle import pyoracle as orc
            r = orc.run_reference_mlp(256, 10, 2)
            if r is not None:
                cpu_ops["config0_mnist_mlp_reference_stack"] = {
                    "images_per_s": r["images_per_s"], "ms_per_step": r["ms_per_step"], "kind": "reference", "cores": 4,
                    "what": "apps/mnist_mlp b256 through NArray -> DagScheduler -> CpuDevice (4 worker threads) -> basic::, compiled from "
                            "/root/reference (oracle/_ref); ReluBackward (no CPU impl in the reference) as a user ComputeFn"}
        except Exception as ex:
            cpu_ops["config0_mnist_mlp_reference_stack"] = {"unavailable": repr(ex)}

    if rank == 0:
        metric = "AlexNet train images/s" if args.workload == "alexnet" else args.workload + " train images/s"
        line = {
            "metric": metric, "value": head["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 tensor-core conv/GEMM, f32 elsewhere",
            "data": "synthetic",
            "config": {"workload": wl["name"], "global_batch": head["global_batch"], "per_gpu_batch": head["per_gpu_batch"],
                       "parallelism": "dp%d" % world, "update": "chain" if args.unfused_update else "fused momentum-SGD kernel",
                       "l2": "working set per step (~2 GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "gradient_merge": head["gradient_merge"],
                       "launch": ("the whole step (forward, backward, loss reduction, update: %d kernel launches) recorded once into a CUDA "
                                  "graph and replayed; `eager` = the same steps launched call by call" % head.get("launches_per_graph_replay", 0))
                                 if head.get("cuda_graph") is True else "every C-ABI call launched from Python",
                       **({"gradient_merge_note": head["gradient_merge_note"]} if "gradient_merge_note" in head else {}),
                       **({"tuning": args.mnv_opt} if args.mnv_opt else {})},
            "clocks": head.get("clocks"), "gpu_launches": head["gpu_launches"], "host_enqueue_ms_per_step": head["host_enqueue_ms_per_step"], "loss": head["loss"],
            "cuda_graph": head.get("cuda_graph", False), "eager": head.get("eager"),   # cuda_graph: True, False, or "failed: ..." (then timed call by call)
            "e2e": head.get("e2e"), "roofline": roofline, "cpu_baseline": cpu_baseline, "op_table": optable,
            "other_configs": others, "cpu_reference_ops": cpu_ops,
            "peaks": {k: pk.get(k) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained", "source")},
        }
        if world > 1:
            line["merge_check"] = head["merge_check"]
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
