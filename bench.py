#!/usr/bin/env python
"""bench.py -- AlexNet (bvlc_alexnet without groups, batch 256 per GPU) training images/s on N B200s,
the metric BASELINE.json names, measured through the owl API over the sm_100a kernel library.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's CPU path on the host cores)

One JSON line on stdout (rank 0).  `value` times K training steps (forward, backward, NCCL gradient
all-reduce overlapped with backward, momentum-SGD update) with inputs resident in HBM, CUDA events
on the compute stream, max over ranks.  `e2e` is the same step fed from pinned host memory (H2D of
the batch every step, prefetched on a copy stream) with the loss read back every step.  `roofline`
describes the dominant kernel (the tcgen05 GEMM/implicit-GEMM conv kernel) from per-launch CUDA
events; `cpu_baseline` is the CPU oracle restatement of the same step on a bounded sample.
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
    "alexnet": dict(builder="build_alexnet", batch=256, classes=1000,
                    name="owl AlexNet (bvlc_alexnet train_val, no groups), batch 256/GPU, random-init weights"),
    "lenet": dict(builder="build_lenet", batch=256, classes=10, name="apps/mnist_cnn LeNet-style CNN, batch 256"),
    "mlp": dict(builder="build_mnist_mlp", batch=256, classes=10, name="apps/mnist_mlp 784-256-10 MLP, batch 256"),
    "googlenet": dict(builder="build_googlenet", batch=120, classes=1000,
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
    ap.add_argument("--merge", default="auto", choices=["auto", "peer", "nccl", "off"], help="N>1 gradient merge (owl/net/merge.py)")
    ap.add_argument("--nccl-ctas", type=int, default=0,
                    help="N>1: SMs left to NCCL (NCCL_MAX_CTAS) and kept out of the persistent tensor-core kernel's grid; 0 = do not manage")
    ap.add_argument("--mnv-opt", action="append", default=[], metavar="KEY=INT",
                    help="tuning: set a mnv_debug_set_option key before the run (recorded in config.tuning)")
    return ap.parse_args()


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the tensor-core kernel, averaged over the launches of one
    ncu launch-list capture of this command (profiles/r01_launch_list_summary.json, tools/launch_summary.py); null for
    workloads that have no committed capture."""
    if workload != "alexnet":
        return None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_launch_list_summary.json")) as f:
            return float(json.load(f)["umma_gemm_kernel"]["dram_bytes_per_launch"])
    except Exception:
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
    x = rs.standard_normal([batch] + list(reversed(shape))).astype(np.float32)
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
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# algorithmic work of one C-ABI call for the roofline of the dominant kernel
# ------------------------------------------------------------------------------------------------
def gemm_flops(name, a):
    if name in ("mnv_matmult", "mnv_matmult_ex"):
        return 2.0 * a[3] * a[4] * a[5]
    if name in ("mnv_conv_forward", "mnv_conv_forward_relu", "mnv_conv_backward_data", "mnv_conv_backward_filter",
                "mnv_conv_backward_filter_bias"):
        off = 4 if name.startswith("mnv_conv_forward") or name.endswith("_bias") else 3
        N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw = a[off:off + 11]
        Ho, Wo = (H + 2 * ph - fh) // sv + 1, (W + 2 * pw - fw) // sh + 1
        return 2.0 * N * Ho * Wo * Co * Ci * fh * fw
    return None


# ------------------------------------------------------------------------------------------------
# CPU leg: the same owl.net graph on the oracle backend (test infrastructure used as a baseline only)
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

    import numpy as np
    import torch
    import torch.distributed as dist

    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    dev_id = owl.create_gpu_device(local)
    owl.set_device(dev_id)
    owl.set_seed(1234)                      # identical initial weights on every rank
    batch = args.batch or wl["batch"]
    net = getattr(onet, wl["builder"])()
    net.batch_size = batch * world          # the update divisor is the global batch
    x, onehot = host_batch(wl, net.input_shape, batch, 100 + rank)
    du = net.get_data_unit()
    du.data, du.label = owl.from_numpy(x), owl.from_numpy(onehot)
    trainer = onet.NetTrainer(net, dist if world > 1 else None, fused_update=not args.unfused_update, merge=args.merge)
    gdev = rt.current_device()

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

    # ---- warm-up, then the timed region: inputs resident in HBM --------------------------------------
    for _ in range(max(3, args.warmup)):
        trainer.step()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.mnv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(gdev.stream)
    for _ in range(args.steps):
        trainer.step()
    e1.record(gdev.stream)
    sync_all()
    launches = lib.mnv_launch_count() - launches0
    sampler.stop_flag = True
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = batch * world * args.steps / (ms_total * 1e-3)
    loss = net.get_loss_units()[-1].getloss()

    # ---- end to end: the batch comes from pinned host memory every step, the loss goes back ----------
    e2e = None
    if not args.no_e2e:
        hx = torch.from_numpy(x.reshape(-1)).pin_memory()
        hy = torch.from_numpy(onehot.reshape(-1)).pin_memory()
        copy_stream = torch.cuda.Stream()
        bufs = [(owl.zeros(du.data.shape), owl.zeros(du.label.shape)) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i])
                bufs[i][0].as_torch().copy_(hx, non_blocking=True)
                bufs[i][1].as_torch().copy_(hy, non_blocking=True)
                ready[i].record(copy_stream)

        for i in range(2):
            consumed[i].record(gdev.stream)
        sync_all()
        steps = args.steps

        host_loss = [torch.zeros(1).pin_memory() for _ in range(2)]
        loss_ready = [torch.cuda.Event(), torch.cuda.Event()]

        def run(nsteps, pipelined):
            """pipelined: the step's loss is reduced on the device, copied to pinned host memory asynchronously and
            READ one step later (every step's loss is read inside the loop, the last one right after it), so the launch
            queue never drains; otherwise a blocking read after every step."""
            loss_val, pending = None, None
            prefetch(0)
            for it in range(nsteps):
                cur = it % 2
                if it + 1 < nsteps:
                    prefetch(1 - cur)
                gdev.stream.wait_event(ready[cur])
                du.data, du.label = bufs[cur]
                trainer.step()
                consumed[cur].record(gdev.stream)
                lu = net.get_loss_units()[-1]
                if not pipelined:
                    loss_val = lu.getloss()                       # blocking D2H read of the step's result
                    continue
                dsum, n = lu.getloss_device()
                host_loss[cur].copy_(dsum.as_torch(), non_blocking=True)    # D2H on the compute stream
                loss_ready[cur].record(gdev.stream)
                if pending is not None:                           # the previous step's loss: read it now
                    loss_ready[pending[0]].synchronize()
                    loss_val = -float(host_loss[pending[0]][0]) / pending[1]
                pending = (cur, n)
            if pending is not None:
                loss_ready[pending[0]].synchronize()
                loss_val = -float(host_loss[pending[0]][0]) / pending[1]
            return loss_val

        run(max(2, min(args.warmup, 4)), True)      # untimed: first-touch cost of the pinned staging path
        sync_all()
        t0 = time.perf_counter()
        step_loss = run(steps, True)
        sync_all()
        dt = max_over_ranks(time.perf_counter() - t0)
        sync_all()
        t0 = time.perf_counter()
        step_loss_b = run(steps, False)
        sync_all()
        dt_b = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": batch * world * steps / dt, "unit": "images/s",
               "h2d_bytes_per_step": int(hx.numel() * 4 + hy.numel() * 4), "d2h_bytes_per_step": 4,
               "last_loss": float(step_loss),
               "loss_read": "every step's loss is reduced on the device, copied to pinned host memory and read one step "
                            "later (double-buffered like the input), inside the timed region",
               "blocking_read_value": batch * world * steps / dt_b,
               "blocking_read_note": "same loop with a blocking loss read after every step (drains the launch queue)"}
        du.data, du.label = bufs[0]

    # ---- per-launch device times of the dominant kernel (rank 0) -------------------------------------
    roofline, optable = None, None
    pk = peaks()
    # every rank takes the two instrumented steps (they contain the gradient all-reduce); rank 0 records
    if rank == 0:
        rt.profiler = rt.EventProfiler()
    for _ in range(2):
        trainer.step()
    sync_all()
    if rank == 0:
        table = rt.profiler.table()
        rt.profiler = None
        per = {}
        g_flops = g_ms = 0.0
        g_n = 0
        for name, a, ms in table:
            d = per.setdefault(name, [0, 0.0])
            d[0] += 1
            d[1] += ms
            fl = gemm_flops(name, a)
            if fl is not None:
                g_flops += fl
                g_ms += ms
                g_n += 1
        total_ms = sum(v[1] for v in per.values())
        optable = {k: {"calls_per_step": v[0] / 2, "ms_per_step": v[1] / 2, "share": v[1] / total_ms}
                   for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])}
        tf32_peak = pk["bf16_tflops_sustained"] / 2.0
        achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms else 0.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": achieved / tf32_peak, "traffic": ncu_traffic(args.workload),
                    "kernel": "mnv::umma_gemm_kernel + mnv::conv_shift_fwd_kernel (tcgen05 TF32: MatMult + conv fwd/bwd-data/bwd-filter, pre-passes included)",
                    "launches_timed": g_n, "avg_launch_ms": g_ms / max(g_n, 1),
                    "share_of_step": g_ms / total_ms if total_ms else None,
                    "peak_source": "%s bf16_tflops_sustained/2 (dense TF32 = half the bf16 rate)" % pk["source"]}
    if world > 1:
        dist.barrier()

    # ---- CPU baseline on a bounded sample (rank 0, N == 1) ---------------------------------------------
    cpu_baseline = None
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

    if rank == 0:
        metric = "AlexNet train images/s" if args.workload == "alexnet" else args.workload + " train images/s"
        line = {
            "metric": metric, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 tensor-core conv/GEMM, f32 elsewhere",
            "data": "synthetic",
            "config": {"workload": wl["name"], "global_batch": batch * world, "per_gpu_batch": batch,
                       "parallelism": "dp%d" % world, "update": "chain" if args.unfused_update else "fused momentum-SGD kernel",
                       "l2": "working set per step (~2 GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "gradient_merge": trainer.merge_kind, **({"gradient_merge_note": trainer.merge_note} if hasattr(trainer, "merge_note") else {}),
                       **({"tuning": args.mnv_opt} if args.mnv_opt else {})},
            "clocks": sampler.summary(), "gpu_launches": int(launches), "loss": float(loss),
            "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu_baseline, "op_table": optable,
            "peaks": {k: pk.get(k) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained", "source")},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
