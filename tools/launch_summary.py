#!/usr/bin/env python
"""Per-kernel summary of an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum[,...]`
launch list: launches, total time, share of the captured window, DRAM bytes per launch.
Usage: launch_summary.py launches.csv [out.json]"""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr, data = rows[0], rows[1:]
ik, im, iv, iid, iu = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launch = collections.OrderedDict()
for r in data:
    d = launch.setdefault(r[iid], {"kernel": r[ik]})
    v = float(r[iv].replace(",", ""))
    if r[im].startswith("gpu__time"):
        v *= TIME.get(r[iu], 1.0)
    elif "bytes" in r[im]:
        v *= BYTES.get(r[iu], 1)
    d[r[im]] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in launch.values():
    k = d["kernel"].split("(")[0].replace("void mnv::", "").replace("mnv::", "").replace("void ", "")
    k = k.replace("(int)", "").replace("(bool)", "")
    a = agg[k]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a[3] += d.get("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * d.get("gpu__time_duration.sum", 0.0)
tot = sum(a[1] for a in agg.values())
out = []
print("%d launches, %.1f us total" % (len(launch), tot))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tc = a[3] / a[1] if a[1] else 0.0
    print("%-62s n=%3d %9.1f us %5.1f%%  dram %8.1f MB/launch  tc %4.1f%%" % (k[:62], a[0], a[1], 100 * a[1] / tot, a[2] / a[0] / 1e6, tc))
    out.append({"kernel": k, "launches": a[0], "time_us": a[1], "share": a[1] / tot, "dram_bytes_per_launch": a[2] / a[0],
                "tensor_pipe_active_pct_time_weighted": tc})
if len(sys.argv) > 2:
    # the tcgen05 kernels: the persistent GEMM / implicit-GEMM kernel and the shift-GEMM convolution kernel
    umma = [o for o in out if o["kernel"].startswith(("umma_gemm_kernel", "conv_shift_fwd_kernel"))]
    n = sum(o["launches"] for o in umma)
    json.dump({"source": sys.argv[1], "launches": len(launch), "total_us": tot, "kernels": out,
               "umma_gemm_kernel": {"launches": n, "share": sum(o["share"] for o in umma),
                                    "dram_bytes_per_launch": sum(o["dram_bytes_per_launch"] * o["launches"] for o in umma) / max(n, 1)}},
              open(sys.argv[2], "w"), indent=1)
