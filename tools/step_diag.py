#!/usr/bin/env python
"""Where does a whole training step on the GPU leave the TF32-operand oracle?  Per-unit forward outputs and gradients of one
net at a tiny batch: GPU vs fp32 oracle, GPU vs TF32-operand oracle, TF32-operand oracle vs fp32 oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minerva_b200.owl as owl
from tests.test_gpu_f_fullsize import _step_nets
owl.set_device(owl.create_gpu_device(0))
name = sys.argv[1] if len(sys.argv) > 1 else "alexnet"
cfg = {"alexnet": ("build_alexnet", [227, 227, 3], 1000, 2, 14), "googlenet": ("build_googlenet", [224, 224, 3], 1000, 2, 13)}[name]
fp32, tf32, gpu = _step_nets(owl, *cfg)
def nr(a, b):
    a, b = a.to_numpy().astype(np.float64), b.to_numpy().astype(np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30)
print("%-28s %12s %12s %12s" % ("unit output", "gpu/tf32orc", "gpu/fp32orc", "tf32/fp32orc"))
for uf, ut, ug in zip(fp32.units, tf32.units, gpu.units):
    if uf.out is None:
        continue
    print("%-28s %12.2e %12.2e %12.2e" % (uf.name, nr(ut.out, ug.out), nr(uf.out, ug.out), nr(uf.out, ut.out)))
print("gradients")
for uid in fp32.get_weighted_unit_ids():
    uf, ut, ug = fp32.units[uid], tf32.units[uid], gpu.units[uid]
    print("%-28s %12.2e %12.2e %12.2e" % (uf.name + ".w", nr(ut.weightgrad, ug.weightgrad), nr(uf.weightgrad, ug.weightgrad), nr(uf.weightgrad, ut.weightgrad)))
