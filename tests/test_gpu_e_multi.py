"""Data-parallel gradient merge on real GPUs (needs >= 2 devices; skipped otherwise): the peer-memory reduce-scatter /
all-gather of owl/net/merge.py against the NCCL all-reduce, same net, same batches, two steps."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, merge, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import minerva_b200.owl as owl
    from minerva_b200.owl.net.net import _default_backend
    from minerva_b200.owl.net.trainer import NetTrainer
    from tests.test_net_cpu import _tiny_net, _batch
    owl.set_device(owl.create_gpu_device(rank))
    owl.set_seed(11)
    B = _default_backend()
    net = _tiny_net(B)
    net.batch_size = 8 * world
    if merge == "peer":          # the tiny net's gradient is a few KB: force the bucketed peer exchange this test is about
        from minerva_b200.owl.net.merge import PeerGradMerge
        PeerGradMerge.SMALL_BYTES = 0
    tr = NetTrainer(net, dist, fused_update=True, merge="peer" if merge == "peer-small" else merge)
    du = net.get_data_unit()
    for step in range(3):        # step 0 builds the peer layout (NCCL-merged), steps 1-2 run the peer exchange
        du.data, du.label = _batch(B, 8, seed=step, lo=8 * rank)
        tr.step()
    owl.wait_for_all()
    if rank == 0:
        np.savez(out, kind=tr.merge_kind, **{"w%d" % uid: net.units[uid].weight.to_numpy() for uid in net.get_weighted_unit_ids()},
                 **{"b%d" % uid: net.units[uid].bias.to_numpy() for uid in net.get_weighted_unit_ids()})
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_merge_matches_nccl(tmp_path):
    res = {}
    for merge in ("peer", "peer-small", "nccl"):
        out = str(tmp_path / (merge + ".npz"))
        mp.spawn(_worker, args=(2, _free_port(), merge, out), nprocs=2, join=True)
        res[merge] = np.load(out)
    assert str(res["peer"]["kind"]).startswith("reduce-scatter") and str(res["nccl"]["kind"]).startswith("NCCL")
    assert "one all-reduce" in str(res["peer-small"]["kind"])      # small nets: gradients in place in one flat buffer, one all-reduce
    for variant in ("peer", "peer-small"):
        for k in res[variant].files:
            if k == "kind":
                continue
            a, b = res[variant][k].astype(np.float64), res["nccl"][k].astype(np.float64)
            assert np.linalg.norm(a - b) <= 1e-5 * max(np.linalg.norm(b), 1e-12), (variant, k)   # two summation orders of two addends: equal up to rounding
