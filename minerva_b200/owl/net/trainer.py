"""Data-parallel training loop (reference: owl/owl/net/trainer.py:101-148).

The reference runs every replica from one Python thread (`for gpuid: owl.set_device(...)`) and
merges gradients by copying them to a per-layer "update GPU" and adding (trainer.py:126-138); the
updated weight is then re-read by the other GPUs on their next use.  Here there is one process per
GPU (torchrun); the merge is an NCCL all-reduce over NVLink/NVSwitch issued per weighted unit the
moment its dW/db exist -- i.e. in reverse layer order, overlapping the rest of backward -- and
every rank applies the identical update to its own replica, so no weight broadcast is needed.
The update divisor stays the GLOBAL batch (net.py:252-254).
"""
import time


class NetTrainer(object):
    def __init__(self, net, dist=None, fused_update=True, merge="auto"):
        """dist: an initialised torch.distributed module (or None for 1 GPU).
        fused_update: one mnv_sgd_momentum_update per tensor instead of the reference's op chain.
        merge: "peer" = reduce-scatter / all-gather over NVLink peer memory on the copy engines (merge.py),
               "nccl" = one NCCL all-reduce per gradient, "auto" = peer when the process group is NCCL on CUDA and the
               symmetric-memory rendezvous works, else nccl (gloo in the CPU tests)."""
        self.net = net
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.fused_update = fused_update
        self._pending = []
        self.peer = None
        self.merge_kind = "none (1 GPU)"
        if self.dist and hasattr(net.B.owl, "set_rank_salt"):
            net.B.owl.set_rank_salt(self.dist.get_rank())     # replicas draw different dropout masks, same weights
        if self.dist:
            self.merge_kind = "NCCL all-reduce per weighted unit, overlapped with backward"
            if merge in ("auto", "peer") and self.dist.get_backend() == "nccl":
                try:
                    from .merge import PeerGradMerge
                    import minerva_b200.owl._runtime as rt
                    self.peer = PeerGradMerge(net, self.dist, rt.current_device())
                    self.merge_kind = ("reduce-scatter + all-gather over NVLink peer (symmetric) memory on the copy engines, "
                                       "per bucket, overlapped with backward; local shard sum = the only kernel")
                except Exception as ex:     # no symmetric-memory support in this build / topology
                    if merge == "peer":
                        raise
                    self.peer = None
                    self.merge_note = "peer merge unavailable: %r" % (ex,)
        net.on_weight_grad = self._on_weight_grad if self.dist else None
        if merge == "off" and self.dist:      # diagnostics only: replicas drift apart; measures the step without any exchange
            net.on_weight_grad, self.peer = None, None
            self.merge_kind = "OFF (diagnostic run, not data-parallel training)"

    # -- gradient merge -------------------------------------------------------------------------
    def _on_weight_grad(self, unit):
        if self.peer is not None:
            try:
                self.peer.on_weight_grad(unit)
                return
            except Exception as ex:
                if self.peer.flat is not None:
                    raise           # half-way through a step: nothing sane to fall back to
                self._peer_failed(ex)
        d = self.dist
        for g in (unit.weightgrad, unit.biasgrad):
            self._pending.append(d.all_reduce(g.as_torch(), op=d.ReduceOp.SUM, async_op=True))

    def _peer_failed(self, ex):
        self.peer = None
        self.merge_kind = "NCCL all-reduce per weighted unit, overlapped with backward"
        self.merge_note = "peer merge unavailable: %r" % (ex,)

    def _wait_merge(self):
        if self.peer is not None:
            try:
                self.peer.finish_step()
            except Exception as ex:
                if self.peer.flat is not None:
                    raise
                # the rendezvous failed while building the layout: this step's gradients were already all-reduced
                self._peer_failed(ex)
        for w in self._pending:
            w.wait()
        self._pending = []

    # -- one iteration ------------------------------------------------------------------------------
    def step(self):
        net = self.net
        if self.peer is not None:
            self.peer.begin_step()
        net.forward("TRAIN")
        net.backward("TRAIN")
        self._wait_merge()
        if not self.fused_update:
            net.weight_update()
            return
        NArray = net.B.owl.NArray
        lr, wd, mom, bs = net.current_lr, net.base_weight_decay, net.momentum, net.batch_size
        entries = []
        for uid in net.get_weighted_unit_ids():
            u = net.units[uid]
            entries.append((u.weight, u.weightdelta, u.weightgrad, lr * u.lr_mult_w / bs, lr * u.lr_mult_w * wd * u.decay_mult_w))
            entries.append((u.bias, u.biasdelta, u.biasgrad, lr * u.lr_mult_b / bs, lr * u.lr_mult_b * wd * u.decay_mult_b))
        if hasattr(NArray, "sgd_update_multi"):          # one launch for every parameter tensor of the net
            NArray.sgd_update_multi(entries, mom)
        else:
            for e in entries:
                NArray.sgd_update(e[0], e[1], e[2], mom, e[3], e[4])
        for uid in net.get_weighted_unit_ids():
            net.units[uid].weightgrad = net.units[uid].biasgrad = None

    def run(self, iters, sync_freq=1, log=None):
        """trainer.py:101-148: img/s = batch_size * sync_freq / wall time between wait_for_all() calls."""
        owl = self.net.B.owl
        last = time.time()
        speeds = []
        for it in range(iters):
            self.step()
            if (it + 1) % sync_freq == 0:
                owl.wait_for_all()
                now = time.time()
                speeds.append(self.net.batch_size * sync_freq / (now - last))
                if log:
                    log("Finished training %d minibatch (speed: %.1f img/s)" % (it + 1, speeds[-1]))
                last = time.time()
        return speeds
