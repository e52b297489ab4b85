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
    BOUND_GRAPHS = 2     # graph mode: recordings bound to recurring input arrays (a double-buffered feed has two)

    def __init__(self, net, dist=None, fused_update=True, merge="auto", graph=False):
        """dist: an initialised torch.distributed module (or None for 1 GPU).
        graph: record the whole step (forward, backward, loss reduction, update) into a CUDA graph once and replay it
               (1 GPU, fused update): the step's ~100..750 launches cost one host call -- see _graph_step.
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
        self.graph = bool(graph)
        self._g = None                   # recorded step: dict(graph, key, seeds, data, label, loss), input copied in
        self._bound = []                 # recordings bound to recurring input arrays (no input copy)
        self._seen = []                  # the last few (data, label) pairs that went through the copying recording
        self._bound_made = 0
        self.graph_replays = 0           # replays so far; launches they stand for = graph_launches_per_step each
        self.graph_launches_per_step = 0
        self.loss_device = None          # graph mode: (1-element NArray of sum(ln(y) o label), batch) of the last step
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
                if self.peer.small and "one all-reduce" not in self.merge_kind:
                    self.merge_kind = ("gradients produced in place in one flat buffer, merged by one all-reduce after backward "
                                       "(whole gradient <= %d MB: the bucketed peer exchange costs more than the step)" % (self.peer.SMALL_BYTES >> 20))
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
        if self.graph:
            return self._graph_step()
        return self._eager_step()

    # -- the step as one CUDA graph -------------------------------------------------------------------
    def _graph_step(self):
        """The small configurations are bound by the host: LeNet / the MLP enqueue ~10..40 launches of a few microseconds
        each, GoogLeNet ~750.  The step's launch sequence does not depend on data, so it is recorded once on the device's
        stream (every buffer the ops allocate comes from the graph's private pool and keeps its address) and replayed.
        What varies between steps stays outside the recording: the input batch is copied into the recorded input arrays,
        the dropout keys are read from a device word stored before every replay (_runtime.GraphSeeds), and a change of the
        learning rate / momentum / weight decay / batch size records a new graph.  Same kernels, same order, same bits as
        the eager step.
        Input arrays that come back (the two device slots of a double-buffered feed) get a recording of their own the second
        time they are seen -- bound to those arrays, so their steps need no input copy (158 MB per AlexNet step, 0.11 ms); at
        most BOUND_GRAPHS of them.  Every recording updates the same weights; `loss_device` is the replayed recording's."""
        import torch
        import minerva_b200.owl._runtime as rt
        net = self.net
        if self.dist is not None or not self.fused_update:
            raise RuntimeError("NetTrainer(graph=True) records the 1-GPU step with the fused update")
        du = net.get_data_unit()
        dev = rt.current_device()
        feed = getattr(du, "feed", None)
        if feed is not None:             # FeedDataUnit: take the uploaded batch here, forward() is not run on a replay
            if not du._started:
                feed.start()
                du._started = True
            feed.done()
            data, label = feed.next()
        else:
            data, label = du.data, du.label
        key = (net.current_lr, net.base_weight_decay, net.momentum, net.batch_size, tuple(data.shape), tuple(label.shape), id(dev))
        if self._g is not None and self._g["key"] != key:
            self._g, self._bound, self._seen, self._bound_made = None, [], [], 0     # hyper-parameters changed: every recording is stale
        for b in self._bound:                                        # a recording bound to exactly these input arrays
            if b["data"] is data and b["label"] is label:
                b["seeds"].arm()
                b["graph"].replay()
                self.loss_device = b["loss"]
                self.graph_replays += 1
                return
        if self._g is not None and data is not self._g["data"] and self._bound_made < 4 * self.BOUND_GRAPHS:
            if any(d is data and l is label for d, l in self._seen):
                if len(self._bound) >= self.BOUND_GRAPHS:            # a new feed took over: its slots replace the oldest ones
                    self._bound.pop(0)
                self._bound_made += 1                                # (bounded: rotating arrays must not re-record forever)
                self._bound.append(self._record(key, data, label, dev, bind=True))
                b = self._bound[-1]
                b["seeds"].arm()
                b["graph"].replay()
                self.loss_device = b["loss"]
                self.graph_replays += 1
                return
            self._seen = (self._seen + [(data, label)])[-4:]
        if self._g is None:
            if any(net.units[uid].weight is None for uid in net.get_weighted_unit_ids()):
                # the very first step runs eagerly: lazy initialisation (weight fillers, kernel attributes, driver entry
                # points) is not recordable; the next call records
                du.data, du.label = data, label
                if feed is not None:
                    du.feed = None
                try:
                    self._eager_step()
                    lu = net.get_loss_units()
                    self.loss_device = lu[-1].getloss_device() if lu and hasattr(lu[-1], "getloss_device") else None
                finally:
                    if feed is not None:
                        du.feed = feed
                return
            self._g = self._record(key, data, label, dev)
        g = self._g
        self.loss_device = g["loss"]
        if data is not g["data"]:
            g["data"].as_torch().copy_(data.as_torch(), non_blocking=True)
        if label is not g["label"]:
            g["label"].as_torch().copy_(label.as_torch(), non_blocking=True)
        g["seeds"].arm()
        g["graph"].replay()
        self.graph_replays += 1

    def _record(self, key, data, label, dev, bind=False):
        """-> the recording (dict).  bind: the graph reads `data` / `label` themselves instead of arrays of its own."""
        import torch
        import minerva_b200.owl._runtime as rt
        from minerva_b200 import _lib
        net = self.net
        du = net.get_data_unit()
        owl = net.B.owl
        static_data, static_label = (data, label) if bind else (owl.zeros(list(data.shape)), owl.zeros(list(label.shape)))
        seeds = rt.GraphSeeds(dev)
        feed = getattr(du, "feed", None)
        dev.stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        du.data, du.label = static_data, static_label
        if feed is not None:
            du.feed = None               # the recording reads the static arrays; the feed is driven by _graph_step
        lib = _lib.load()
        n0 = lib.mnv_launch_count()
        rt.capture_seeds = seeds
        # graph nodes already follow each other without a launch gap; programmatic edges measured 1 % slower than plain ones
        pdl = lib.mnv_set_dependent_launch(0)
        try:
            with torch.cuda.graph(graph, stream=dev.stream):
                self._eager_step()
                lu = net.get_loss_units()
                loss = lu[-1].getloss_device() if lu and hasattr(lu[-1], "getloss_device") else None
        finally:
            rt.capture_seeds = None
            lib.mnv_set_dependent_launch(pdl)
            if feed is not None:
                du.feed = feed
        rt.set_device(rt._devices.index(dev))      # torch.cuda.graph restores ITS entry stream; make the device's current again
        self.graph_launches_per_step = int(lib.mnv_launch_count() - n0)
        return dict(graph=graph, key=key, seeds=seeds, data=static_data, label=static_label, loss=loss)

    def _eager_step(self):
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
