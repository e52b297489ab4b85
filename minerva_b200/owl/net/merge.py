"""Gradient merge over NVLink peer memory without taking SMs from the compute stream.

The reference merges gradients by copying every replica's dW to one "update GPU" and adding there
(owl/owl/net/trainer.py:126-138).  The first B200 version used one NCCL all-reduce per weighted unit; measured at
1 -> 2 -> 8 GPUs the step grew 5.87 -> 6.25 -> 6.67 ms, and the per-op table showed where: the convolution
backward-filter and the FC backward GEMMs that run next to the all-reduce kernels lost 0.3-0.5 ms, i.e. the NCCL kernels'
SMs and HBM bandwidth came straight out of the persistent tensor-core kernel (one CTA per SM).

Here the exchange is a reduce-scatter + all-gather written on symmetric (peer-mapped) memory in which every byte that
crosses NVLink is moved by the copy engines (cudaMemcpyAsync between peer-mapped buffers), and the only kernel is the
local sum of a 1/W shard:

  flat   one symmetric fp32 buffer holding every gradient of the net, in the order backward produces them, cut into
         buckets (the FC gradients, 94 % of the bytes, are complete a third of the way through backward);
  per bucket, on a side stream, as soon as its last gradient has been copied into `flat`:
      barrier                                     every rank's bucket is in place
      pull   recv[p] <- peer p's flat[bucket shard r]          (W-1 copy-engine transfers, shard r = this rank's)
      sum    flat[shard r] += recv[p], p in rank order         (one mnv_add_n on the side stream, 1/W of the bucket)
      barrier                                     every shard is reduced, nobody still reads unreduced data
      pull   flat[shard p] <- peer p's flat[shard p]           (W-1 copy-engine transfers)
      barrier                                     nobody overwrites `flat` while a peer still pulls
  the update then reads the merged gradients from `flat` (every rank holds identical bits: a shard is summed once).

torch is used for what the task allows it for: the symmetric allocation and rendezvous
(torch.distributed._symmetric_memory), streams, events and the peer-mapped tensor views.
"""
import torch

from ... import _lib


def plan_layout(sizes, world, last_fc):
    """sizes: [(unit name, weight-gradient elements, bias-gradient elements)] in backward order; last_fc: index of the
    last fully-connected unit in that order (-1: none).  -> (slices, buckets, total elements):
    slices[(name, "w" | "b")] = (offset, elements), 16-byte aligned; buckets = [(offset, elements, closing unit)], each a
    multiple of 4 * world elements so that it splits into `world` shards of whole 16-byte vectors.
    Bucket boundaries: after the last fully-connected unit (the FC gradients, most of the bytes, come first in backward),
    before the last unit, and at the end -- the last bucket's exchange is the only one nothing hides, so it holds a
    single unit (AlexNet conv1: 140 KB)."""
    align = 4 * world
    slices, buckets, off, cur_start = {}, [], 0, 0
    cuts = {last_fc, len(sizes) - 2, len(sizes) - 1}
    for i, (name, nw, nb) in enumerate(sizes):
        for tag, n in (("w", nw), ("b", nb)):
            slices[(name, tag)] = (off, n)
            off += (n + 3) // 4 * 4
        if i in cuts:
            end = (off + align - 1) // align * align
            buckets.append((cur_start, end - cur_start, name))
            off = cur_start = end
    return slices, buckets, off


class PeerGradMerge(object):
    SMALL_BYTES = 8 << 20      # total gradient bytes up to which one all-reduce replaces the bucketed peer exchange

    def __init__(self, net, dist, dev):
        import torch.distributed._symmetric_memory as symm_mem
        self.net, self.dist, self.dev = net, dist, dev
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.symm_mem = symm_mem
        self.comm = torch.cuda.Stream(device=dev.device)
        self.flat = None          # built on the first step, when every gradient's size is known
        self.slices = {}          # (unit name, "w" | "b") -> (offset, numel)
        self.buckets = []         # (offset, numel, name of the unit whose gradients complete the bucket)
        self.done = torch.cuda.Event()
        self.lanes = []           # side streams for concurrent peer transfers
        self.small = False
        self._arrived = 0

    # -- layout ---------------------------------------------------------------------------------------------------
    def _build(self, order):
        """order: [(unit, weightgrad, biasgrad)] in the order backward produced them on the first (NCCL-merged) step."""
        W = self.world
        from .net import FullyConnection
        last_fc = max([i for i, (u, _, _) in enumerate(order) if isinstance(u, FullyConnection)], default=-1)
        slices, buckets, off = plan_layout([(u.name, gw.size, gb.size) for u, gw, gb in order], W, last_fc)
        self.slices = slices
        self.buckets = buckets
        total = off
        # A net whose whole gradient is a few MB (LeNet 1.7 MB, the MLP 0.9 MB) finishes backward in ~0.1 ms: three bucket
        # exchanges of three barriers each cost more than the step.  Its gradients are still produced in place in `flat`,
        # and finish_step merges the whole buffer with ONE NCCL all-reduce (latency-optimal at this size; NVLS on NVSwitch).
        self.small = total * 4 <= self.SMALL_BYTES
        if self.small:
            self.buckets = []
        if self.small:
            self.flat = torch.zeros(total, dtype=torch.float32, device=self.dev.device)     # plain memory: nothing is peer-mapped
        else:
            shard_max = max(n // W for _, n, _ in buckets)
            sm = self.symm_mem
            self.flat = sm.empty(total, dtype=torch.float32, device=self.dev.device)
            self.flat.zero_()
            self.hdl = sm.rendezvous(self.flat, self.dist.group.WORLD)
            self.recv = torch.empty((W, shard_max), dtype=torch.float32, device=self.dev.device)
        for u, _, _ in order:     # from now on the units compute their gradients straight into `flat`
            u.grad_out = tuple(self.flat[o:o + n] for o, n in (self.slices[(u.name, "w")], self.slices[(u.name, "b")]))
        torch.cuda.synchronize(self.dev.device)

    # -- per step -------------------------------------------------------------------------------------------------
    def begin_step(self):
        self._arrived = 0
        self._first = []
        self.net.B.owl.NArray.place_next()      # no stale placement survives an aborted step

    def on_weight_grad(self, unit):
        """Called by Net.backward as soon as `unit`'s gradients exist (compute stream)."""
        if self.flat is None:
            self._first.append((unit, unit.weightgrad, unit.biasgrad))
            return
        NArray = self.net.B.owl.NArray
        for tag, attr in (("w", "weightgrad"), ("b", "biasgrad")):
            off, n = self.slices[(unit.name, tag)]
            g = getattr(unit, attr)
            view = self.flat[off:off + n]
            if g.as_torch().data_ptr() != view.data_ptr():            # not produced in place (unit without grad_out support)
                view.copy_(g.as_torch(), non_blocking=True)          # device-to-device, on the compute stream
                setattr(unit, attr, NArray(view, g.shape, self.dev))    # the update reads the merged values from `flat`
        for b, (boff, bn, last) in enumerate(self.buckets):
            if last == unit.name:
                self._exchange(b, boff, bn)

    def _pull_all(self, pairs):
        """dst.copy_(src) for every (dst, src) pair, each on its own side stream so that the transfers (one per peer) run
        on as many copy engines as the device has; joined back into the comm stream."""
        if len(pairs) <= 1:
            for dst, src in pairs:
                dst.copy_(src, non_blocking=True)
            return
        while len(self.lanes) < len(pairs):
            self.lanes.append(torch.cuda.Stream(device=self.dev.device))
        fork = torch.cuda.Event()
        fork.record(self.comm)
        for lane, (dst, src) in zip(self.lanes, pairs):
            lane.wait_event(fork)
            with torch.cuda.stream(lane):
                dst.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(lane)
            self.comm.wait_event(ev)

    def _exchange(self, b, boff, bn):
        W, r, hdl = self.world, self.rank, self.hdl
        shard = bn // W
        lib = _lib.load()
        ready = torch.cuda.Event()
        ready.record(self.dev.stream)
        self.comm.wait_event(ready)
        with torch.cuda.stream(self.comm):
            mine = self.flat[boff + r * shard: boff + (r + 1) * shard]
            hdl.barrier(channel=0)
            peers = [(r - step) % W for step in range(1, W)]
            self._pull_all([(self.recv[p, :shard], hdl.get_buffer(p, (shard,), torch.float32, boff + r * shard)) for p in peers])
            # ((mine + recv[p0]) + recv[p1]) + ... with p in rank order -- the bits of W - 1 chained accumulates -- in one
            # pass over the shard (mnv_add_n, in place on srcs[0]): (W + 1) x 4 bytes per element instead of (W - 1) x 12
            others = [p for p in range(W) if p != r]
            import ctypes
            while others:
                take, others = others[:7], others[7:]
                ptrs = (ctypes.c_void_p * (len(take) + 1))(mine.data_ptr(), *[self.recv[p].data_ptr() for p in take])
                rc = lib.mnv_add_n(ptrs, len(take) + 1, mine.data_ptr(), shard, self.comm.cuda_stream)
                if rc:
                    _lib.check(rc, "mnv_add_n")
            hdl.barrier(channel=0)
            self._pull_all([(self.flat[boff + p * shard: boff + (p + 1) * shard],
                             hdl.get_buffer(p, (shard,), torch.float32, boff + p * shard)) for p in peers])
            hdl.barrier(channel=0)
            self.done.record(self.comm)

    def finish_step(self):
        """Everything merged before the update reads it.  On the first step the layout is built and the gradients are
        merged by a plain all-reduce (no overlap, once)."""
        if self.flat is None:
            d = self.dist
            for u, gw, gb in self._first:
                d.all_reduce(gw.as_torch(), op=d.ReduceOp.SUM)
                d.all_reduce(gb.as_torch(), op=d.ReduceOp.SUM)
            # the switch to the peer exchange is collective: if the symmetric allocation / rendezvous fails on ANY
            # rank, every rank stays on the NCCL merge (a split decision would deadlock the first barrier)
            err = None
            try:
                self._build(self._first)
            except Exception as ex:
                err = ex
                self.flat = None
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.dev.device)
            d.all_reduce(ok, op=d.ReduceOp.MIN)
            self._first = []
            if int(ok.item()) == 0:
                self.flat = None
                for u in self.net.units:
                    if getattr(u, "grad_out", None) is not None:
                        u.grad_out = None
                raise RuntimeError("peer merge not available on every rank: %r" % (err,))
            return
        if self.small:
            self.dist.all_reduce(self.flat, op=self.dist.ReduceOp.SUM)      # stream-ordered after backward, before the update
            return
        self.dev.stream.wait_event(self.done)
