"""Data layer (reference: owl/owl/net/netio.py:259-339 LMDBDataProvider + the data units of
owl/owl/net/net.py, and ArrayLoader's upload on the default stream, minerva/op/impl/cuda.cpp:592-597).

The reference converts every minibatch on the host -- uint8 Datum minus the mean image, random crop, optional mirror,
astype(float32) -- and uploads 4 bytes per pixel with a blocking copy.  Here:

  * a provider fills PINNED host buffers with the STORED uint8 images (1 byte per pixel), the per-image (crop_y, crop_x,
    mirror) triples the reference draws with numpy (netio.py:303-311) and the labels;
  * `HostFeed` uploads them on a copy stream into one of two device buffer sets and runs the transform kernel
    (mnv_image_transform_u8: mean-subtract, crop, mirror, scale -> fp32) behind the copy, on the same stream, so the
    compute stream only ever waits on an event -- never the host -- and step k+1's upload overlaps step k's compute;
  * `FeedDataUnit` is the Net's data unit over such a feed (hand-over and buffer release are stream-ordered events).

No LMDB / JPEG decoding here (no such library in the image): `SyntheticImageProvider` stands in for the
database cursor with images of the stored shape.  The fp32 path (`data_f32=`) uploads an already converted batch, for
comparison with what the reference moves over PCIe.
"""
import numpy as np
import torch

from .net import DataUnit


class SyntheticImageProvider(object):
    """Stands in for LMDBDataProvider.get_mb (netio.py:289-331): yields (uint8 images [N,C,S,S], one-hot labels [N,classes])
    of the stored size, a fresh minibatch per call drawn from a fixed pool (no per-step host RNG cost in the timed loop)."""

    def __init__(self, batch, channels, stored_hw, classes, seed=0, pool=2):
        rs = np.random.RandomState(seed)
        self.images = [rs.randint(0, 256, (batch, channels) + tuple(stored_hw), dtype=np.uint8) for _ in range(pool)]
        self.labels = []
        for _ in range(pool):
            lab = np.zeros((batch, classes), np.float32)
            lab[np.arange(batch), rs.randint(0, classes, batch)] = 1
            self.labels.append(lab)
        self.k = 0

    def get_mb(self):
        i = self.k % len(self.images)
        self.k += 1
        return self.images[i], self.labels[i]


class HostFeed(object):
    """Double-buffered, stream-ordered host -> device feed of one training process.

        feed.start()                     upload step 0
        data, label = feed.next()        this step's NArrays (compute stream waits on the upload's event); step k+1's upload starts
        ... enqueue the step ...
        feed.done()                      the step's reads are enqueued: its buffers may be overwritten after them

    data_u8 [N,C,S,S] uint8 (+ mean [C,S,S] fp32, scale, crop (h, w), mirror): the stored images are uploaded and
    transformed on the device.  data_f32: an fp32 batch uploaded as is.  `provider` (optional): object with get_mb() ->
    (uint8 images, labels) called once per step instead of re-sending the fixed batch."""

    def __init__(self, owl, rt, data_u8=None, data_f32=None, label=None, mean=None, scale=1.0, crop=None, mirror=False,
                 provider=None, train=True, seed=0):
        from ... import _lib
        self.owl, self.rt, self.lib, self._check = owl, rt, _lib.load(), _lib.check
        self.dev = rt.current_device()
        dev = self.dev.device
        self.provider, self.train, self.mirror = provider, train, bool(mirror)
        self.rs = np.random.RandomState(seed)
        self.copy_stream = torch.cuda.Stream(device=dev)
        if provider is not None and data_u8 is None:
            data_u8, label = provider.get_mb()
        self.u8 = data_u8 is not None
        src = data_u8 if self.u8 else data_f32
        self.N = int(src.shape[0])
        if self.u8:
            if src.ndim == 2:                       # flat vectors (MNIST MLP): one channel, one row
                src = src.reshape(self.N, 1, 1, src.shape[1])
            _, self.C, self.sh, self.sw = src.shape
            self.ch, self.cw = (crop if crop else (self.sh, self.sw))
            self.scale = float(scale)
            self.h_img = torch.from_numpy(np.ascontiguousarray(src)).pin_memory()
            # pinned staging is per slot wherever the host rewrites it between uploads (an async copy reads it later)
            self.h_off = [torch.zeros((self.N, 3), dtype=torch.int32).pin_memory() for _ in range(2)]
            self.need_off = (self.ch, self.cw) != (self.sh, self.sw) or self.mirror
            self.mean = None
            if mean is not None:                    # [C, S, S] mean image (mean_file) or C per-channel values (mean_value), netio.py:271-283
                m = np.asarray(mean, np.float32)
                if m.size == self.C:
                    m = np.broadcast_to(m.reshape(self.C, 1, 1), (self.C, self.sh, self.sw))
                assert m.size == self.C * self.sh * self.sw, "mean must be [C, stored_h, stored_w] or C values"
                self.mean = torch.from_numpy(np.ascontiguousarray(m).reshape(-1)).to(dev)
            self.d_img = [torch.empty_like(self.h_img, device=dev) for _ in range(2)]
            self.d_off = [torch.zeros((self.N, 3), dtype=torch.int32, device=dev) for _ in range(2)]
            numel = self.N * self.C * self.ch * self.cw
        else:
            self.h_img = torch.from_numpy(np.ascontiguousarray(src, np.float32).reshape(-1)).pin_memory()
            numel = self.h_img.numel()
        self.src_shape = tuple(src.shape)
        self.h_lab = torch.from_numpy(np.ascontiguousarray(label, np.float32).reshape(-1)).pin_memory()
        self.h_img_slot, self.h_lab_slot = [self.h_img, self.h_img], [self.h_lab, self.h_lab]
        if provider is not None:                    # fresh data every step: one pinned image / label buffer per slot
            self.h_img_slot[1], self.h_lab_slot[1] = self.h_img.clone().pin_memory(), self.h_lab.clone().pin_memory()
        self.lab_shape = list(reversed(label.shape))
        NArray = owl.NArray
        self.bufs = [(NArray(torch.empty(numel, dtype=torch.float32, device=dev), self._data_shape(), self.dev),
                      NArray(torch.empty(self.h_lab.numel(), dtype=torch.float32, device=dev), self.lab_shape, self.dev))
                     for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.cur = 0
        self.issued = 0
        self.taken = 0
        self.outstanding = None

    def _data_shape(self):
        if self.u8:
            if self.sh == 1 and self.C == 1:
                return [self.cw, self.N]
            return [self.cw, self.ch, self.C, self.N]
        return list(reversed(self.src_shape))

    # ------------------------------------------------------------------------------------------------------------
    def _draw_offsets(self, i):
        """netio.py:303-311: TRAIN draws a random crop corner per image (and a mirror coin), TEST takes the centre."""
        o = self.h_off[i].numpy()
        if self.train:
            o[:, 0] = self.rs.randint(0, self.sh - self.ch + 1, self.N)
            o[:, 1] = self.rs.randint(0, self.sw - self.cw + 1, self.N)
            o[:, 2] = (self.rs.rand(self.N) > 0.5) if self.mirror else 0
        else:
            o[:, 0], o[:, 1], o[:, 2] = (self.sh - self.ch) // 2, (self.sw - self.cw) // 2, 0

    def _prefetch(self, i):
        if self.issued >= 2:
            self.ready[i].synchronize()                # slot i's previous upload has left its pinned staging buffers
        if self.provider is not None and self.issued > 0:
            img, lab = self.provider.get_mb()          # the host-side copy into pinned memory is the loader's only per-step work
            self.h_img_slot[i].numpy()[...] = img.reshape(self.h_img.shape)
            self.h_lab_slot[i].numpy()[...] = lab.reshape(-1)
        if self.u8 and self.need_off:
            self._draw_offsets(i)
        data, label = self.bufs[i]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[i])
            label.as_torch().copy_(self.h_lab_slot[i], non_blocking=True)
            if self.u8:
                self.d_img[i].copy_(self.h_img_slot[i], non_blocking=True)
                if self.need_off:
                    self.d_off[i].copy_(self.h_off[i], non_blocking=True)
                rc = self.lib.mnv_image_transform_u8(self.d_img[i].data_ptr(), self.mean.data_ptr() if self.mean is not None else None,
                                                     self.d_off[i].data_ptr() if self.need_off else None, data.as_torch().data_ptr(),
                                                     self.N, self.C, self.sh, self.sw, self.ch, self.cw, self.scale,
                                                     self.copy_stream.cuda_stream)
                if rc:
                    self._check(rc, "mnv_image_transform_u8")
            else:
                data.as_torch().copy_(self.h_img_slot[i], non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.issued += 1

    def start(self):
        for i in range(2):
            self.consumed[i].record(self.dev.stream)
        self.cur, self.issued, self.taken, self.outstanding = 0, 0, 0, None
        self._prefetch(0)

    def next(self):
        i = self.cur
        self._prefetch(1 - i)                       # the NEXT step's upload overlaps this step's compute
        self.dev.stream.wait_event(self.ready[i])
        self.outstanding = i
        self.cur = 1 - i
        self.taken += 1
        return self.bufs[i]

    def done(self):
        if self.outstanding is not None:
            self.consumed[self.outstanding].record(self.dev.stream)
            self.outstanding = None

    def stop(self):
        self.done()
        self.copy_stream.synchronize()

    def resident(self):
        """A batch left resident on the device (after stop()): what the device-timed loop reads every step."""
        return self.bufs[0]

    def bytes_per_step(self):
        return int(self.h_img.numel() * self.h_img.element_size() + self.h_lab.numel() * 4 + (self.N * 12 if self.u8 and self.need_off else 0))


class FeedDataUnit(DataUnit):
    """The Net's data unit over a HostFeed: forward() releases the previous step's buffers (every read of them is enqueued
    by then), takes the next uploaded batch and starts the upload after it.  Reference: LMDBDataUnit.forward calls
    get_mb() and owl.from_numpy() -- a blocking conversion + copy -- every iteration (owl/owl/net/net.py data units)."""

    def __init__(self, name, tops, feed=None):
        super().__init__(name, tops)
        self.feed = feed
        self._started = False

    def forward(self, from_btm, to_top, phase):
        if self.feed is not None:
            if not self._started:
                self.feed.start()
                self._started = True
            self.feed.done()
            self.data, self.label = self.feed.next()
        super().forward(from_btm, to_top, phase)

    def close(self):
        if self.feed is not None and self._started:
            self.feed.stop()
            self._started = False
