"""Device layer of the owl binding: the role of minerva/device (GpuDevice streams + pooled memory,
minerva/device/device.cpp:129-222) with one stream and one kernel workspace per device."""
import torch

from .. import _lib

CPU_DEVICE = -1


class _Gpu:
    def __init__(self, index):
        self.index = index
        self.device = torch.device("cuda", index)
        with torch.cuda.device(index):
            self.stream = torch.cuda.Stream(device=index)
            self.workspace = torch.empty(_lib.load().mnv_workspace_bytes_hint(), dtype=torch.uint8, device=self.device)
        self.stream_ptr = self.stream.cuda_stream
        self.ws_ptr = self.workspace.data_ptr()
        self.ws_bytes = self.workspace.numel()


_devices = []      # id -> _Gpu or CPU_DEVICE
_current = None
_seed = [0x5EED, 0, 0]   # seed, call counter, rank salt (dropout masks only)
profiler = None    # set to an object with begin(name, args, dev) / end(token, dev) to time every C-ABI call


class EventProfiler:
    """Per-op-name device time table from CUDA events on the launching stream -- what the reference's
    ExecutionProfiler (minerva/profiler/execution_profiler.cpp:17-53) kept with wall clocks."""

    def __init__(self):
        self.records = []   # (name, args, start_event, end_event)

    def begin(self, name, args, dev):
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(dev.stream)
        return (name, args, e0)

    def end(self, tok, dev):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(dev.stream)
        self.records.append(tok + (e1,))

    def table(self):
        """-> list of (name, args, milliseconds); call after the streams are synchronised."""
        return [(n, a, e0.elapsed_time(e1)) for (n, a, e0, e1) in self.records]


def has_cuda():
    return int(torch.cuda.is_available())


def get_gpu_device_count():
    return torch.cuda.device_count()


def create_cpu_device():
    _devices.append(CPU_DEVICE)
    return len(_devices) - 1


def create_gpu_device(which):
    if not torch.cuda.is_available():
        raise _lib.MnvError("create_gpu_device(%d): no CUDA device; this build has no CPU path" % which)
    _lib.load()
    _devices.append(_Gpu(which))
    dev_id = len(_devices) - 1
    if _current is None:
        set_device(dev_id)
    return dev_id


def set_device(dev):
    global _current
    d = _devices[dev]
    _current = d
    if d is not CPU_DEVICE:
        torch.cuda.set_device(d.index)
        torch.cuda.set_stream(d.stream)


def current_device():
    if _current is None:
        raise _lib.MnvError("no device: call owl.create_gpu_device() first")
    if _current is CPU_DEVICE:
        raise _lib.MnvError("the CPU (basic) device is not provided by this build: the GPU path has no CPU fallback")
    return _current


def wait_for_all():
    for d in _devices:
        if d is not CPU_DEVICE:
            d.stream.synchronize()


def set_seed(seed):
    _seed[0] = int(seed) & 0xFFFFFFFF
    _seed[1] = 0


def set_rank_salt(rank):
    """Data-parallel replicas share one seed (identical initial weights) but must not share dropout masks:
    the trainer mixes the rank into the Bernoulli generator's seed only."""
    _seed[2] = int(rank) & 0xFFFF


class GraphSeeds:
    """Generator keys for a training step recorded into a CUDA graph: the k-th draw of the step uses
    (*word + k * 0x85EBCA77) ^ salt on the device (mnv_rand_bernoulli_ds) -- the key next_seed() would have returned for
    it -- and arm() stores the step's base word before every replay and advances the host counter by the step's draws."""

    def __init__(self, dev):
        self.word = torch.zeros(1, dtype=torch.int32, device=dev.device)
        self.calls = 0

    def next(self, salted=False):
        self.calls += 1
        return (self.calls * 0x85EBCA77) & 0xFFFFFFFF, ((_seed[2] * 0xC2B2AE35) & 0xFFFFFFFF) if salted else 0

    def arm(self):
        base = (_seed[0] * 0x9E3779B1 + _seed[1] * 0x85EBCA77) & 0xFFFFFFFF
        self.word.fill_(base - (1 << 32) if base >= (1 << 31) else base)      # enqueued on the device's stream, before the replay
        _seed[1] += self.calls


capture_seeds = None     # a GraphSeeds while NetTrainer records a step: generators take their key from the device word


def next_seed(salted=False):
    _seed[1] += 1
    s = (_seed[0] * 0x9E3779B1 + _seed[1] * 0x85EBCA77) & 0xFFFFFFFF
    return (s ^ (_seed[2] * 0xC2B2AE35)) & 0xFFFFFFFF if salted else s
