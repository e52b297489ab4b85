#include "device/stream_device.h"
#include <stdexcept>
#include <string>
#include "mnv.h"

namespace minerva {
namespace {
void CudaOk(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace

struct StreamDevice::Ev {
  cudaEvent_t ev;
  StreamDevice* owner;
  ~Ev() {   // back to the pool: events are created once and recycled
    std::lock_guard<std::mutex> l(owner->ev_mu_);
    owner->event_pool_.push_back(ev);
  }
};

size_t StreamDevice::SizeClass(size_t bytes) {
  if (bytes <= 256) return 256;
  size_t p = 256;
  while (p * 2 <= bytes) p *= 2;          // p <= bytes < 2p
  const size_t step = p / 8;              // 8 classes per power of two
  return (bytes + step - 1) / step * step;
}

StreamDevice::StreamDevice(uint64_t device_id, DeviceListener* listener, int gpu_id, Completion mode)
    : device_id_(device_id), listener_(listener), gpu_(gpu_id), mode_(mode), workspace_bytes_(mnv_workspace_bytes_hint()) {
  CudaOk(cudaSetDevice(gpu_), "cudaSetDevice");
  CudaOk(cudaFree(0), "context init");
  for (int i = 0; i < kStreams; ++i) CudaOk(cudaStreamCreateWithFlags(&streams_[i], cudaStreamNonBlocking), "cudaStreamCreate");
  if (mode_ == Completion::kEvent) completer_ = std::thread(&StreamDevice::CompletionLoop, this);
}

StreamDevice::~StreamDevice() {
  cudaSetDevice(gpu_);
  for (int i = 0; i < kStreams; ++i) cudaStreamSynchronize(streams_[i]);
  if (completer_.joinable()) {
    { std::lock_guard<std::mutex> l(cmu_); stop_ = true; }
    ccv_.notify_all();
    completer_.join();
  }
  {
    std::lock_guard<std::mutex> l(mu_);
    for (auto& kv : data_) cudaFree(kv.second.block.ptr);
    data_.clear();
    for (auto& kv : free_) for (Block& b : kv.second) cudaFree(b.ptr);
    free_.clear();
  }
  for (int i = 0; i < kStreams; ++i) { if (workspace_[i]) cudaFree(workspace_[i]); cudaStreamDestroy(streams_[i]); }
  std::lock_guard<std::mutex> l(ev_mu_);
  for (cudaEvent_t e : event_pool_) cudaEventDestroy(e);
}

StreamDevice::EvPtr StreamDevice::NewEvent() {
  cudaEvent_t e = nullptr;
  {
    std::lock_guard<std::mutex> l(ev_mu_);
    if (!event_pool_.empty()) { e = event_pool_.back(); event_pool_.pop_back(); }
  }
  if (!e) CudaOk(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
  return EvPtr(new Ev{e, this});
}

void* StreamDevice::Workspace(int i) {
  if (!workspace_[i]) CudaOk(cudaMalloc(&workspace_[i], workspace_bytes_), "workspace cudaMalloc");
  return workspace_[i];
}

// mu_ held.  A recycled block is ordered after every earlier use on OTHER streams by device-side waits; uses on the
// same stream are ordered by the stream itself.
StreamDevice::Block StreamDevice::Alloc(size_t bytes, int stream_idx) {
  const size_t cls = SizeClass(bytes);
  auto it = free_.find(cls);
  if (it != free_.end() && !it->second.empty()) {
    Block b = std::move(it->second.back());
    it->second.pop_back();
    for (int s = 0; s < kStreams; ++s) {
      if (s != stream_idx && b.last_use[s]) {
        CudaOk(cudaStreamWaitEvent(streams_[stream_idx], b.last_use[s]->ev, 0), "cudaStreamWaitEvent");
        ++stats_.cross_stream_waits;
      }
      if (s != stream_idx) b.last_use[s].reset();
    }
    ++stats_.pool_hits;
    return b;
  }
  Block b;
  b.bytes = cls;
  CudaOk(cudaMalloc(&b.ptr, cls), "cudaMalloc");
  ++stats_.cuda_mallocs;
  stats_.bytes_reserved += cls;
  return b;
}

void StreamDevice::PushTask(Task* task) {
  CudaOk(cudaSetDevice(gpu_), "cudaSetDevice");          // PreExecute (device.cpp:201-203)
  EvPtr ev;
  int si;
  {
    std::lock_guard<std::mutex> lck(mu_);
    // stream affinity: run where the first local input was produced (no cross-stream wait on the critical chain),
    // round-robin for tasks without inputs so independent chains spread over the streams
    si = -1;
    for (auto& i : task->inputs) {
      if (i.physical_data.device_id != device_id_) continue;
      auto it = data_.find(i.physical_data.data_id);
      if (it != data_.end() && it->second.stream >= 0) { si = it->second.stream; break; }
    }
    if (si < 0) si = rr_++ % kStreams;
    cudaStream_t st = streams_[si];
    DataList in, out;
    in.reserve(task->inputs.size());
    out.reserve(task->outputs.size());
    std::vector<Data*> touched;
    for (auto& i : task->inputs) {
      const PhysicalData& pd = i.physical_data;
      auto it = data_.find(pd.data_id);
      if (it == data_.end()) {
        if (pd.device_id == device_id_) throw std::runtime_error("input data is not on this device");
        if (!resolver_) throw std::runtime_error("remote input without a resolver");
        // pull a remote input (device.cpp:75-91): peer copy on this task's stream, after the producer's event
        const size_t bytes = static_cast<size_t>(pd.size.Prod()) * sizeof(float);
        Data d;
        d.block = Alloc(bytes, si);
        std::pair<float*, cudaEvent_t> src = resolver_(pd.device_id, pd.data_id);
        if (src.second) CudaOk(cudaStreamWaitEvent(st, src.second, 0), "cudaStreamWaitEvent(remote)");
        CudaOk(cudaMemcpyAsync(d.block.ptr, src.first, bytes, cudaMemcpyDefault, st), "cudaMemcpyAsync(peer)");
        d.stream = si;
        d.ready = NewEvent();
        CudaOk(cudaEventRecord(d.ready->ev, st), "cudaEventRecord");
        it = data_.emplace(pd.data_id, std::move(d)).first;
      } else if (it->second.stream != si && it->second.ready) {
        CudaOk(cudaStreamWaitEvent(st, it->second.ready->ev, 0), "cudaStreamWaitEvent");
        ++stats_.cross_stream_waits;
      }
      in.emplace_back(static_cast<float*>(it->second.block.ptr), pd.size);
      touched.push_back(&it->second);
    }
    for (auto& o : task->outputs) {
      const PhysicalData& pd = o.physical_data;
      if (data_.count(pd.data_id)) throw std::runtime_error("data already existed");
      Data d;
      d.block = Alloc(static_cast<size_t>(pd.size.Prod()) * sizeof(float), si);
      d.stream = si;
      auto it = data_.emplace(pd.data_id, std::move(d)).first;
      out.emplace_back(static_cast<float*>(it->second.block.ptr), pd.size);
    }
    // unordered_map rehash invalidates iterators but not element addresses; `touched` holds addresses taken before the
    // output insertions -- element addresses are stable in std::unordered_map
    Context ctx;
    ctx.impl_type = ImplType::kCuda;
    ctx.stream = st;
    ctx.workspace = Workspace(si);
    ctx.workspace_bytes = workspace_bytes_;
    if (!task->op.compute_fn) throw std::runtime_error("task without a compute function");
    task->op.compute_fn->Execute(in, out, ctx);          // enqueue-only (include/mnv.h)
    ev = NewEvent();
    CudaOk(cudaEventRecord(ev->ev, st), "cudaEventRecord");
    for (Data* d : touched) Touch(d->block, si, ev);
    for (auto& o : task->outputs) {
      Data& d = data_.at(o.physical_data.data_id);
      d.ready = ev;
      Touch(d.block, si, ev);
    }
    ++stats_.tasks;
  }
  switch (mode_) {
    case Completion::kBlocking:                          // device.cpp:221 then :118
      CudaOk(cudaStreamSynchronize(streams_[si]), task->op.compute_fn->Name().c_str());
      listener_->OnOperationComplete(task);
      break;
    case Completion::kEvent: {
      { std::lock_guard<std::mutex> l(cmu_); pending_.emplace_back(ev, task); }
      ccv_.notify_one();
      break;
    }
    case Completion::kEnqueue:
      listener_->OnOperationComplete(task);
      break;
  }
}

void StreamDevice::CompletionLoop() {
  cudaSetDevice(gpu_);
  for (;;) {
    std::pair<EvPtr, Task*> job;
    {
      std::unique_lock<std::mutex> l(cmu_);
      ccv_.wait(l, [&] { return stop_ || !pending_.empty(); });
      if (pending_.empty()) return;
      job = std::move(pending_.front());
      pending_.pop_front();
    }
    cudaEventSynchronize(job.first->ev);
    listener_->OnOperationComplete(job.second);
  }
}

void StreamDevice::FreeDataIfExist(uint64_t data_id) {
  std::lock_guard<std::mutex> lck(mu_);
  auto it = data_.find(data_id);
  if (it == data_.end()) return;
  free_[it->second.block.bytes].push_back(std::move(it->second.block));   // the block keeps its last-use events
  data_.erase(it);
}

float* StreamDevice::GetPtr(uint64_t data_id) {
  EvPtr ready;
  float* p;
  {
    std::lock_guard<std::mutex> lck(mu_);
    Data& d = data_.at(data_id);
    ready = d.ready;
    p = static_cast<float*>(d.block.ptr);
  }
  if (ready) CudaOk(cudaEventSynchronize(ready->ev), "cudaEventSynchronize");
  return p;
}

std::pair<float*, cudaEvent_t> StreamDevice::Export(uint64_t data_id) {
  std::lock_guard<std::mutex> lck(mu_);
  Data& d = data_.at(data_id);
  return std::make_pair(static_cast<float*>(d.block.ptr), d.ready ? d.ready->ev : nullptr);
}

void StreamDevice::CopyToHost(uint64_t data_id, float* dst, size_t floats) {
  float* src = GetPtr(data_id);
  CudaOk(cudaSetDevice(gpu_), "cudaSetDevice");
  CudaOk(cudaMemcpy(dst, src, floats * sizeof(float), cudaMemcpyDeviceToHost), "cudaMemcpy D2H");
}

void StreamDevice::WaitForAll() {
  CudaOk(cudaSetDevice(gpu_), "cudaSetDevice");
  for (int i = 0; i < kStreams; ++i) CudaOk(cudaStreamSynchronize(streams_[i]), "cudaStreamSynchronize");
  if (mode_ == Completion::kEvent) {
    for (;;) {
      { std::lock_guard<std::mutex> l(cmu_); if (pending_.empty()) break; }
      std::this_thread::yield();
    }
  }
}

StreamDevice::Stats StreamDevice::stats() const {
  std::lock_guard<std::mutex> lck(mu_);
  return stats_;
}

}  // namespace minerva

extern "C" size_t mnv_host_size_class(size_t bytes) { return minerva::StreamDevice::SizeClass(bytes); }
