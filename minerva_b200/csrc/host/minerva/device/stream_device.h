// StreamDevice -- the per-op host path of the device layer, redesigned for a GPU whose ops take microseconds
// (SURVEY 8f-1).  Reference: ThreadedDevice::Execute + GpuDevice::DoExecute (minerva/device/device.cpp:68-119,
// 214-222) and PooledDataStore (minerva/device/pooled_data_store.cpp:19-65).
//
// What the reference does per task: hop to one of 4 worker threads, cudaMalloc / exact-size free-list the outputs,
// launch, cudaStreamSynchronize (the host blocks for the whole op), hop back to a dispatcher thread through
// OnOperationComplete, which only then releases the successors.  With 5-100 us kernels the GPU idles between ops.
//
// Here PushTask is enqueue-only on the caller's thread:
//   * outputs come from a SIZE-CLASS pool whose blocks remember the events of their last uses, so a freed block is
//     handed to the next owner without any host synchronisation: the new owner's stream waits on those events
//     (stream-ordered reuse); nothing is ever cudaFree'd in steady state;
//   * every task records one event; a consumer on another stream (or another device, or the host through GetPtr /
//     CopyToHost) waits on the producer's event -- dependencies are enforced ON THE DEVICE, not by the dispatcher;
//   * completion is reported to the DeviceListener according to `Completion`:
//       kBlocking  after cudaStreamSynchronize, on the calling thread            (the reference's behaviour)
//       kEvent     by a completion thread when the task's event has fired         (same contract, host never blocks in PushTask)
//       kEnqueue   as soon as the task is enqueued                                (successors are released immediately; the
//                  listener contract "the result is readable" holds because every way of reading it -- a later task on any
//                  stream, a peer copy, GetPtr / CopyToHost -- is ordered after the producing event)
// The same Task / DeviceListener types as the reference, so the DAG scheduler drives it unchanged.
#pragma once
#include <cuda_runtime.h>
#include <array>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>
#include "device/task.h"

namespace minerva {

enum class Completion { kBlocking = 0, kEvent = 1, kEnqueue = 2 };

class StreamDevice {
 public:
  static constexpr int kStreams = 4;   // the reference's kParallelism (device.cpp:138)
  StreamDevice(uint64_t device_id, DeviceListener* listener, int gpu_id, Completion mode = Completion::kEnqueue);
  ~StreamDevice();
  StreamDevice(const StreamDevice&) = delete;
  StreamDevice& operator=(const StreamDevice&) = delete;

  void PushTask(Task* task);                       // device.h: Device::PushTask
  void FreeDataIfExist(uint64_t data_id);          // device.h: Device::FreeDataIfExist (stream-ordered, never blocks)
  float* GetPtr(uint64_t data_id);                 // blocks until the data is readable (NArray::Get's path)
  void CopyToHost(uint64_t data_id, float* dst, size_t floats);
  void WaitForAll();
  uint64_t device_id() const { return device_id_; }
  // resolves an input that lives on another device: -> (pointer, event after which it is readable)
  using RemoteResolver = std::function<std::pair<float*, cudaEvent_t>(uint64_t device_id, uint64_t data_id)>;
  void SetRemoteResolver(RemoteResolver r) { resolver_ = std::move(r); }
  std::pair<float*, cudaEvent_t> Export(uint64_t data_id);   // what a peer's resolver returns for data held here

  struct Stats { uint64_t tasks = 0, cuda_mallocs = 0, pool_hits = 0, cross_stream_waits = 0, bytes_reserved = 0; };
  Stats stats() const;
  static size_t SizeClass(size_t bytes);           // 256 B floor, then 8 classes per power of two (<= 12.5 % slack)

 private:
  struct Ev;                                        // pooled cudaEvent_t
  using EvPtr = std::shared_ptr<Ev>;
  struct Block { void* ptr = nullptr; size_t bytes = 0; std::array<EvPtr, kStreams> last_use; };
  struct Data { Block block; EvPtr ready; int stream = -1; };

  EvPtr NewEvent();
  Block Alloc(size_t bytes, int stream_idx);
  void Touch(Block& b, int stream_idx, const EvPtr& ev) { b.last_use[stream_idx] = ev; }
  void* Workspace(int stream_idx);
  void CompletionLoop();

  const uint64_t device_id_;
  DeviceListener* const listener_;
  const int gpu_;
  const Completion mode_;
  std::array<cudaStream_t, kStreams> streams_;
  std::array<void*, kStreams> workspace_{};
  size_t workspace_bytes_;
  mutable std::mutex mu_;
  std::unordered_map<uint64_t, Data> data_;
  std::map<size_t, std::vector<Block>> free_;      // size class -> blocks
  std::mutex ev_mu_;                               // guards event_pool_ only (an Ev may die while mu_ is held)
  std::vector<cudaEvent_t> event_pool_;
  int rr_ = 0;
  Stats stats_;
  RemoteResolver resolver_;
  // kEvent completion thread
  std::thread completer_;
  std::mutex cmu_;
  std::condition_variable ccv_;
  std::deque<std::pair<EvPtr, Task*>> pending_;
  bool stop_ = false;
};

}  // namespace minerva
