// GpuDevice -- the device layer pieces north_star keeps in C++: per-worker CUDA streams, a pooled data store and
// (new) a per-stream kernel workspace handed to the ops through Context.  Reference: minerva/device/device.cpp:129-222
// (GpuDevice::Impl: kParallelism = 4 x {stream, cublas, cudnn}), minerva/device/pooled_data_store.cpp:19-65.
// The DAG scheduler / listener plumbing around it is out of scope; DoExecute() is the call the
// reference's ThreadedDevice::Execute makes for every task.
#pragma once
#include <cuda_runtime.h>
#include <array>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "op/hotpath.h"

namespace minerva {

// Exact-size free lists like the reference pool, but the free space is released only when an allocation
// would exceed the threshold (the reference flushes the whole pool past a 5 GiB total).
class PooledDataStore {
 public:
  PooledDataStore(int gpu, size_t threshold_bytes);
  ~PooledDataStore();
  float* CreateData(uint64_t id, size_t length_bytes);
  float* GetData(uint64_t id);
  bool ExistData(uint64_t id) const;
  void FreeData(uint64_t id);
  size_t GetTotalBytes() const;

 private:
  void ReleaseFreeSpace();
  struct DataState { void* ptr; size_t length; };
  const int gpu_;
  const size_t threshold_;
  size_t total_ = 0;
  mutable std::mutex mu_;
  std::unordered_map<uint64_t, DataState> data_;
  std::map<size_t, std::vector<void*>> free_;
};

class GpuDevice {
 public:
  static constexpr size_t kParallelism = 4;   // worker threads == streams, as in the reference
  explicit GpuDevice(int gpu_id, size_t pool_threshold = static_cast<size_t>(64) << 30);
  ~GpuDevice();
  GpuDevice(const GpuDevice&) = delete;
  GpuDevice& operator=(const GpuDevice&) = delete;

  std::string Name() const;
  PooledDataStore& data_store() { return *store_; }
  cudaStream_t stream(int thrid) const { return streams_[thrid]; }
  // One task: Context{kCuda, stream[thrid], workspace[thrid]} -> op.compute_fn->Execute -> stream sync
  // (the completion point of the reference, device.cpp:214-222).  Set `sync` false to keep the op asynchronous.
  void DoExecute(const DataList& in, const DataList& out, PhysicalOp& op, int thrid, bool sync = true);
  void Barrier(int thrid);
  void DoCopyRemoteData(float* dst, float* src, size_t bytes, int thrid);

 private:
  const int gpu_;
  std::array<cudaStream_t, kParallelism> streams_;
  std::array<void*, kParallelism> workspace_;
  size_t workspace_bytes_;
  PooledDataStore* store_;
};

}  // namespace minerva
