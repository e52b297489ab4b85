#include "device/gpu_device.h"
#include <sstream>
#include <stdexcept>
#include "mnv.h"

namespace minerva {
namespace {
void CudaOk(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace

PooledDataStore::PooledDataStore(int gpu, size_t threshold_bytes) : gpu_(gpu), threshold_(threshold_bytes) {}

PooledDataStore::~PooledDataStore() {
  cudaSetDevice(gpu_);
  for (auto& kv : data_) cudaFree(kv.second.ptr);
  ReleaseFreeSpace();
}

float* PooledDataStore::CreateData(uint64_t id, size_t length) {
  std::lock_guard<std::mutex> lck(mu_);
  if (data_.count(id)) throw std::runtime_error("data already existed");
  DataState ds{nullptr, length};
  auto it = free_.find(length);
  if (it != free_.end() && !it->second.empty()) {   // reuse an exact-size block
    ds.ptr = it->second.back();
    it->second.pop_back();
  } else {
    if (threshold_ < total_ + length) ReleaseFreeSpace();
    CudaOk(cudaSetDevice(gpu_), "cudaSetDevice");
    CudaOk(cudaMalloc(&ds.ptr, length ? length : 4), "cudaMalloc");
    total_ += length;
  }
  data_.emplace(id, ds);
  return static_cast<float*>(ds.ptr);
}

float* PooledDataStore::GetData(uint64_t id) {
  std::lock_guard<std::mutex> lck(mu_);
  return static_cast<float*>(data_.at(id).ptr);
}

bool PooledDataStore::ExistData(uint64_t id) const {
  std::lock_guard<std::mutex> lck(mu_);
  return data_.count(id) != 0;
}

void PooledDataStore::FreeData(uint64_t id) {
  std::lock_guard<std::mutex> lck(mu_);
  auto it = data_.find(id);
  if (it == data_.end()) throw std::runtime_error("freeing unknown data");
  free_[it->second.length].push_back(it->second.ptr);
  data_.erase(it);
}

size_t PooledDataStore::GetTotalBytes() const {
  std::lock_guard<std::mutex> lck(mu_);
  return total_;
}

void PooledDataStore::ReleaseFreeSpace() {
  cudaSetDevice(gpu_);
  for (auto& kv : free_) {
    for (void* p : kv.second) { cudaFree(p); total_ -= kv.first; }
  }
  free_.clear();
}

GpuDevice::GpuDevice(int gpu_id, size_t pool_threshold) : gpu_(gpu_id), workspace_bytes_(mnv_workspace_bytes_hint()) {
  CudaOk(cudaSetDevice(gpu_), "cudaSetDevice");
  CudaOk(cudaFree(0), "context init");
  for (size_t i = 0; i < kParallelism; ++i) {
    CudaOk(cudaStreamCreateWithFlags(&streams_[i], cudaStreamNonBlocking), "cudaStreamCreate");
    CudaOk(cudaMalloc(&workspace_[i], workspace_bytes_), "workspace cudaMalloc");
  }
  store_ = new PooledDataStore(gpu_, pool_threshold);
}

GpuDevice::~GpuDevice() {
  cudaSetDevice(gpu_);
  for (size_t i = 0; i < kParallelism; ++i) {
    cudaStreamSynchronize(streams_[i]);
    cudaFree(workspace_[i]);
    cudaStreamDestroy(streams_[i]);
  }
  delete store_;
}

std::string GpuDevice::Name() const {
  std::ostringstream os;
  os << "GPU device #" << gpu_;
  return os.str();
}

void GpuDevice::DoExecute(const DataList& in, const DataList& out, PhysicalOp& op, int thrid, bool sync) {
  CudaOk(cudaSetDevice(gpu_), "cudaSetDevice");   // PreExecute (device.cpp:201-203)
  Context ctx;
  ctx.impl_type = ImplType::kCuda;
  ctx.stream = streams_[thrid];
  ctx.workspace = workspace_[thrid];
  ctx.workspace_bytes = workspace_bytes_;
  op.compute_fn->Execute(in, out, ctx);
  if (sync) CudaOk(cudaStreamSynchronize(streams_[thrid]), op.compute_fn->Name().c_str());
}

void GpuDevice::Barrier(int thrid) { CudaOk(cudaStreamSynchronize(streams_[thrid]), "Barrier"); }

void GpuDevice::DoCopyRemoteData(float* dst, float* src, size_t bytes, int thrid) {
  CudaOk(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, streams_[thrid]), "cudaMemcpyAsync");
  CudaOk(cudaStreamSynchronize(streams_[thrid]), "copy sync");
}

}  // namespace minerva
