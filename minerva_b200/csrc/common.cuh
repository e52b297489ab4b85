// common.cuh -- shared launch plumbing for the mnv_* C ABI (include/mnv.h).
// Replaces the reference's FindConfiguration (<=128 blocks of <=1024 threads, scalar 4-byte
// accesses; minerva/op/impl/cuda/cuda_perform.cu:12-30) with grids sized from the 148-SM B200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "mnv.h"

namespace mnv {

constexpr int kNumSMs = 148;           // B200: 2 dies x 74 SMs
constexpr int kBlock = 256;            // default CTA size for streaming kernels
constexpr int kBlocksPerSM = 8;        // 8 x 256 threads = 2048 resident threads per SM

extern std::atomic<uint64_t> g_launches;

inline int finish_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? MNV_OK : static_cast<int>(e);
}

inline cudaStream_t as_stream(mnv_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Grid for a grid-stride streaming kernel: enough CTAs to cover `work` items once, capped at one
// full wave of resident CTAs (a multiple of the SM count).
inline int stream_grid(size_t work_items, int per_block = kBlock) {
  size_t blocks = (work_items + per_block - 1) / per_block;
  size_t cap = static_cast<size_t>(kNumSMs) * kBlocksPerSM;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// Programmatic dependent launch: a kernel started through launch_pdl() may begin (its CTAs take the SM resources the
// previous kernel in the stream frees, run their prologue: barrier init, TMEM allocation, index math) before that kernel
// has finished; it must call pdl_wait() before its first global-memory access -- the wait returns once the previous
// grid has completed and its writes are visible, so the ordering the stream promises is unchanged.  pdl_enter() =
// "let MY successor start early too" + the wait.  A ~2 us launch gap per kernel boundary goes away (a step is
// 100..750 dependent launches), inside a recorded CUDA graph as well (programmatic edges).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }

extern std::atomic<int> g_pdl;     // 0: plain stream-ordered launches (mnv_set_dependent_launch)

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// the same with a thread-block cluster of `cluster_x` CTAs along x (grid.x a multiple of it)
template <typename... KArgs, typename... Args>
inline void launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = g_pdl.load(std::memory_order_relaxed) ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = cluster_x; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// max with the reference's predicate (keep `a` unless a < b), order-independent for non-NaN data
__device__ __forceinline__ float ref_max(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = ref_max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace mnv

#define MNV_CHECK_PTR(p) do { if ((p) == nullptr) return MNV_EINVAL; } while (0)
