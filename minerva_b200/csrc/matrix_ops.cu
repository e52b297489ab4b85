// matrix_ops.cu -- a5 Reduction, a6 MaxIndex, a7 NormArithmetic, a19 Transpose, a23 Concat/Slice
// (strided copy), Select.  Column-major {m,n}: element (i,j) at i + j*m
// (reference layout rule: minerva/common/scale.cpp:182-191).
//
// Replaces K4-K9 (one thread per row/column looping serially, "TODO: this is inefficient",
// minerva/op/impl/cuda/cuda_kernel.h:107-198), cublasSgeam (L5) and the per-image cublasScopy
// loops of Concat/Slice (minerva/op/impl/cuda.cpp:80-155).
#include "common.cuh"

namespace mnv {

// ---- NormArithmetic ---------------------------------------------------------------------------
struct NAdd { __device__ float operator()(float x, float y) const { return __fadd_rn(x, y); } };
struct NSub { __device__ float operator()(float x, float y) const { return __fsub_rn(x, y); } };
struct NMul { __device__ float operator()(float x, float y) const { return __fmul_rn(x, y); } };
struct NDiv { __device__ float operator()(float x, float y) const { return __fdiv_rn(x, y); } };

// grid.y strides over columns, grid.x*block over rows: no integer division per element, coalesced
// along i.  ON_ROW: vec[i]; else vec[j].
template <bool ON_ROW, class Op>
__global__ void __launch_bounds__(kBlock) norm_kernel(const float* __restrict__ mat, const float* __restrict__ vec,
                                                      float* __restrict__ res, int m, int n, Op op) {
  for (int j = blockIdx.y; j < n; j += gridDim.y) {
    const float* src = mat + static_cast<size_t>(j) * m;
    float* dst = res + static_cast<size_t>(j) * m;
    float vj = ON_ROW ? 0.f : __ldg(vec + j);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
      dst[i] = op(__ldg(src + i), ON_ROW ? __ldg(vec + i) : vj);
  }
}

template <bool ON_ROW, class Op>
int launch_norm(const float* mat, const float* vec, float* res, int m, int n, cudaStream_t s) {
  if (m < 0 || n < 0) return MNV_EINVAL;
  if (m == 0 || n == 0) return MNV_OK;
  if (!mat || !vec || !res) return MNV_EINVAL;
  int gx = (m + kBlock - 1) / kBlock;
  if (gx > kNumSMs * kBlocksPerSM) gx = kNumSMs * kBlocksPerSM;
  int gy = (kNumSMs * kBlocksPerSM + gx - 1) / gx;
  if (gy > n) gy = n;
  if (gy > 65535) gy = 65535;
  norm_kernel<ON_ROW, Op><<<dim3(gx, gy), kBlock, 0, s>>>(mat, vec, res, m, n, Op{});
  return finish_launch();
}

// ---- Reduction / MaxIndex ---------------------------------------------------------------------
struct ValIdx { float v; int i; };
// reference predicate: replace only when strictly greater; equal values keep the smaller index
__device__ __forceinline__ ValIdx better(ValIdx a, ValIdx b) {
  if (a.v < b.v || (a.v == b.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ ValIdx warp_argmax(ValIdx x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  return x;
}

// MODE 0 sum, 1 max, 2 argmax.  One warp per column, lanes stride down the (contiguous) column.
template <int MODE>
__global__ void __launch_bounds__(kBlock) reduce_col_kernel(const float* __restrict__ in, float* __restrict__ out, int m, int n) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int j = warp; j < n; j += nwarps) {
    const float* col = in + static_cast<size_t>(j) * m;
    if (MODE == 0) {
      float acc = 0.f;
      for (int i = lane; i < m; i += 32) acc += __ldg(col + i);
      acc = warp_sum(acc);
      if (lane == 0) out[j] = acc;
    } else if (MODE == 1) {
      float acc = __ldg(col);
      for (int i = lane; i < m; i += 32) acc = ref_max(acc, __ldg(col + i));
      acc = warp_max(acc);
      if (lane == 0) out[j] = acc;
    } else {
      ValIdx b{__ldg(col), 0};
      for (int i = lane; i < m; i += 32) b = better(b, ValIdx{__ldg(col + i), i});
      b = warp_argmax(b);
      if (lane == 0) out[j] = static_cast<float>(b.i);
    }
  }
}

// Reduce across columns for each row: block = 32 rows x 8 column slices, coalesced along rows,
// slices combined through shared memory in slice order.
template <int MODE>
__global__ void __launch_bounds__(kBlock) reduce_row_kernel(const float* __restrict__ in, float* __restrict__ out, int m, int n) {
  __shared__ float sv[8][33];
  __shared__ int si[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i0 = blockIdx.x * 32; i0 < m; i0 += gridDim.x * 32) {
    int i = i0 + tx;
    float acc = 0.f;
    ValIdx b{0.f, 0};
    if (i < m) {
      if (MODE != 0) { acc = __ldg(in + i); b = ValIdx{acc, 0}; }
      for (int j = ty; j < n; j += 8) {
        float v = __ldg(in + i + static_cast<size_t>(j) * m);
        if (MODE == 0) acc += v;
        else if (MODE == 1) acc = ref_max(acc, v);
        else b = better(b, ValIdx{v, j});
      }
    }
    sv[ty][tx] = MODE == 2 ? b.v : acc;
    if (MODE == 2) si[ty][tx] = b.i;
    __syncthreads();
    if (ty == 0 && i < m) {
      if (MODE == 2) {
        ValIdx r{sv[0][tx], si[0][tx]};
        for (int t = 1; t < 8; ++t) r = better(r, ValIdx{sv[t][tx], si[t][tx]});
        out[i] = static_cast<float>(r.i);
      } else {
        float r = sv[0][tx];
        for (int t = 1; t < 8; ++t) r = MODE == 0 ? r + sv[t][tx] : ref_max(r, sv[t][tx]);
        out[i] = r;
      }
    }
    __syncthreads();
  }
}

template <int MODE>
int launch_reduce(bool on_row, const float* in, float* out, int m, int n, cudaStream_t s) {
  if (m <= 0 || n <= 0) return (m < 0 || n < 0) ? MNV_EINVAL : MNV_OK;
  if (!in || !out) return MNV_EINVAL;
  if (on_row) {
    int grid = (m + 31) / 32;
    if (grid > kNumSMs * kBlocksPerSM) grid = kNumSMs * kBlocksPerSM;
    reduce_row_kernel<MODE><<<grid, kBlock, 0, s>>>(in, out, m, n);
  } else {
    int grid = (n + 7) / 8;  // 8 warps per CTA
    if (grid > kNumSMs * kBlocksPerSM) grid = kNumSMs * kBlocksPerSM;
    reduce_col_kernel<MODE><<<grid, kBlock, 0, s>>>(in, out, m, n);
  }
  return finish_launch();
}

// ---- Transpose: c{n,m}[j + i*n] = a{m,n}[i + j*m], 32x32 tiles through padded smem --------------
__global__ void __launch_bounds__(kBlock) transpose_kernel(const float* __restrict__ a, float* __restrict__ c, int m, int n,
                                                           int tiles_m, int tiles_n) {
  __shared__ float tile[32][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int total = tiles_m * tiles_n;
  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    int ti = t % tiles_m, tj = t / tiles_m;
    int i0 = ti * 32, j0 = tj * 32;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int j = j0 + ty + r * 8, i = i0 + tx;
      if (i < m && j < n) tile[ty + r * 8][tx] = __ldg(a + i + static_cast<size_t>(j) * m);
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int i = i0 + ty + r * 8, j = j0 + tx;
      if (i < m && j < n) c[j + static_cast<size_t>(i) * n] = tile[tx][ty + r * 8];
    }
    __syncthreads();
  }
}

// ---- strided block copy (Concat / Slice) -----------------------------------------------------
__global__ void __launch_bounds__(kBlock) copy_strided_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t inner,
                                                              size_t outer, size_t src_stride, size_t dst_stride, int vec) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec) {
    size_t inner4 = inner / 4, total = inner4 * outer;
    for (size_t e = tid; e < total; e += stride) {
      size_t b = e / inner4, k = e - b * inner4;
      reinterpret_cast<float4*>(dst + b * dst_stride)[k] = __ldg(reinterpret_cast<const float4*>(src + b * src_stride) + k);
    }
  } else {
    size_t total = inner * outer;
    for (size_t e = tid; e < total; e += stride) {
      size_t b = e / inner, k = e - b * inner;
      dst[b * dst_stride + k] = __ldg(src + b * src_stride + k);
    }
  }
}

// up to 8 strided block copies that share `outer` in one launch (blockIdx.y = the copy): the four inputs of an inception
// concat, or the four slices of its gradient, are one kernel instead of four
struct CopySegs { const float* src[8]; float* dst[8]; size_t inner[8], src_stride[8], dst_stride[8]; int vec[8]; };
__global__ void __launch_bounds__(kBlock) copy_strided_n_kernel(const CopySegs sg, size_t outer) {
  const int j = blockIdx.y;
  const float* __restrict__ src = sg.src[j];
  float* __restrict__ dst = sg.dst[j];
  const size_t inner = sg.inner[j], ss = sg.src_stride[j], ds = sg.dst_stride[j];
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (sg.vec[j]) {
    const size_t inner4 = inner / 4, total = inner4 * outer;
    for (size_t e = tid; e < total; e += stride) {
      const size_t b = e / inner4, k = e - b * inner4;
      reinterpret_cast<float4*>(dst + b * ds)[k] = __ldg(reinterpret_cast<const float4*>(src + b * ss) + k);
    }
  } else {
    const size_t total = inner * outer;
    for (size_t e = tid; e < total; e += stride) {
      const size_t b = e / inner, k = e - b * inner;
      dst[b * ds + k] = __ldg(src + b * ss + k);
    }
  }
}

__global__ void __launch_bounds__(kBlock) select_kernel(float* __restrict__ dst, const float* __restrict__ src, const int* __restrict__ indices,
                                                        size_t n_idx, size_t rows) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t total = n_idx * rows;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    size_t j = e / rows, r = e - j * rows;
    dst[e] = __ldg(src + r + static_cast<size_t>(__ldg(indices + j)) * rows);
  }
}

}  // namespace mnv

using namespace mnv;

extern "C" {

#define MNV_NORM(name, on_row, Op)                                                                  \
  int name(const float* matrix, const float* vec, float* res, int m, int n, mnv_stream_t s) {      \
    return launch_norm<on_row, Op>(matrix, vec, res, m, n, as_stream(s));                           \
  }
MNV_NORM(mnv_norm_add_on_col, false, NAdd)
MNV_NORM(mnv_norm_sub_on_col, false, NSub)
MNV_NORM(mnv_norm_mult_on_col, false, NMul)
MNV_NORM(mnv_norm_div_on_col, false, NDiv)
MNV_NORM(mnv_norm_add_on_row, true, NAdd)
MNV_NORM(mnv_norm_sub_on_row, true, NSub)
MNV_NORM(mnv_norm_mult_on_row, true, NMul)
MNV_NORM(mnv_norm_div_on_row, true, NDiv)

int mnv_reduction_sum_on_col(const float* in, float* out, int m, int n, mnv_stream_t s) {
  return launch_reduce<0>(false, in, out, m, n, as_stream(s));
}
int mnv_reduction_max_on_col(const float* in, float* out, int m, int n, mnv_stream_t s) {
  return launch_reduce<1>(false, in, out, m, n, as_stream(s));
}
int mnv_reduction_sum_on_row(const float* in, float* out, int m, int n, mnv_stream_t s) {
  return launch_reduce<0>(true, in, out, m, n, as_stream(s));
}
int mnv_reduction_max_on_row(const float* in, float* out, int m, int n, mnv_stream_t s) {
  return launch_reduce<1>(true, in, out, m, n, as_stream(s));
}
int mnv_max_index_on_col(const float* in, float* out, int m, int n, mnv_stream_t s) {
  return launch_reduce<2>(false, in, out, m, n, as_stream(s));
}
int mnv_max_index_on_row(const float* in, float* out, int m, int n, mnv_stream_t s) {
  return launch_reduce<2>(true, in, out, m, n, as_stream(s));
}

int mnv_transpose(const float* a, float* c, int m, int n, mnv_stream_t s) {
  if (m < 0 || n < 0) return MNV_EINVAL;
  if (m == 0 || n == 0) return MNV_OK;
  if (!a || !c) return MNV_EINVAL;
  int tm = (m + 31) / 32, tn = (n + 31) / 32;
  long long total = static_cast<long long>(tm) * tn;
  int grid = static_cast<int>(total < kNumSMs * kBlocksPerSM ? total : kNumSMs * kBlocksPerSM);
  transpose_kernel<<<grid, kBlock, 0, as_stream(s)>>>(a, c, m, n, tm, tn);
  return finish_launch();
}

int mnv_copy_strided(const float* src, float* dst, size_t inner, size_t outer, size_t src_stride,
                     size_t dst_stride, mnv_stream_t s) {
  if (inner == 0 || outer == 0) return MNV_OK;
  if (!src || !dst) return MNV_EINVAL;
  int vec = aligned16(src) && aligned16(dst) && inner % 4 == 0 && src_stride % 4 == 0 && dst_stride % 4 == 0;
  size_t work = vec ? inner / 4 * outer : inner * outer;
  copy_strided_kernel<<<stream_grid(work), kBlock, 0, as_stream(s)>>>(src, dst, inner, outer, src_stride, dst_stride, vec);
  return finish_launch();
}

int mnv_copy_strided_n(const mnv_copy_seg_t* segs, int count, size_t outer, mnv_stream_t s) {
  if (count < 0 || count > 8) return MNV_EINVAL;
  if (count == 0 || outer == 0) return MNV_OK;
  if (!segs) return MNV_EINVAL;
  CopySegs sg;
  size_t max_work = 0;
  int n = 0;
  for (int j = 0; j < count; ++j) {
    const mnv_copy_seg_t& g = segs[j];
    if (g.inner == 0) continue;
    if (!g.src || !g.dst) return MNV_EINVAL;
    sg.src[n] = g.src; sg.dst[n] = g.dst; sg.inner[n] = g.inner; sg.src_stride[n] = g.src_stride; sg.dst_stride[n] = g.dst_stride;
    sg.vec[n] = aligned16(g.src) && aligned16(g.dst) && g.inner % 4 == 0 && g.src_stride % 4 == 0 && g.dst_stride % 4 == 0;
    const size_t work = sg.vec[n] ? g.inner / 4 * outer : g.inner * outer;
    if (work > max_work) max_work = work;
    ++n;
  }
  if (n == 0) return MNV_OK;
  int gx = stream_grid(max_work);
  if (gx * n > kNumSMs * kBlocksPerSM) gx = (kNumSMs * kBlocksPerSM + n - 1) / n;     // one wave over all the copies together
  copy_strided_n_kernel<<<dim3(gx, n), kBlock, 0, as_stream(s)>>>(sg, outer);
  return finish_launch();
}

int mnv_select(float* dst, const float* src, const int* indices, size_t n_idx, size_t cols, size_t rows,
               mnv_stream_t s) {
  (void)cols;
  if (n_idx == 0 || rows == 0) return MNV_OK;
  if (!dst || !src || !indices) return MNV_EINVAL;
  select_kernel<<<stream_grid(n_idx * rows), kBlock, 0, as_stream(s)>>>(dst, src, indices, n_idx, rows);
  return finish_launch();
}

}  // extern "C"
