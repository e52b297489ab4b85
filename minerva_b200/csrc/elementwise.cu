// elementwise.cu -- a1 Arithmetic, a2/a3 ArithmeticConst, a4 Elewise, a10/a11 Activation,
// a20 copy, a21 Fill / Philox generators, 8(f) fused SGD.  All HBM-bound streaming kernels:
// 128-bit coalesced accesses, 4 independent vectors in flight per thread, grid = one wave of
// 148 SMs x 8 CTAs.  Replaces K1-K3, K10, K11 and the cuBLAS copy+axpy / copy+scal two-pass
// paths L2-L4, L6 and cuDNN activations L12.
//
// Bit-exactness vs minerva/op/impl/basic.cpp: every arithmetic step is an explicit IEEE
// round-to-nearest intrinsic (no FMA contraction, true division), so +,-,*,/,neg,relu,fill,copy
// match the CPU reference bit for bit.  exp/ln/tanh use CUDA's expf/logf/tanhf (<= 2 ulp from
// glibc; the reference's own tests allow 4 ulp, tests/unittest_elewise.cpp:17).
#include "common.cuh"

// ---- exact mode: glibc's expf / logf / tanhf and the reference's sigmoid, bit for bit -------------------------
// glibc_math.h is one text for host and device; here its arithmetic macros are the IEEE round-to-nearest intrinsics, which
// nvcc neither contracts nor reassociates.  The host build of the same text is compared with libm on all 2^32 inputs by
// tests/cpp/check_glibc_math.c (0 mismatches), so these kernels return what minerva/op/impl/basic.cpp:134,139,416,444 return.
#define MNV_GM_FN __device__ __forceinline__
#define MNV_GM_CONST __device__ const
#define MNV_INFF __int_as_float(0x7f800000)
#define MNV_NANF __int_as_float(0x7fffffff)
#define MNV_MUL(a, b) __dmul_rn((a), (b))
#define MNV_ADD(a, b) __dadd_rn((a), (b))
#define MNV_SUB(a, b) __dsub_rn((a), (b))
#define MNV_DIV(a, b) __ddiv_rn((a), (b))
#define MNV_FMA(a, b, c) __fma_rn((a), (b), (c))
#define MNV_FMULF(a, b) __fmul_rn((a), (b))
#define MNV_FADDF(a, b) __fadd_rn((a), (b))
#define MNV_FSUBF(a, b) __fsub_rn((a), (b))
#define MNV_FDIVF(a, b) __fdiv_rn((a), (b))
#include "glibc_math.h"

namespace mnv {

std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_pdl{1};

constexpr int kUnroll = 4;

template <int NIN, class Op>
__global__ void __launch_bounds__(kBlock) ew_vec4_kernel(const float4* __restrict__ a,
                                                         const float4* __restrict__ b,
                                                         const float4* __restrict__ c,
                                                         float4* __restrict__ out, size_t n4, Op op) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t base = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; base < n4;
       base += stride * kUnroll) {
    float4 va[kUnroll], vb[kUnroll], vc[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      size_t i = base + u * stride;
      if (i < n4) {
        va[u] = __ldg(a + i);
        if (NIN > 1) vb[u] = __ldg(b + i);
        if (NIN > 2) vc[u] = __ldg(c + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      size_t i = base + u * stride;
      if (i < n4) {
        float4 r;
        r.x = op(va[u].x, vb[u].x, vc[u].x);
        r.y = op(va[u].y, vb[u].y, vc[u].y);
        r.z = op(va[u].z, vb[u].z, vc[u].z);
        r.w = op(va[u].w, vb[u].w, vc[u].w);
        out[i] = r;
      }
    }
  }
}

template <int NIN, class Op>
__global__ void __launch_bounds__(kBlock) ew_scalar_kernel(const float* __restrict__ a,
                                                           const float* __restrict__ b,
                                                           const float* __restrict__ c,
                                                           float* __restrict__ out, size_t begin,
                                                           size_t n, Op op) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = begin + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float x = a[i];
    float y = NIN > 1 ? b[i] : 0.f;
    float z = NIN > 2 ? c[i] : 0.f;
    out[i] = op(x, y, z);
  }
}

// One entry for every elementwise op: vector body + scalar tail (or all-scalar when a pointer is
// not 16-byte aligned, e.g. a sliced view).
template <int NIN, class Op>
int launch_ew(const float* a, const float* b, const float* c, float* out, size_t n, Op op, cudaStream_t s) {
  if (n == 0) return MNV_OK;
  if (!a || !out || (NIN > 1 && !b) || (NIN > 2 && !c)) return MNV_EINVAL;
  bool vec = aligned16(a) && aligned16(out) && (NIN < 2 || aligned16(b)) && (NIN < 3 || aligned16(c));
  size_t n4 = vec ? n / 4 : 0;
  if (n4) {
    int grid = stream_grid((n4 + kUnroll - 1) / kUnroll);
    ew_vec4_kernel<NIN, Op><<<grid, kBlock, 0, s>>>(reinterpret_cast<const float4*>(a),
                                                     reinterpret_cast<const float4*>(b),
                                                     reinterpret_cast<const float4*>(c),
                                                     reinterpret_cast<float4*>(out), n4, op);
    int rc = finish_launch();
    if (rc) return rc;
  }
  if (n4 * 4 < n) {
    size_t rest = n - n4 * 4;
    ew_scalar_kernel<NIN, Op><<<stream_grid(rest), kBlock, 0, s>>>(a, b, c, out, n4 * 4, n, op);
    return finish_launch();
  }
  return MNV_OK;
}

// dst = ((src0 + src1) + src2) + ...: the left-to-right chain of mnv_add calls in one pass (count + 1 streams instead of
// 3 * (count - 1)).  owl.net sums the sensitivities a blob receives from its consumers with it (an inception input has four).
struct AddNArgs { const float* src[8]; int count; };
template <bool VEC>
__global__ void __launch_bounds__(kBlock) add_n_kernel(const AddNArgs a, float* out, size_t n) {   // out may be src[0] (plain loads, no restrict)
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  if (VEC) {
    const size_t n4 = n / 4;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 acc = reinterpret_cast<const float4*>(a.src[0])[i];
#pragma unroll 8
      for (int j = 1; j < a.count; ++j) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.src[j]) + i);
        acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
      }
      reinterpret_cast<float4*>(out)[i] = acc;
    }
    for (size_t i = n4 * 4 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
      float acc = a.src[0][i];
      for (int j = 1; j < a.count; ++j) acc = __fadd_rn(acc, a.src[j][i]);
      out[i] = acc;
    }
  } else {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
      float acc = a.src[0][i];
      for (int j = 1; j < a.count; ++j) acc = __fadd_rn(acc, a.src[j][i]);
      out[i] = acc;
    }
  }
}

// ---- functors (x, y, z) -> result; unused operands are ignored ------------------------------
struct AddOp { __device__ float operator()(float x, float y, float) const { return __fadd_rn(x, y); } };
struct SubOp { __device__ float operator()(float x, float y, float) const { return __fsub_rn(x, y); } };
struct MulOp { __device__ float operator()(float x, float y, float) const { return __fmul_rn(x, y); } };
struct DivOp { __device__ float operator()(float x, float y, float) const { return __fdiv_rn(x, y); } };
struct ConstAddOp { float v; __device__ float operator()(float x, float, float) const { return __fadd_rn(x, v); } };
struct LeftConstSubOp { float v; __device__ float operator()(float x, float, float) const { return __fsub_rn(v, x); } };
struct LeftConstDivOp { float v; __device__ float operator()(float x, float, float) const { return __fdiv_rn(v, x); } };
struct ScaleOp { float v; __device__ float operator()(float x, float, float) const { return __fmul_rn(x, v); } };
struct ConstDivOp { float v; __device__ float operator()(float x, float, float) const { return __fdiv_rn(x, v); } };
struct ExpOp { __device__ float operator()(float x, float, float) const { return expf(x); } };
struct LnOp { __device__ float operator()(float x, float, float) const { return logf(x); } };
struct NegOp { __device__ float operator()(float x, float, float) const { return -x; } };
struct CopyOp { __device__ float operator()(float x, float, float) const { return x; } };
// basic.cpp:416 -- float expf, then the reciprocal in double, rounded once to float
struct SigmoidOp {
  __device__ float operator()(float x, float, float) const {
    return static_cast<float>(1.0 / (1.0 + static_cast<double>(expf(-x))));
  }
};
struct ExpExactOp { __device__ float operator()(float x, float, float) const { return mnv_glibc_expf(x); } };
struct LnExactOp { __device__ float operator()(float x, float, float) const { return mnv_glibc_logf(x); } };
struct TanhExactOp { __device__ float operator()(float x, float, float) const { return mnv_glibc_tanhf(x); } };
struct SigmoidExactOp { __device__ float operator()(float x, float, float) const { return mnv_ref_sigmoidf(x); } };
struct ReluOp { __device__ float operator()(float x, float, float) const { return x > 0.f ? x : 0.f; } };  // basic.cpp:430
struct TanhOp { __device__ float operator()(float x, float, float) const { return tanhf(x); } };
// backward functors take (dy, y, x)
struct SigmoidBackOp {
  __device__ float operator()(float dy, float y, float) const {
    return __fmul_rn(__fmul_rn(dy, y), __fsub_rn(1.0f, y));
  }
};
struct ReluBackOp { __device__ float operator()(float dy, float x, float) const { return x > 0.f ? dy : 0.f; } };
struct TanhBackOp {
  __device__ float operator()(float dy, float y, float) const {
    return __fmul_rn(dy, __fsub_rn(1.0f, __fmul_rn(y, y)));
  }
};

// ---- fill -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) fill_kernel(float* __restrict__ dst, size_t n, float v) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // dst is 4-byte aligned at least; peel to a 16-byte boundary, then 128-bit stores
  size_t head = ((16 - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u) / 4;
  if (head > n) head = n;
  if (tid < head) dst[tid] = v;
  float4* d4 = reinterpret_cast<float4*>(dst + head);
  size_t n4 = (n - head) / 4;
  float4 vv = make_float4(v, v, v, v);
  for (size_t i = tid; i < n4; i += stride) d4[i] = vv;
  size_t done = head + n4 * 4;
  if (tid < n - done) dst[done + tid] = v;
}

// ---- Philox4x32-10 (same stream layout as oracle/mnv_oracle.c:orc_philox4x32) -------------------
__device__ __forceinline__ uint4 philox4x32(uint32_t seed, uint64_t block, uint32_t stream_id) {
  uint32_t c0 = static_cast<uint32_t>(block), c1 = static_cast<uint32_t>(block >> 32), c2 = stream_id, c3 = 0u;
  uint32_t k0 = seed, k1 = 0x0B200B20u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ void store4_guarded(float* dst, size_t blk, size_t n, const float v[4]) {
  size_t i = blk * 4;
  if (i + 4 <= n && (reinterpret_cast<uintptr_t>(dst + i) & 15u) == 0) {
    *reinterpret_cast<float4*>(dst + i) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int j = 0; j < 4 && i + j < n; ++j) dst[i + j] = v[j];
  }
}

// seed_base != null (mnv_rand_bernoulli_ds): the stream's key is (*seed_base + seed) ^ seed_xor, read on the device -- a launch
// recorded in a CUDA graph draws a new mask on every replay once the caller has stored the step's base word.
__global__ void __launch_bounds__(kBlock) bernoulli_kernel(float* __restrict__ dst, size_t n, uint32_t seed, float p,
                                                           const uint32_t* __restrict__ seed_base, uint32_t seed_xor) {
  if (seed_base) seed = (__ldg(seed_base) + seed) ^ seed_xor;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t nblk = (n + 3) / 4;
  for (size_t blk = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; blk < nblk; blk += stride) {
    uint4 r = philox4x32(seed, blk, 1u);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float u = __fmul_rn(static_cast<float>(w[j] >> 8), 1.0f / 16777216.0f);
      v[j] = u < p ? 1.0f : 0.0f;
    }
    store4_guarded(dst, blk, n, v);
  }
}

__global__ void __launch_bounds__(kBlock) randn_kernel(float* __restrict__ dst, size_t n, uint32_t seed, float mean, float sd) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t nblk = (n + 3) / 4;
  for (size_t blk = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; blk < nblk; blk += stride) {
    uint4 r = philox4x32(seed, blk, 2u);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    float v[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float u1 = __fmul_rn(static_cast<float>((w[2 * h] >> 8) + 1u), 1.0f / 16777216.0f);
      float u2 = __fmul_rn(static_cast<float>(w[2 * h + 1] >> 8), 1.0f / 16777216.0f);
      float rad = sqrtf(__fmul_rn(-2.0f, logf(u1)));
      float ang = __fmul_rn(6.283185307179586f, u2);
      v[2 * h] = __fadd_rn(mean, __fmul_rn(sd, __fmul_rn(rad, cosf(ang))));
      v[2 * h + 1] = __fadd_rn(mean, __fmul_rn(sd, __fmul_rn(rad, sinf(ang))));
    }
    store4_guarded(dst, blk, n, v);
  }
}

// ---- in-place binary ops: acc[i] = op(acc[i], x[i]) -------------------------------------------------------------
// The reference-shaped entries promise "outputs never alias inputs" and read through the non-coherent path
// (__ldg, __restrict__); these two are the explicit in-place forms the gradient merge (acc += peer shard) and the
// pooling-backward ReLU fallback need: the aliased pointer is read with plain loads and is not __restrict__.
template <class Op>
__global__ void __launch_bounds__(kBlock) ew_inplace_kernel(float* acc, const float* __restrict__ x, size_t n, int vec, Op op) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t n4 = vec ? n / 4 : 0;
  float4* a4 = reinterpret_cast<float4*>(acc);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (size_t base = tid; base < n4; base += stride * kUnroll) {
    float4 va[kUnroll], vx[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t i = base + u * stride;
      if (i < n4) { va[u] = a4[i]; vx[u] = __ldg(x4 + i); }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t i = base + u * stride;
      if (i < n4) a4[i] = make_float4(op(va[u].x, vx[u].x, 0.f), op(va[u].y, vx[u].y, 0.f), op(va[u].z, vx[u].z, 0.f), op(va[u].w, vx[u].w, 0.f));
    }
  }
  for (size_t i = n4 * 4 + tid; i < n; i += stride) acc[i] = op(acc[i], x[i], 0.f);
}
template <class Op>
static int launch_inplace(float* acc, const float* x, size_t n, Op op, cudaStream_t s) {
  if (n == 0) return MNV_OK;
  if (!acc || !x) return MNV_EINVAL;
  const int vec = aligned16(acc) && aligned16(x);
  ew_inplace_kernel<Op><<<stream_grid(n / (4 * kUnroll) + 1), kBlock, 0, s>>>(acc, x, n, vec, op);
  return finish_launch();
}

// ---- data-layer transform: uint8 stored images -> mean-subtracted, cropped, mirrored fp32 batch -----------------------
// One thread per 4 consecutive output pixels of a row: 4 byte loads (or one 32-bit load when the source is aligned and
// not mirrored), 4 mean loads, one 16-byte store.  HBM-bound: 1 + 4 (mean, L2-resident: one image) + 4 B per element.
__global__ void __launch_bounds__(kBlock) image_transform_u8_kernel(const unsigned char* __restrict__ src, const float* __restrict__ mean,
                                                                    const int* __restrict__ off, float* __restrict__ dst, int C, int sh, int sw,
                                                                    int ch, int cw, float scale, size_t groups, int gpr) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t g = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const int xg = static_cast<int>(g % gpr);
    size_t r = g / gpr;
    const int y = static_cast<int>(r % ch);
    r /= ch;
    const int c = static_cast<int>(r % C);
    const size_t n = r / C;
    int oy = 0, ox = 0, mir = 0;
    if (off) { oy = __ldg(off + 3 * n); ox = __ldg(off + 3 * n + 1); mir = __ldg(off + 3 * n + 2); }
    const size_t srow = ((n * C + c) * sh + oy + y) * static_cast<size_t>(sw) + ox;
    const size_t mrow = (static_cast<size_t>(c) * sh + oy + y) * static_cast<size_t>(sw) + ox;
    float* d = dst + ((n * C + c) * ch + y) * static_cast<size_t>(cw) + 4 * xg;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = 4 * xg + j;
      if (x < cw) {
        const int xs = mir ? cw - 1 - x : x;
        const float px = static_cast<float>(__ldg(src + srow + xs));
        v[j] = __fmul_rn(mean ? __fsub_rn(px, __ldg(mean + mrow + xs)) : px, scale);
      } else {
        v[j] = 0.f;
      }
    }
    if (4 * xg + 4 <= cw && (reinterpret_cast<uintptr_t>(d) & 15u) == 0) {
      *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int j = 0; j < 4 && 4 * xg + j < cw; ++j) d[j] = v[j];
    }
  }
}

// ---- fused momentum SGD (in place): 3 reads + 2 writes = 20 B/param ------------------------------
__global__ void __launch_bounds__(kBlock) sgd_kernel(float* __restrict__ w, float* __restrict__ delta,
                                                     const float* __restrict__ grad, size_t n, float mom,
                                                     float lrb, float lrwd, int vec) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  auto upd = [&](float& wv, float& dv, float g) {
    float d = __fsub_rn(__fsub_rn(__fmul_rn(mom, dv), __fmul_rn(lrb, g)), __fmul_rn(lrwd, wv));
    dv = d;
    wv = __fadd_rn(wv, d);
  };
  size_t n4 = vec ? n / 4 : 0;
  float4* w4 = reinterpret_cast<float4*>(w);
  float4* d4 = reinterpret_cast<float4*>(delta);
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  for (size_t i = tid; i < n4; i += stride) {
    float4 wv = w4[i], dv = d4[i], gv = __ldg(g4 + i);
    upd(wv.x, dv.x, gv.x); upd(wv.y, dv.y, gv.y); upd(wv.z, dv.z, gv.z); upd(wv.w, dv.w, gv.w);
    w4[i] = wv; d4[i] = dv;
  }
  for (size_t i = n4 * 4 + tid; i < n; i += stride) {
    float wv = w[i], dv = delta[i];
    upd(wv, dv, grad[i]);
    w[i] = wv; delta[i] = dv;
  }
}

// The same update for up to kSgdMaxTensors tensors in ONE launch: a step's 16 (AlexNet) / 128 (GoogLeNet) parameter
// tensors are mostly small (biases, 1x1 filters), so one launch per tensor is launch-bound (16 launches: 0.29 ms for
// 1.25 GB = 65 % of the HBM peak; one launch streams at the rate of the large tensors).  Work is cut into chunks of
// kSgdChunk elements; a CTA walks chunks grid-stride and finds each chunk's tensor in the (<= 64 entry) prefix table.
constexpr int kSgdMaxTensors = 64;
constexpr int kSgdChunk = 4096;
struct SgdTable {
  float* w[kSgdMaxTensors];
  float* delta[kSgdMaxTensors];
  const float* grad[kSgdMaxTensors];
  unsigned long long n[kSgdMaxTensors];
  unsigned int first_chunk[kSgdMaxTensors + 1];
  float lrb[kSgdMaxTensors], lrwd[kSgdMaxTensors];
  int count;
};
__global__ void __launch_bounds__(kBlock) sgd_multi_kernel(const __grid_constant__ SgdTable t, float mom) {
  const unsigned total = t.first_chunk[t.count];
  for (unsigned c = blockIdx.x; c < total; c += gridDim.x) {
    int lo = 0, hi = t.count - 1;
    while (lo < hi) {                       // last tensor whose first chunk is <= c
      const int mid = (lo + hi + 1) >> 1;
      if (t.first_chunk[mid] <= c) lo = mid; else hi = mid - 1;
    }
    float* w = t.w[lo];
    float* d = t.delta[lo];
    const float* g = t.grad[lo];
    const float lrb = t.lrb[lo], lrwd = t.lrwd[lo];
    const size_t begin = static_cast<size_t>(c - t.first_chunk[lo]) * kSgdChunk;
    const size_t end = min(begin + static_cast<size_t>(kSgdChunk), static_cast<size_t>(t.n[lo]));
    auto upd = [&](float& wv, float& dv, float gv) {
      const float nd = __fsub_rn(__fsub_rn(__fmul_rn(mom, dv), __fmul_rn(lrb, gv)), __fmul_rn(lrwd, wv));
      dv = nd;
      wv = __fadd_rn(wv, nd);
    };
    const bool vec = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(g)) & 15u) == 0;
    if (vec) {                              // chunk starts are multiples of 4 elements
      const size_t b4 = begin / 4, e4 = end / 4;
      float4* w4 = reinterpret_cast<float4*>(w);
      float4* d4 = reinterpret_cast<float4*>(d);
      const float4* g4 = reinterpret_cast<const float4*>(g);
      for (size_t i = b4 + threadIdx.x; i < e4; i += kBlock) {
        float4 wv = w4[i], dv = d4[i], gv = __ldg(g4 + i);
        upd(wv.x, dv.x, gv.x); upd(wv.y, dv.y, gv.y); upd(wv.z, dv.z, gv.z); upd(wv.w, dv.w, gv.w);
        w4[i] = wv; d4[i] = dv;
      }
      for (size_t i = e4 * 4 + threadIdx.x; i < end; i += kBlock) { float wv = w[i], dv = d[i]; upd(wv, dv, g[i]); w[i] = wv; d[i] = dv; }
    } else {
      for (size_t i = begin + threadIdx.x; i < end; i += kBlock) { float wv = w[i], dv = d[i]; upd(wv, dv, g[i]); w[i] = wv; d[i] = dv; }
    }
  }
}

}  // namespace mnv

using namespace mnv;

extern "C" {

int mnv_add(const float* a, const float* b, float* c, size_t n, mnv_stream_t s) {
  return launch_ew<2>(a, b, nullptr, c, n, AddOp{}, as_stream(s));
}
int mnv_add_n(const float* const* srcs, int count, float* dst, size_t n, mnv_stream_t s) {
  if (count < 1 || count > 8) return MNV_EINVAL;
  if (n == 0) return MNV_OK;
  if (!srcs || !dst) return MNV_EINVAL;
  AddNArgs a;
  bool vec = aligned16(dst);
  for (int j = 0; j < 8; ++j) a.src[j] = nullptr;
  for (int j = 0; j < count; ++j) {
    if (!srcs[j]) return MNV_EINVAL;
    a.src[j] = srcs[j];
    vec = vec && aligned16(srcs[j]);
  }
  a.count = count;
  if (vec) add_n_kernel<true><<<stream_grid(n / 4 + 1), kBlock, 0, as_stream(s)>>>(a, dst, n);
  else add_n_kernel<false><<<stream_grid(n), kBlock, 0, as_stream(s)>>>(a, dst, n);
  return finish_launch();
}
int mnv_sub(const float* a, const float* b, float* c, size_t n, mnv_stream_t s) {
  return launch_ew<2>(a, b, nullptr, c, n, SubOp{}, as_stream(s));
}
int mnv_dot_mult(const float* a, const float* b, float* c, size_t n, mnv_stream_t s) {
  return launch_ew<2>(a, b, nullptr, c, n, MulOp{}, as_stream(s));
}
int mnv_dot_div(const float* a, const float* b, float* c, size_t n, mnv_stream_t s) {
  return launch_ew<2>(a, b, nullptr, c, n, DivOp{}, as_stream(s));
}
int mnv_const_add(const float* in, float* out, float v, size_t n, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, ConstAddOp{v}, as_stream(s));
}
int mnv_left_const_sub(const float* in, float* out, float v, size_t n, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, LeftConstSubOp{v}, as_stream(s));
}
int mnv_left_const_div(const float* in, float* out, float v, size_t n, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, LeftConstDivOp{v}, as_stream(s));
}
int mnv_scale(const float* in, float* out, size_t n, float v, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, ScaleOp{v}, as_stream(s));
}
int mnv_const_div(const float* in, float* out, float v, size_t n, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, ConstDivOp{v}, as_stream(s));
}
int mnv_elewise_exp(const float* in, float* out, size_t n, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, ExpOp{}, as_stream(s));
}
int mnv_elewise_ln(const float* in, float* out, size_t n, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, LnOp{}, as_stream(s));
}
int mnv_elewise_exp_exact(const float* in, float* out, size_t n, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, ExpExactOp{}, as_stream(s));
}
int mnv_elewise_ln_exact(const float* in, float* out, size_t n, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, LnExactOp{}, as_stream(s));
}
int mnv_elewise_negative(const float* in, float* out, size_t n, mnv_stream_t s) {
  return launch_ew<1>(in, nullptr, nullptr, out, n, NegOp{}, as_stream(s));
}
int mnv_copy(const float* src, float* dst, size_t n, mnv_stream_t s) {
  return launch_ew<1>(src, nullptr, nullptr, dst, n, CopyOp{}, as_stream(s));
}
int mnv_reshape(const float* in, float* out, size_t bytes, mnv_stream_t s) {
  if (bytes % sizeof(float)) return MNV_EINVAL;
  return launch_ew<1>(in, nullptr, nullptr, out, bytes / sizeof(float), CopyOp{}, as_stream(s));
}

static inline size_t prod4(int a, int b, int c, int d) {
  return static_cast<size_t>(a) * b * c * static_cast<size_t>(d);
}
#define MNV_DIMS_OK(a, b, c, d) ((a) >= 0 && (b) >= 0 && (c) >= 0 && (d) >= 0)

int mnv_sigmoid_forward(const float* x, float* y, int N, int C, int H, int W, mnv_stream_t s) {
  if (!MNV_DIMS_OK(N, C, H, W)) return MNV_EINVAL;
  return launch_ew<1>(x, nullptr, nullptr, y, prod4(N, C, H, W), SigmoidOp{}, as_stream(s));
}
int mnv_relu_forward(const float* x, float* y, int N, int C, int H, int W, mnv_stream_t s) {
  if (!MNV_DIMS_OK(N, C, H, W)) return MNV_EINVAL;
  return launch_ew<1>(x, nullptr, nullptr, y, prod4(N, C, H, W), ReluOp{}, as_stream(s));
}
int mnv_tanh_forward(const float* x, float* y, int N, int C, int H, int W, mnv_stream_t s) {
  if (!MNV_DIMS_OK(N, C, H, W)) return MNV_EINVAL;
  return launch_ew<1>(x, nullptr, nullptr, y, prod4(N, C, H, W), TanhOp{}, as_stream(s));
}
int mnv_sigmoid_forward_exact(const float* x, float* y, int N, int C, int H, int W, mnv_stream_t s) {
  if (!MNV_DIMS_OK(N, C, H, W)) return MNV_EINVAL;
  return launch_ew<1>(x, nullptr, nullptr, y, prod4(N, C, H, W), SigmoidExactOp{}, as_stream(s));
}
int mnv_tanh_forward_exact(const float* x, float* y, int N, int C, int H, int W, mnv_stream_t s) {
  if (!MNV_DIMS_OK(N, C, H, W)) return MNV_EINVAL;
  return launch_ew<1>(x, nullptr, nullptr, y, prod4(N, C, H, W), TanhExactOp{}, as_stream(s));
}
// Only the operands the formula needs are read (12 B/elem): sigmoid/tanh use (dy, y), relu (dy, x).
int mnv_sigmoid_backward(const float* x, const float* y, const float* dy, float* dx, int N, int C, int H,
                         int W, mnv_stream_t s) {
  (void)x;
  if (!MNV_DIMS_OK(N, C, H, W)) return MNV_EINVAL;
  return launch_ew<2>(dy, y, nullptr, dx, prod4(N, C, H, W), SigmoidBackOp{}, as_stream(s));
}
int mnv_relu_backward(const float* x, const float* y, const float* dy, float* dx, int N, int C, int H,
                      int W, mnv_stream_t s) {
  (void)y;
  if (!MNV_DIMS_OK(N, C, H, W)) return MNV_EINVAL;
  return launch_ew<2>(dy, x, nullptr, dx, prod4(N, C, H, W), ReluBackOp{}, as_stream(s));
}
int mnv_tanh_backward(const float* x, const float* y, const float* dy, float* dx, int N, int C, int H,
                      int W, mnv_stream_t s) {
  (void)x;
  if (!MNV_DIMS_OK(N, C, H, W)) return MNV_EINVAL;
  return launch_ew<2>(dy, y, nullptr, dx, prod4(N, C, H, W), TanhBackOp{}, as_stream(s));
}

int mnv_fill(float* dst, size_t n, float v, mnv_stream_t s) {
  if (n == 0) return MNV_OK;
  MNV_CHECK_PTR(dst);
  fill_kernel<<<stream_grid(n / 4 + 8), kBlock, 0, as_stream(s)>>>(dst, n, v);
  return finish_launch();
}
int mnv_rand_bernoulli(float* dst, size_t n, unsigned int seed, float p, mnv_stream_t s) {
  if (n == 0) return MNV_OK;
  MNV_CHECK_PTR(dst);
  bernoulli_kernel<<<stream_grid((n + 3) / 4), kBlock, 0, as_stream(s)>>>(dst, n, seed, p, nullptr, 0u);
  return finish_launch();
}
int mnv_rand_bernoulli_ds(float* dst, size_t n, const unsigned int* seed_base, unsigned int seed_add, unsigned int seed_xor, float p,
                          mnv_stream_t s) {
  if (n == 0) return MNV_OK;
  MNV_CHECK_PTR(dst);
  MNV_CHECK_PTR(seed_base);
  bernoulli_kernel<<<stream_grid((n + 3) / 4), kBlock, 0, as_stream(s)>>>(dst, n, seed_add, p, seed_base, seed_xor);
  return finish_launch();
}
int mnv_randn(float* dst, size_t n, unsigned int seed, float mean, float var, mnv_stream_t s) {
  if (n == 0) return MNV_OK;
  MNV_CHECK_PTR(dst);
  randn_kernel<<<stream_grid((n + 3) / 4), kBlock, 0, as_stream(s)>>>(dst, n, seed, mean, var);
  return finish_launch();
}
int mnv_sgd_momentum_update(float* w, float* delta, const float* grad, size_t n, float momentum,
                            float lr_over_batch, float lr_times_wd, mnv_stream_t s) {
  if (n == 0) return MNV_OK;
  if (!w || !delta || !grad) return MNV_EINVAL;
  int vec = aligned16(w) && aligned16(delta) && aligned16(grad);
  sgd_kernel<<<stream_grid(n / 4 + 1), kBlock, 0, as_stream(s)>>>(w, delta, grad, n, momentum,
                                                                    lr_over_batch, lr_times_wd, vec);
  return finish_launch();
}

int mnv_sgd_momentum_update_multi(const mnv_sgd_tensor_t* tensors, int count, float momentum, mnv_stream_t s) {
  if (count < 0 || (count > 0 && !tensors)) return MNV_EINVAL;
  for (int i = 0; i < count;) {
    SgdTable t;
    t.count = 0;
    unsigned chunks = 0;
    for (; i < count && t.count < kSgdMaxTensors; ++i) {
      const mnv_sgd_tensor_t& e = tensors[i];
      if (e.n == 0) continue;
      if (!e.w || !e.delta || !e.grad) return MNV_EINVAL;
      const size_t nch = (e.n + kSgdChunk - 1) / kSgdChunk;
      if (nch > 0x7fffffffu - chunks) return MNV_EUNSUPPORTED;
      const int k = t.count++;
      t.w[k] = e.w; t.delta[k] = e.delta; t.grad[k] = e.grad; t.n[k] = e.n; t.lrb[k] = e.lr_over_batch; t.lrwd[k] = e.lr_times_wd;
      t.first_chunk[k] = chunks;
      chunks += static_cast<unsigned>(nch);
    }
    if (t.count == 0) continue;
    t.first_chunk[t.count] = chunks;
    sgd_multi_kernel<<<stream_grid(static_cast<size_t>(chunks) * kBlock), kBlock, 0, as_stream(s)>>>(t, momentum);
    int rc = finish_launch();
    if (rc) return rc;
  }
  return MNV_OK;
}
int mnv_accumulate(float* acc, const float* x, size_t n, mnv_stream_t s) {
  return launch_inplace(acc, x, n, AddOp{}, as_stream(s));
}
int mnv_relu_mask_inplace(float* dx, const float* x, size_t n, mnv_stream_t s) {
  return launch_inplace(dx, x, n, ReluBackOp{}, as_stream(s));
}

int mnv_image_transform_u8(const unsigned char* src, const float* mean, const int* crop_mirror, float* dst, int N, int C,
                           int src_h, int src_w, int crop_h, int crop_w, float scale, mnv_stream_t s) {
  if (N < 0 || C <= 0 || src_h <= 0 || src_w <= 0 || crop_h <= 0 || crop_w <= 0 || crop_h > src_h || crop_w > src_w) return MNV_EINVAL;
  if (N == 0) return MNV_OK;
  if (!src || !dst) return MNV_EINVAL;
  const int gpr = (crop_w + 3) / 4;
  const size_t groups = static_cast<size_t>(N) * C * crop_h * gpr;
  image_transform_u8_kernel<<<stream_grid(groups), kBlock, 0, as_stream(s)>>>(src, mean, crop_mirror, dst, C, src_h, src_w, crop_h, crop_w,
                                                                              scale, groups, gpr);
  return finish_launch();
}

int mnv_abi_version(void) { return 2; }
const char* mnv_build_info(void) {
  return "minerva_b200 kernels: sm_100a, nvcc " __VERSION__ ", built " __DATE__;
}
size_t mnv_workspace_bytes_hint(void) { return static_cast<size_t>(768) << 20; }
uint64_t mnv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int mnv_set_dependent_launch(int enabled) { return g_pdl.exchange(enabled ? 1 : 0, std::memory_order_relaxed); }

}  // extern "C"
