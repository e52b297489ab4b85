// nn_ops.cu -- a8/a9 Softmax, a12/a13 Pooling, a17 ConvBackwardBias, a22 LRN.
// These replace the cuDNN-v2 calls L10, L11, L13 and the Caffe-derived LRN kernels K12-K14 of the
// reference.  Every wrapper there created/destroyed descriptors and synchronised
// the stream per call (minerva/op/impl/cuda/cuda_perform.cu:339-615); here each op is one
// enqueue-only launch (two for bias-grad with a workspace).
#include <math_constants.h>
#include <stdlib.h>
#include "common.cuh"

namespace mnv {

// ------------------------------------------------------------------------------------------------
// Softmax.  Recipe of minerva/op/impl/basic.cpp:236-260: m = max, e = expf(x - m), s = sum e,
// y = e / s.  The sum is a tree instead of sequential (<= 1e-5 relative, north_star).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* smem) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();  // protects smem reuse between consecutive reductions
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = smem[0];
  for (int w = 1; w < nw; ++w) r = is_max ? ref_max(r, smem[w]) : r + smem[w];
  return r;
}

// instance mode: one CTA per image over `g` contiguous values
__global__ void __launch_bounds__(kBlock) softmax_instance_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int num, int g) {
  __shared__ float red[32];
  for (int n = blockIdx.x; n < num; n += gridDim.x) {
    const float* xp = x + static_cast<size_t>(n) * g;
    float* yp = y + static_cast<size_t>(n) * g;
    float mx = __ldg(xp);
    for (int i = threadIdx.x; i < g; i += blockDim.x) mx = ref_max(mx, __ldg(xp + i));
    mx = block_reduce(mx, true, red);
    float sum = 0.f;
    for (int i = threadIdx.x; i < g; i += blockDim.x) {
      float e = expf(__fsub_rn(__ldg(xp + i), mx));
      yp[i] = e;  // re-read below by the same thread
      sum += e;
    }
    sum = block_reduce(sum, false, red);
    for (int i = threadIdx.x; i < g; i += blockDim.x) yp[i] = __fdiv_rn(yp[i], sum);
  }
}

__global__ void __launch_bounds__(kBlock) softmax_instance_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                      float* __restrict__ dx, int num, int g) {
  __shared__ float red[32];
  for (int n = blockIdx.x; n < num; n += gridDim.x) {
    size_t off = static_cast<size_t>(n) * g;
    float dot = 0.f;
    for (int i = threadIdx.x; i < g; i += blockDim.x) dot += __fmul_rn(__ldg(dy + off + i), __ldg(y + off + i));
    dot = block_reduce(dot, false, red);
    for (int i = threadIdx.x; i < g; i += blockDim.x)
      dx[off + i] = __fmul_rn(__ldg(y + off + i), __fsub_rn(__ldg(dy + off + i), dot));
  }
}

// channel mode: one thread per (n, h, w), walking C with stride H*W (coalesced across threads)
__global__ void __launch_bounds__(kBlock) softmax_channel_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int num, int ch, int hw) {
  size_t total = static_cast<size_t>(num) * hw;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t n = t / hw, p = t - n * hw, base = n * ch * hw + p;
    float mx = __ldg(x + base);
    for (int c = 1; c < ch; ++c) mx = ref_max(mx, __ldg(x + base + static_cast<size_t>(c) * hw));
    float sum = 0.f;
    for (int c = 0; c < ch; ++c) {
      float e = expf(__fsub_rn(__ldg(x + base + static_cast<size_t>(c) * hw), mx));
      y[base + static_cast<size_t>(c) * hw] = e;
      sum = __fadd_rn(sum, e);
    }
    for (int c = 0; c < ch; ++c) y[base + static_cast<size_t>(c) * hw] = __fdiv_rn(y[base + static_cast<size_t>(c) * hw], sum);
  }
}

__global__ void __launch_bounds__(kBlock) softmax_channel_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                     float* __restrict__ dx, int num, int ch, int hw) {
  size_t total = static_cast<size_t>(num) * hw;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t n = t / hw, p = t - n * hw, base = n * ch * hw + p;
    float dot = 0.f;
    for (int c = 0; c < ch; ++c) {
      size_t o = base + static_cast<size_t>(c) * hw;
      dot = __fadd_rn(dot, __fmul_rn(__ldg(dy + o), __ldg(y + o)));
    }
    for (int c = 0; c < ch; ++c) {
      size_t o = base + static_cast<size_t>(c) * hw;
      dx[o] = __fmul_rn(__ldg(y + o), __fsub_rn(__ldg(dy + o), dot));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pooling.  Geometry: minerva/narray/convolution.cpp:107-114.  One thread per output (forward) /
// per input (backward, gather form => deterministic, no atomics, no memset of bottom_diff).
// ------------------------------------------------------------------------------------------------
struct PoolGeom {
  int H, W, Ho, Wo, sv, sh, wh, ww, ph, pw;
};

__host__ __device__ inline int pooled_size(int x, int pad, int window, int stride) {
  int p = (x + 2 * pad - window + stride - 1) / stride + 1;
  if (0 <= (p - 1) * stride - x - pad) --p;
  return p;
}

template <bool IS_MAX>
__global__ void __launch_bounds__(kBlock) pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t planes, PoolGeom g) {
  size_t per = static_cast<size_t>(g.Ho) * g.Wo, total = planes * per;
  const float inv_div = static_cast<float>(g.wh * g.ww);
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t pl = t / per;
    int r = static_cast<int>(t - pl * per);
    int i = r / g.Wo, j = r - i * g.Wo;
    const float* xp = x + pl * g.H * g.W;
    int h0 = i * g.sv - g.ph, w0 = j * g.sh - g.pw;
    int hs = max(h0, 0), he = min(h0 + g.wh, g.H), ws = max(w0, 0), we = min(w0 + g.ww, g.W);
    float acc = IS_MAX ? -CUDART_INF_F : 0.f;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        float v = __ldg(xp + h * g.W + w);
        if (IS_MAX) { if (v > acc) acc = v; } else { acc = __fadd_rn(acc, v); }
      }
    y[t] = IS_MAX ? acc : __fdiv_rn(acc, inv_div);
  }
}

// max backward: dx[h,w] = sum over windows (i,j) containing (h,w), in (i-major, j-minor) order, of
// dy[i,j] when (h,w) is the FIRST position of that window (h-major, w-minor scan) equal to the
// window maximum y[i,j].  `y` must be the forward output (as cuDNN requires).
__global__ void __launch_bounds__(kBlock) maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             const float* __restrict__ dy, float* __restrict__ dx, size_t planes, PoolGeom g) {
  size_t per = static_cast<size_t>(g.H) * g.W, total = planes * per;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t pl = t / per;
    int r = static_cast<int>(t - pl * per);
    int h = r / g.W, w = r - h * g.W;
    const float* xp = x + pl * per;
    const float* yp = y + pl * g.Ho * g.Wo;
    const float* dyp = dy + pl * g.Ho * g.Wo;
    float xv = __ldg(xp + r);
    // windows with i*sv - ph <= h < i*sv - ph + wh
    int i_lo = (h + g.ph - g.wh + g.sv) / g.sv;  // ceil((h+ph-wh+1)/sv) for non-negative numerators
    if (h + g.ph - g.wh + 1 <= 0) i_lo = 0;
    int i_hi = min((h + g.ph) / g.sv, g.Ho - 1);
    int j_lo = (w + g.pw - g.ww + g.sh) / g.sh;
    if (w + g.pw - g.ww + 1 <= 0) j_lo = 0;
    int j_hi = min((w + g.pw) / g.sh, g.Wo - 1);
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i)
      for (int j = j_lo; j <= j_hi; ++j) {
        float top = __ldg(yp + i * g.Wo + j);
        if (xv != top) continue;
        // is there an earlier in-range position of this window that also equals the maximum?
        int h0 = i * g.sv - g.ph, w0 = j * g.sh - g.pw;
        int hs = max(h0, 0), ws = max(w0, 0), we = min(w0 + g.ww, g.W);
        bool first = true;
        for (int hh = hs; hh <= h && first; ++hh) {
          int wend = hh == h ? w : we;
          for (int wc = ws; wc < wend; ++wc)
            if (__ldg(xp + hh * g.W + wc) == top) { first = false; break; }
        }
        if (first) acc = __fadd_rn(acc, __ldg(dyp + i * g.Wo + j));
      }
    dx[t] = acc;
  }
}

__global__ void __launch_bounds__(kBlock) avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, size_t planes, PoolGeom g) {
  size_t per = static_cast<size_t>(g.H) * g.W, total = planes * per;
  const float div = static_cast<float>(g.wh * g.ww);
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t pl = t / per;
    int r = static_cast<int>(t - pl * per);
    int h = r / g.W, w = r - h * g.W;
    const float* dyp = dy + pl * g.Ho * g.Wo;
    int i_lo = (h + g.ph - g.wh + g.sv) / g.sv;
    if (h + g.ph - g.wh + 1 <= 0) i_lo = 0;
    int i_hi = min((h + g.ph) / g.sv, g.Ho - 1);
    int j_lo = (w + g.pw - g.ww + g.sh) / g.sh;
    if (w + g.pw - g.ww + 1 <= 0) j_lo = 0;
    int j_hi = min((w + g.pw) / g.sh, g.Wo - 1);
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i)
      for (int j = j_lo; j <= j_hi; ++j) acc = __fadd_rn(acc, __fdiv_rn(__ldg(dyp + i * g.Wo + j), div));
    dx[t] = acc;
  }
}

// ---- shared-memory staged variants -------------------------------------------------------------
// A CTA stages G consecutive (n,c) planes (they are contiguous in NCHW) with coalesced loads, works
// out of shared memory, and writes its outputs coalesced: every HBM byte is touched exactly once
// (algorithmic traffic), window re-reads hit shared memory.  G is chosen so that small planes
// (13x13, 6x6) still give the CTA a few thousand elements per pass.  Index decomposition uses
// multiply-high by a host-computed reciprocal instead of integer division.
struct FastDiv {
  uint32_t d, magic;   // q = umulhi(n, magic) is exact for n*d < 2^32
};
static inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = static_cast<uint32_t>(d);
  f.magic = d == 1 ? 0u : static_cast<uint32_t>((1ull << 32) / static_cast<uint32_t>(d)) + 1u;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, FastDiv f) { return f.d == 1 ? n : __umulhi(n, f.magic); }

// Copy n floats global -> shared.  `dst` must satisfy (dst index) == (src index) mod 4 in units of
// floats relative to 16-byte boundaries (callers offset the smem base by the source misalignment), so
// the body moves 128-bit vectors on both sides; at most 3 scalars at either end.
__device__ __forceinline__ int misalign4(const float* p) { return static_cast<int>((reinterpret_cast<uintptr_t>(p) >> 2) & 3u); }
__device__ __forceinline__ void stage_in(float* __restrict__ dst, const float* __restrict__ src, int n) {
  int head = (4 - misalign4(src)) & 3;
  if (head > n) head = n;
  if (static_cast<int>(threadIdx.x) < head) dst[threadIdx.x] = __ldg(src + threadIdx.x);
  int n4 = (n - head) >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src + head);
  float4* d4 = reinterpret_cast<float4*>(dst + head);
  for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = __ldg(s4 + i);
  int done = head + (n4 << 2);
  if (static_cast<int>(threadIdx.x) < n - done) dst[done + threadIdx.x] = __ldg(src + done + threadIdx.x);
}

// The same copy with cp.async: nothing waits on a register, so every load of the group is in flight at once and the
// copy of the NEXT plane group can overlap the arithmetic on the current one (callers double-buffer and close each
// group with stage_commit(); stage_wait_prev() = all but the newest group have landed).
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async_f32(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void stage_in_async(float* __restrict__ dst, const float* __restrict__ src, int n) {
  int head = (4 - misalign4(src)) & 3;
  if (head > n) head = n;
  if (static_cast<int>(threadIdx.x) < head) cp_async_f32(dst + threadIdx.x, src + threadIdx.x);
  const int n4 = (n - head) >> 2;
  const uint32_t d4 = smem_addr(dst + head);
  const float* s4 = src + head;
  for (int i = threadIdx.x; i < n4; i += blockDim.x)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d4 + 16u * i), "l"(s4 + 4 * i) : "memory");
  const int done = head + (n4 << 2);
  if (static_cast<int>(threadIdx.x) < n - done) cp_async_f32(dst + done + threadIdx.x, src + done + threadIdx.x);
}
__device__ __forceinline__ void stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void stage_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <bool IS_MAX>
__global__ void __launch_bounds__(kBlock) pool_fwd_smem_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, PoolGeom g,
                                                                int G, FastDiv d_howo, FastDiv d_wo) {
  extern __shared__ float sm[];
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  const float div = static_cast<float>(g.wh * g.ww);
  for (int p0 = blockIdx.x * G; p0 < planes; p0 += gridDim.x * G) {
    const int cnt = min(G, planes - p0);
    const float* xsrc = x + static_cast<size_t>(p0) * HW;
    float* sx = sm + misalign4(xsrc);
    stage_in(sx, xsrc, cnt * HW);
    __syncthreads();
    float* yp = y + static_cast<size_t>(p0) * HoWo;
    for (int o = threadIdx.x; o < cnt * HoWo; o += blockDim.x) {
      int gq = fdiv(o, d_howo), r = o - gq * HoWo, i = fdiv(r, d_wo), j = r - i * g.Wo;
      const float* xp = sx + gq * HW;
      int h0 = i * g.sv - g.ph, w0 = j * g.sh - g.pw;
      int hs = max(h0, 0), he = min(h0 + g.wh, g.H), ws = max(w0, 0), we = min(w0 + g.ww, g.W);
      float acc = IS_MAX ? -CUDART_INF_F : 0.f;
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) {
          float v = xp[h * g.W + w];
          if (IS_MAX) { if (v > acc) acc = v; } else { acc = __fadd_rn(acc, v); }
        }
      yp[o] = IS_MAX ? acc : __fdiv_rn(acc, div);
    }
    __syncthreads();
  }
}

// Max backward, two phases per staged plane group (same rule as maxpool_bwd_kernel, bit-identical):
//  A. one thread per window: argmax position (first maximum in h-major, w-minor scan) -> arg[] in smem;
//  B. one thread per bottom element: sum, in (i-major, j-minor) window order, of dy over the windows whose
//     argmax is this element.  No float equality tests, `top` is not read (4*E_out bytes less traffic).
// WH/WW/SV/SH > 0 are compile-time window/stride (3x3/2, 2x2/2, 3x3/3, 3x3/1: divisions become shifts,
// loops unroll); 0 = run-time geometry.
template <int WH, int WW, int SV, int SH>
__global__ void __launch_bounds__(kBlock) maxpool_bwd_smem_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  float* __restrict__ dx, int planes, PoolGeom g, int G,
                                                                  FastDiv d_hw, FastDiv d_w, FastDiv d_howo, FastDiv d_wo,
                                                                  FastDiv d_sv, FastDiv d_sh) {
  extern __shared__ float sm[];
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  const int wh = WH ? WH : g.wh, ww = WW ? WW : g.ww, sv = SV ? SV : g.sv, sh = SH ? SH : g.sh;
  float* sx_base = sm;                                               // G*HW + 4, rounded to 16 bytes
  float* sdy_base = sm + ((G * HW + 4 + 3) & ~3);                    // G*HoWo + 4, rounded to 16 bytes
  int* sarg = reinterpret_cast<int*>(sdy_base + ((G * HoWo + 4 + 3) & ~3));   // G*HoWo
  for (int p0 = blockIdx.x * G; p0 < planes; p0 += gridDim.x * G) {
    const int cnt = min(G, planes - p0);
    const float* xsrc = x + static_cast<size_t>(p0) * HW;
    const float* dysrc = dy + static_cast<size_t>(p0) * HoWo;
    float* sx = sx_base + misalign4(xsrc);
    float* sdy = sdy_base + misalign4(dysrc);
    stage_in(sx, xsrc, cnt * HW);
    stage_in(sdy, dysrc, cnt * HoWo);
    __syncthreads();
    for (int o = threadIdx.x; o < cnt * HoWo; o += blockDim.x) {       // phase A
      int gq = fdiv(o, d_howo), r = o - gq * HoWo, i = fdiv(r, d_wo), j = r - i * g.Wo;
      const float* xp = sx + gq * HW;
      int h0 = i * sv - g.ph, w0 = j * sh - g.pw;
      float best = -CUDART_INF_F;
      int arg = -1;
#pragma unroll
      for (int kh = 0; kh < (WH ? WH : 1); ++kh)
#pragma unroll
        for (int kw = 0; kw < (WW ? WW : 1); ++kw) {
          if (WH) {
            int h = h0 + kh, w = w0 + kw;
            if (static_cast<unsigned>(h) < static_cast<unsigned>(g.H) && static_cast<unsigned>(w) < static_cast<unsigned>(g.W)) {
              float v = xp[h * g.W + w];
              if (v > best) { best = v; arg = h * g.W + w; }
            }
          }
        }
      if (!WH) {
        int hs = max(h0, 0), he = min(h0 + wh, g.H), ws = max(w0, 0), we = min(w0 + ww, g.W);
        for (int h = hs; h < he; ++h)
          for (int w = ws; w < we; ++w) {
            float v = xp[h * g.W + w];
            if (v > best) { best = v; arg = h * g.W + w; }
          }
      }
      sarg[o] = arg;
    }
    __syncthreads();
    float* dxp = dx + static_cast<size_t>(p0) * HW;
    for (int e = threadIdx.x; e < cnt * HW; e += blockDim.x) {         // phase B
      int gq = fdiv(e, d_hw), r = e - gq * HW, h = fdiv(r, d_w), w = r - h * g.W;
      const float* dyp = sdy + gq * HoWo;
      const int* ap = sarg + gq * HoWo;
      int hn = h + g.ph, wn = w + g.pw;
      int i_lo = hn - wh + 1 <= 0 ? 0 : (SV ? (hn - wh + sv) / SV : static_cast<int>(fdiv(hn - wh + sv, d_sv)));
      int i_hi = min(SV ? hn / SV : static_cast<int>(fdiv(hn, d_sv)), g.Ho - 1);
      int j_lo = wn - ww + 1 <= 0 ? 0 : (SH ? (wn - ww + sh) / SH : static_cast<int>(fdiv(wn - ww + sh, d_sh)));
      int j_hi = min(SH ? wn / SH : static_cast<int>(fdiv(wn, d_sh)), g.Wo - 1);
      float acc = 0.f;
      for (int i = i_lo; i <= i_hi; ++i)
        for (int j = j_lo; j <= j_hi; ++j)
          if (ap[i * g.Wo + j] == r) acc = __fadd_rn(acc, dyp[i * g.Wo + j]);
      dxp[e] = acc;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kBlock) avgpool_bwd_smem_kernel(const float* __restrict__ dy, float* __restrict__ dx, int planes, PoolGeom g,
                                                                  int G, FastDiv d_hw, FastDiv d_w, FastDiv d_sv, FastDiv d_sh) {
  extern __shared__ float sm[];
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo;
  const float div = static_cast<float>(g.wh * g.ww);
  for (int p0 = blockIdx.x * G; p0 < planes; p0 += gridDim.x * G) {
    const int cnt = min(G, planes - p0);
    const float* dysrc = dy + static_cast<size_t>(p0) * HoWo;
    float* sdy = sm + misalign4(dysrc);
    stage_in(sdy, dysrc, cnt * HoWo);
    __syncthreads();
    // each term of a bottom element's sum is dy / (wh * ww), rounded: divide every pooled element once here (an IEEE
    // division is ~25 instructions) instead of once per window that contains a bottom element -- the same terms, the same bits
    for (int e = threadIdx.x; e < cnt * HoWo; e += blockDim.x) sdy[e] = __fdiv_rn(sdy[e], div);
    __syncthreads();
    float* dxp = dx + static_cast<size_t>(p0) * HW;
    for (int e = threadIdx.x; e < cnt * HW; e += blockDim.x) {
      int gq = fdiv(e, d_hw), r = e - gq * HW, h = fdiv(r, d_w), w = r - h * g.W;
      const float* dyp = sm + misalign4(dysrc) + gq * HoWo;
      int i_lo = fdiv(h + g.ph - g.wh + g.sv, d_sv);
      if (h + g.ph - g.wh + 1 <= 0) i_lo = 0;
      int i_hi = min(static_cast<int>(fdiv(h + g.ph, d_sv)), g.Ho - 1);
      int j_lo = fdiv(w + g.pw - g.ww + g.sh, d_sh);
      if (w + g.pw - g.ww + 1 <= 0) j_lo = 0;
      int j_hi = min(static_cast<int>(fdiv(w + g.pw, d_sh)), g.Wo - 1);
      float acc = 0.f;
      for (int i = i_lo; i <= i_hi; ++i)
        for (int j = j_lo; j <= j_hi; ++j) acc = __fadd_rn(acc, dyp[i * g.Wo + j]);
      dxp[e] = acc;
    }
    __syncthreads();
  }
}

// ---- 3x3 / stride 2 / pad 0 max pooling (AlexNet pool1/2/5, GoogLeNet pool1-4): unrolled variants --------
// Forward: 9 unrolled LDS per output; bounds tests only when the last window overhangs (CHECK).
template <bool CHECK>
__global__ void __launch_bounds__(kBlock) maxpool332_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, PoolGeom g,
                                                                 int G, FastDiv d_howo, FastDiv d_wo, unsigned char* __restrict__ idx) {
  extern __shared__ float sm[];
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo, W = g.W;
  const int buf = (G * HW + 4 + 3) & ~3, stride = gridDim.x * G;   // two staging buffers: group k+1 streams in under group k's maxima
  int p0 = blockIdx.x * G, it = 0;
  if (p0 < planes) {
    const float* xsrc = x + static_cast<size_t>(p0) * HW;
    stage_in_async(sm + misalign4(xsrc), xsrc, min(G, planes - p0) * HW);
  }
  stage_commit();
  for (; p0 < planes; p0 += stride, it ^= 1) {
    const int cnt = min(G, planes - p0), nxt = p0 + stride;
    const float* xsrc = x + static_cast<size_t>(p0) * HW;
    if (nxt < planes) {
      const float* nsrc = x + static_cast<size_t>(nxt) * HW;
      stage_in_async(sm + (it ^ 1) * buf + misalign4(nsrc), nsrc, min(G, planes - nxt) * HW);
    }
    stage_commit();
    stage_wait_prev();
    __syncthreads();
    const float* sx = sm + it * buf + misalign4(xsrc);
    float* yp = y + static_cast<size_t>(p0) * HoWo;
    for (int o = threadIdx.x; o < cnt * HoWo; o += blockDim.x) {
      int gq = fdiv(o, d_howo), r = o - gq * HoWo, i = fdiv(r, d_wo), j = r - i * g.Wo;
      const float* xp = sx + gq * HW + (2 * i) * W + 2 * j;
      float best = -CUDART_INF_F;
      int arg = 255;   // window position kh*3+kw of the first maximum in scan order; 255 = none (a window of NaNs)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          if (!CHECK || (2 * i + kh < g.H && 2 * j + kw < W)) {
            float v = xp[kh * W + kw];
            if (v > best) { best = v; arg = kh * 3 + kw; }
          }
        }
      yp[o] = best;
      if (idx) idx[static_cast<size_t>(p0) * HoWo + o] = static_cast<unsigned char>(arg);
    }
    __syncthreads();   // this buffer is refilled by the next iteration's prefetch
  }
}

// Backward: phase A as the generic kernel (arg-max plane), phase B one thread per 2x2 input block (2i..2i+1,
// 2j..2j+1): the four windows that can own those elements are (i-1..i, j-1..j), so 4 arg + 4 dy LDS serve four
// outputs.  Contributions are added in (i-major, j-minor) window order => bit-identical to the oracle.
template <bool CHECK>
__global__ void __launch_bounds__(kBlock) maxpool332_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 float* __restrict__ dx, int planes, PoolGeom g, int G,
                                                                 FastDiv d_howo, FastDiv d_wo, FastDiv d_blk, FastDiv d_bw, int relu) {
  extern __shared__ float sm[];
  const int H = g.H, W = g.W, HW = H * W, Ho = g.Ho, Wo = g.Wo, HoWo = Ho * Wo;
  const int BH = (H + 1) >> 1, BW = (W + 1) >> 1, BLK = BH * BW;   // 2x2 blocks per plane
  // two (x, dy) staging buffers + one arg-max plane: group k+1 streams in under group k's two phases
  const int xbuf = (G * HW + 4 + 3) & ~3, dbuf = (G * HoWo + 4 + 3) & ~3, buf = xbuf + dbuf, stride = gridDim.x * G;
  int* sarg = reinterpret_cast<int*>(sm + 2 * buf);
  int p0 = blockIdx.x * G, it = 0;
  if (p0 < planes) {
    const float* xsrc = x + static_cast<size_t>(p0) * HW;
    const float* dysrc = dy + static_cast<size_t>(p0) * HoWo;
    const int cnt = min(G, planes - p0);
    stage_in_async(sm + misalign4(xsrc), xsrc, cnt * HW);
    stage_in_async(sm + xbuf + misalign4(dysrc), dysrc, cnt * HoWo);
  }
  stage_commit();
  for (; p0 < planes; p0 += stride, it ^= 1) {
    const int cnt = min(G, planes - p0), nxt = p0 + stride;
    const float* xsrc = x + static_cast<size_t>(p0) * HW;
    const float* dysrc = dy + static_cast<size_t>(p0) * HoWo;
    if (nxt < planes) {
      const float* nx = x + static_cast<size_t>(nxt) * HW;
      const float* nd = dy + static_cast<size_t>(nxt) * HoWo;
      const int ncnt = min(G, planes - nxt);
      stage_in_async(sm + (it ^ 1) * buf + misalign4(nx), nx, ncnt * HW);
      stage_in_async(sm + (it ^ 1) * buf + xbuf + misalign4(nd), nd, ncnt * HoWo);
    }
    stage_commit();
    stage_wait_prev();
    __syncthreads();
    const float* sx = sm + it * buf + misalign4(xsrc);
    const float* sdy = sm + it * buf + xbuf + misalign4(dysrc);
    for (int o = threadIdx.x; o < cnt * HoWo; o += blockDim.x) {   // phase A: arg-max of every window
      int gq = fdiv(o, d_howo), r = o - gq * HoWo, i = fdiv(r, d_wo), j = r - i * Wo;
      const int base = (2 * i) * W + 2 * j;
      const float* xp = sx + gq * HW + base;
      float best = -CUDART_INF_F;
      int arg = -1;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          if (!CHECK || (2 * i + kh < H && 2 * j + kw < W)) {
            float v = xp[kh * W + kw];
            if (v > best) { best = v; arg = base + kh * W + kw; }
          }
        }
      sarg[o] = arg;
    }
    __syncthreads();
    float* dxp = dx + static_cast<size_t>(p0) * HW;
    for (int b = threadIdx.x; b < cnt * BLK; b += blockDim.x) {    // phase B: one 2x2 input block per thread
      int gq = fdiv(b, d_blk), rb = b - gq * BLK, bi = fdiv(rb, d_bw), bj = rb - bi * BW;
      const int* ap = sarg + gq * HoWo;
      const float* dyp = sdy + gq * HoWo;
      // windows (bi-1, bj-1), (bi-1, bj), (bi, bj-1), (bi, bj); invalid ones get arg -2 (matches nothing)
      const bool iu = bi >= 1 && bi - 1 < Ho, id = bi < Ho, jl = bj >= 1 && bj - 1 < Wo, jr = bj < Wo;
      int a00 = -2, a01 = -2, a10 = -2, a11 = -2;
      float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
      if (iu && jl) { a00 = ap[(bi - 1) * Wo + bj - 1]; d00 = dyp[(bi - 1) * Wo + bj - 1]; }
      if (iu && jr) { a01 = ap[(bi - 1) * Wo + bj]; d01 = dyp[(bi - 1) * Wo + bj]; }
      if (id && jl) { a10 = ap[bi * Wo + bj - 1]; d10 = dyp[bi * Wo + bj - 1]; }
      if (id && jr) { a11 = ap[bi * Wo + bj]; d11 = dyp[bi * Wo + bj]; }
      const int h = 2 * bi, w = 2 * bj, r00 = h * W + w;
      float* out = dxp + gq * HW + r00;
      const float* xin = sx + gq * HW + r00;   // relu: the pooled tensor is a ReLU output, x > 0 is the ReLU-backward mask
      // (h, w): in all four windows
      float acc = 0.f;
      if (a00 == r00) acc = __fadd_rn(acc, d00);
      if (a01 == r00) acc = __fadd_rn(acc, d01);
      if (a10 == r00) acc = __fadd_rn(acc, d10);
      if (a11 == r00) acc = __fadd_rn(acc, d11);
      out[0] = (relu && !(xin[0] > 0.f)) ? 0.f : acc;
      if (w + 1 < W) {   // (h, w+1): windows (bi-1, bj), (bi, bj)
        acc = 0.f;
        if (a01 == r00 + 1) acc = __fadd_rn(acc, d01);
        if (a11 == r00 + 1) acc = __fadd_rn(acc, d11);
        out[1] = (relu && !(xin[1] > 0.f)) ? 0.f : acc;
      }
      if (h + 1 < H) {   // (h+1, w): windows (bi, bj-1), (bi, bj)
        acc = 0.f;
        if (a10 == r00 + W) acc = __fadd_rn(acc, d10);
        if (a11 == r00 + W) acc = __fadd_rn(acc, d11);
        out[W] = (relu && !(xin[W] > 0.f)) ? 0.f : acc;
        if (w + 1 < W) {   // (h+1, w+1): window (bi, bj)
          acc = (a11 == r00 + W + 1) ? __fadd_rn(0.f, d11) : 0.f;
          out[W + 1] = (relu && !(xin[W + 1] > 0.f)) ? 0.f : acc;
        }
      }
    }
    __syncthreads();
  }
}

// Backward from the forward pass's arg-max bytes (mnv_max_pooling_forward_idx): neither the bottom nor a staging
// buffer is needed -- a thread owns one 2x2 input block, reads the (at most four) windows that can own its elements
// straight through L1 (one byte + one float each, all loads issued before anything depends on them) and writes its
// four outputs.  Traffic: 5 B per pooled element + 4 B per input element, against 4 + 8 for the recomputing kernel.
// Same (i-major, j-minor) accumulation order, so the result is bit-identical.  mask != null: ReLU backward folded in --
// an element passes only if its window's maximum (= the element itself, being the arg-max) is > 0.
// (Assembling the planes in shared memory for 16-byte stores was measured slower: 0.29 vs 0.27 ms per AlexNet step.  Round 2
// tried two more organisations on the 55 -> 27 planes, 0.176 ms here: column strips that walk down the 2-row blocks with
// the window row above kept in registers -- 0.191 ms, the stride-2 stores are the limit, not the index arithmetic -- and a
// warp per plane with a thread per input element, whose stores are consecutive floats -- 0.371 ms, the 1-4 dependent
// byte -> float loads per element serialise.)
__global__ void __launch_bounds__(kBlock) maxpool332_bwd_idx_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx,
                                                                     const float* __restrict__ mask, float* __restrict__ dx, size_t planes,
                                                                     int G, PoolGeom g, FastDiv d_blk, FastDiv d_bw) {
  const int H = g.H, W = g.W, HW = H * W, Ho = g.Ho, Wo = g.Wo, HoWo = Ho * Wo;
  const int BW = (W + 1) >> 1, BLK = ((H + 1) >> 1) * BW;
  // a CTA walks groups of G planes; G * BLK^2 < 2^32 is checked by the host, so the fast divisions are exact
  for (size_t p0 = static_cast<size_t>(blockIdx.x) * G; p0 < planes; p0 += static_cast<size_t>(gridDim.x) * G) {
    const int cnt = static_cast<int>(planes - p0 < static_cast<size_t>(G) ? planes - p0 : G);
    for (int o = threadIdx.x; o < cnt * BLK; o += blockDim.x) {
      const int gq = static_cast<int>(fdiv(o, d_blk));
      const size_t plane = p0 + gq;
      const int rb = o - gq * BLK, bi = static_cast<int>(fdiv(rb, d_bw)), bj = rb - bi * BW;
      const unsigned char* ip = idx + plane * HoWo;
      const float* dyp = dy + plane * HoWo;
      const float* mp = mask ? mask + plane * HoWo : nullptr;
      // windows (bi-1, bj-1), (bi-1, bj), (bi, bj-1), (bi, bj)
      const bool iu = bi >= 1 && bi - 1 < Ho, id = bi < Ho, jl = bj >= 1 && bj - 1 < Wo, jr = bj < Wo;
      const int w00 = (bi - 1) * Wo + bj - 1, w01 = w00 + 1, w10 = w00 + Wo, w11 = w10 + 1;
      const bool v00 = iu && jl, v01 = iu && jr, v10 = id && jl, v11 = id && jr;
      const int a00 = v00 ? __ldg(ip + w00) : 255, a01 = v01 ? __ldg(ip + w01) : 255;
      const int a10 = v10 ? __ldg(ip + w10) : 255, a11 = v11 ? __ldg(ip + w11) : 255;
      const float d00 = v00 ? __ldg(dyp + w00) : 0.f, d01 = v01 ? __ldg(dyp + w01) : 0.f;
      const float d10 = v10 ? __ldg(dyp + w10) : 0.f, d11 = v11 ? __ldg(dyp + w11) : 0.f;
      float m00 = 1.f, m01 = 1.f, m10 = 1.f, m11 = 1.f;
      if (mp) {
        m00 = v00 ? __ldg(mp + w00) : 0.f; m01 = v01 ? __ldg(mp + w01) : 0.f;
        m10 = v10 ? __ldg(mp + w10) : 0.f; m11 = v11 ? __ldg(mp + w11) : 0.f;
      }
      // window (bi-1, bj-1) reaches the block through its position (2, 2) only; (bi-1, bj) through (2, 0..1);
      // (bi, bj-1) through (0..1, 2); (bi, bj) through (0..1, 0..1).  255 (no arg-max) matches nothing.
      const bool p00 = m00 > 0.f, p01 = m01 > 0.f, p10 = m10 > 0.f, p11 = m11 > 0.f;
      float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
      if (a00 == 8 && p00) o00 = __fadd_rn(o00, d00);
      if (a01 == 6 && p01) o00 = __fadd_rn(o00, d01);
      if (a10 == 2 && p10) o00 = __fadd_rn(o00, d10);
      if (a11 == 0 && p11) o00 = __fadd_rn(o00, d11);
      if (a01 == 7 && p01) o01 = __fadd_rn(o01, d01);
      if (a11 == 1 && p11) o01 = __fadd_rn(o01, d11);
      if (a10 == 5 && p10) o10 = __fadd_rn(o10, d10);
      if (a11 == 3 && p11) o10 = __fadd_rn(o10, d11);
      if (a11 == 4 && p11) o11 = __fadd_rn(o11, d11);
      const int h = 2 * bi, w = 2 * bj;
      float* out = dx + plane * HW + h * W + w;
      out[0] = o00;
      if (w + 1 < W) out[1] = o01;
      if (h + 1 < H) {
        out[W] = o10;
        if (w + 1 < W) out[W + 1] = o11;
      }
    }
  }
}

// ---- 3x3 / stride 1 / pad 1 max pooling (GoogLeNet's nine inception pools): column strips ---------------------------
// The generic staged kernels spend ~80 instructions per output on index arithmetic and bounds tests (15 % of the HBM peak on
// 28 x 28 planes).  Here a thread owns one COLUMN of one staged plane and walks down its rows: the column's two edge tests
// are loop invariants, a row costs three shared-memory loads, and the (value, position) of the horizontal first-maximum of
// the last three rows slides through registers.  Scan order is the oracle's (kh-major, kw-minor, strict >), so outputs and
// arg-max bytes are bit-identical to the generic kernels.  idx == null: plain forward.
__global__ void __launch_bounds__(kBlock) maxpool331_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int H, int W,
                                                                 int G, FastDiv d_w, unsigned char* __restrict__ idx) {
  extern __shared__ float sm[];
  const int HW = H * W;
  for (int p0 = blockIdx.x * G; p0 < planes; p0 += gridDim.x * G) {
    const int cnt = min(G, planes - p0);
    const float* xsrc = x + static_cast<size_t>(p0) * HW;
    float* sx = sm + misalign4(xsrc);
    stage_in(sx, xsrc, cnt * HW);
    __syncthreads();
    for (int c = threadIdx.x; c < cnt * W; c += blockDim.x) {
      const int gq = fdiv(c, d_w), w = c - gq * W;
      const bool hl = w > 0, hr = w + 1 < W;
      const float* col = sx + gq * HW + w;
      float* yp = y + (static_cast<size_t>(p0) + gq) * HW + w;
      unsigned char* ip = idx ? idx + (static_cast<size_t>(p0) + gq) * HW + w : nullptr;
      // (rv, ra): first maximum of input row r over kw = 0..2, ra = kw or 255 when nothing beat -inf
      float v0 = -CUDART_INF_F, v1 = -CUDART_INF_F, v2;
      int a0 = 255, a1 = 255, a2;
      auto row_max = [&](int r, float& rv, int& ra) {
        rv = -CUDART_INF_F; ra = 255;
        if (r < H) {
          const float* q = col + r * W;
          if (hl) { const float t = q[-1]; if (t > rv) { rv = t; ra = 0; } }
          { const float t = q[0]; if (t > rv) { rv = t; ra = 1; } }
          if (hr) { const float t = q[1]; if (t > rv) { rv = t; ra = 2; } }
        }
      };
      row_max(0, v1, a1);                       // window row kh = 1 of output row 0 (kh = 0 is padding)
      for (int h = 0; h < H; ++h) {
        row_max(h + 1, v2, a2);
        float best = -CUDART_INF_F;
        int arg = 255;
        if (v0 > best) { best = v0; arg = a0; }
        if (v1 > best) { best = v1; arg = 3 + a1; }
        if (v2 > best) { best = v2; arg = 6 + a2; }
        yp[h * W] = best;
        if (ip) ip[h * W] = static_cast<unsigned char>(arg);
        v0 = v1; a0 = a1; v1 = v2; a1 = a2;
      }
    }
    __syncthreads();
  }
}

// Backward from the arg-max bytes: dx[h][w] = sum over the (up to nine) windows (i, j), i = h-1..h+1, j = w-1..w+1, in
// (i-major, j-minor) order, of dy[i][j] where idx[i][j] names (h, w) -- position (h-i+1)*3 + (w-j+1).  Same column-strip walk:
// three (dy, idx) pairs enter per row, the 3 x 3 neighbourhood slides through registers.  dy and idx are staged (idx as
// bytes behind the floats).  Bit-identical to the generic backward kernel (same accumulation order, no float compares).
__global__ void __launch_bounds__(kBlock) maxpool331_bwd_idx_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ idx,
                                                                     const float* __restrict__ mask, float* __restrict__ dx, int planes, int H,
                                                                     int W, int G, FastDiv d_w) {
  extern __shared__ float sm[];
  const int HW = H * W;
  unsigned char* sidx_base = reinterpret_cast<unsigned char*>(sm + ((G * HW + 4 + 3) & ~3));
  for (int p0 = blockIdx.x * G; p0 < planes; p0 += gridDim.x * G) {
    const int cnt = min(G, planes - p0);
    const float* dsrc = dy + static_cast<size_t>(p0) * HW;
    const unsigned char* isrc = idx + static_cast<size_t>(p0) * HW;
    float* sdy = sm + misalign4(dsrc);
    stage_in(sdy, dsrc, cnt * HW);
    {  // the byte plane: 16-byte vectors where source and destination agree mod 16, bytes at the ends
      const int n = cnt * HW, mis = static_cast<int>(reinterpret_cast<uintptr_t>(isrc) & 15u);
      unsigned char* sidx = sidx_base + mis;
      int head = (16 - mis) & 15;
      if (head > n) head = n;
      if (static_cast<int>(threadIdx.x) < head) sidx[threadIdx.x] = __ldg(isrc + threadIdx.x);
      const int n16 = (n - head) >> 4;
      const uint4* s16 = reinterpret_cast<const uint4*>(isrc + head);
      uint4* d16 = reinterpret_cast<uint4*>(sidx + head);
      for (int i = threadIdx.x; i < n16; i += blockDim.x) d16[i] = __ldg(s16 + i);
      const int done = head + (n16 << 4);
      if (static_cast<int>(threadIdx.x) < n - done) sidx[done + threadIdx.x] = __ldg(isrc + done + threadIdx.x);
    }
    __syncthreads();
    const unsigned char* sidx = sidx_base + static_cast<int>(reinterpret_cast<uintptr_t>(isrc) & 15u);
    for (int c = threadIdx.x; c < cnt * W; c += blockDim.x) {
      const int gq = fdiv(c, d_w), w = c - gq * W;
      const bool hl = w > 0, hr = w + 1 < W;
      const float* dcol = sdy + gq * HW + w;
      const unsigned char* icol = sidx + gq * HW + w;
      float* out = dx + (static_cast<size_t>(p0) + gq) * HW + w;
      // mask != null (the pooled tensor is a ReLU output): window (i, j) passes only if its maximum top[i][j] is > 0
      const float* mcol = mask ? mask + (static_cast<size_t>(p0) + gq) * HW + w : nullptr;
      // window row i holds (d[j], a[j]) for j = w-1, w, w+1; arg 255 / out-of-range matches nothing
      float d0[3] = {0.f, 0.f, 0.f}, d1[3] = {0.f, 0.f, 0.f}, d2[3];
      int a0[3] = {255, 255, 255}, a1[3] = {255, 255, 255}, a2[3];
      auto load_row = [&](int i, float (&d)[3], int (&a)[3]) {
        d[0] = d[1] = d[2] = 0.f; a[0] = a[1] = a[2] = 255;
        if (i < H) {
          const float* q = dcol + i * W;
          const unsigned char* b = icol + i * W;
          if (hl) { d[0] = q[-1]; a[0] = b[-1]; }
          d[1] = q[0]; a[1] = b[0];
          if (hr) { d[2] = q[1]; a[2] = b[1]; }
          if (mcol) {
            const float* m = mcol + i * W;
            if (hl && !(__ldg(m - 1) > 0.f)) a[0] = 255;
            if (!(__ldg(m) > 0.f)) a[1] = 255;
            if (hr && !(__ldg(m + 1) > 0.f)) a[2] = 255;
          }
        }
      };
      load_row(0, d1, a1);
      for (int h = 0; h < H; ++h) {
        load_row(h + 1, d2, a2);
        // window (i, j) sees (h, w) at kh = h - i + 1, kw = w - j + 1: i = h-1 -> kh 2, j = w-1 -> kw 2
        float acc = 0.f;
        if (a0[0] == 8) acc = __fadd_rn(acc, d0[0]);
        if (a0[1] == 7) acc = __fadd_rn(acc, d0[1]);
        if (a0[2] == 6) acc = __fadd_rn(acc, d0[2]);
        if (a1[0] == 5) acc = __fadd_rn(acc, d1[0]);
        if (a1[1] == 4) acc = __fadd_rn(acc, d1[1]);
        if (a1[2] == 3) acc = __fadd_rn(acc, d1[2]);
        if (a2[0] == 2) acc = __fadd_rn(acc, d2[0]);
        if (a2[1] == 1) acc = __fadd_rn(acc, d2[1]);
        if (a2[2] == 0) acc = __fadd_rn(acc, d2[2]);
        out[h * W] = acc;
#pragma unroll
        for (int j = 0; j < 3; ++j) { d0[j] = d1[j]; a0[j] = a1[j]; d1[j] = d2[j]; a1[j] = a2[j]; }
      }
    }
    __syncthreads();
  }
}

constexpr int kPoolSmemMax = 96 * 1024;      // opt-in dynamic shared memory ceiling for the staged kernels
constexpr int kPoolSmemBig = 200 * 1024;     // ceiling for the double-buffered 3x3/2 kernels (GoogLeNet pool1: two 50 KB planes per CTA)
constexpr int kPoolSmemTarget = 24 * 1024;   // aim: ~6K floats per CTA pass, several CTAs per SM

// planes per CTA pass for `per_plane` floats of staging; 0 => does not fit, use the global-memory kernel
static int pool_group(size_t per_plane_floats, size_t planes, size_t max_index) {
  size_t bytes = per_plane_floats * sizeof(float);
  if (bytes > static_cast<size_t>(kPoolSmemMax)) return 0;
  size_t G = kPoolSmemTarget / bytes;
  if (G < 1) G = 1;
  // keep >= ~4 CTAs per SM busy and the fast-division precondition n*d < 2^32
  while (G > 1 && (planes / G < static_cast<size_t>(kNumSMs) * 4 || G * max_index * max_index >= (1ull << 32))) --G;
  if (max_index * max_index * G >= (1ull << 32)) return 0;
  return static_cast<int>(G);
}
template <class K>
static int pool_smem_attr(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return MNV_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoolSmemBig);
  return e == cudaSuccess ? MNV_OK : static_cast<int>(e);
}
static int pool_grid(size_t planes, int G) {
  size_t groups = (planes + G - 1) / G;
  size_t cap = static_cast<size_t>(kNumSMs) * kBlocksPerSM;
  return static_cast<int>(groups < cap ? groups : cap);
}

static int make_geom(PoolGeom* g, int N, int C, int H, int W, int sv, int sh, int wh, int ww, int ph, int pw) {
  if (N < 0 || C < 0 || H <= 0 || W <= 0 || sv <= 0 || sh <= 0 || wh <= 0 || ww <= 0 || ph < 0 || pw < 0)
    return MNV_EINVAL;
  if (ph >= wh || pw >= ww) return MNV_EUNSUPPORTED;  // a window could lie entirely in padding
  g->H = H; g->W = W; g->sv = sv; g->sh = sh; g->wh = wh; g->ww = ww; g->ph = ph; g->pw = pw;
  g->Ho = pooled_size(H, ph, wh, sv);
  g->Wo = pooled_size(W, pw, ww, sh);
  return (g->Ho > 0 && g->Wo > 0) ? MNV_OK : MNV_EINVAL;
}

// ------------------------------------------------------------------------------------------------
// ConvBackwardBias: db[c] = sum_{n,h,w} dy[n,c,h,w].  Stage 1: CTA (c, split) sums its share of the
// images; stage 2 folds the partials in split order (deterministic).  Without a workspace a single
// CTA per channel does the whole sum.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) bias_grad_partial_kernel(const float* __restrict__ dy, float* __restrict__ partial,
                                                                   int N, int C, int hw, int splits) {
  __shared__ float red[32];
  const int c = blockIdx.x, sp = blockIdx.y;
  const int n_per = (N + splits - 1) / splits;
  const int n0 = sp * n_per, n1 = min(N, n0 + n_per);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // A warp walks whole (image, channel) planes, lanes along the contiguous pixels: two planes and four 32-pixel runs
  // per step keep 8 independent 128-byte-per-warp loads in flight per lane whatever the plane size (13 x 13 planes
  // leave most of a 256-thread block idle when the block strides one plane).
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
  const size_t img_stride = static_cast<size_t>(C) * hw;
  const float* base = dy + static_cast<size_t>(c) * hw;
  int n = n0 + 2 * warp;
  for (; n + 1 < n1; n += 2 * nwarps) {
    const float* p0 = base + n * img_stride;
    const float* p1 = p0 + img_stride;
    int i = lane;
    for (; i + 96 < hw; i += 128) {
      float v0 = __ldg(p0 + i), v1 = __ldg(p0 + i + 32), v2 = __ldg(p0 + i + 64), v3 = __ldg(p0 + i + 96);
      float w0 = __ldg(p1 + i), w1 = __ldg(p1 + i + 32), w2 = __ldg(p1 + i + 64), w3 = __ldg(p1 + i + 96);
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
      b0 += w0; b1 += w1; b2 += w2; b3 += w3;
    }
    for (; i < hw; i += 32) { a0 += __ldg(p0 + i); b0 += __ldg(p1 + i); }
  }
  if (n < n1) {   // odd plane out
    const float* p0 = base + n * img_stride;
    int i = lane;
    for (; i + 96 < hw; i += 128) {
      float v0 = __ldg(p0 + i), v1 = __ldg(p0 + i + 32), v2 = __ldg(p0 + i + 64), v3 = __ldg(p0 + i + 96);
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    for (; i < hw; i += 32) a0 += __ldg(p0 + i);
  }
  float acc = block_reduce(((a0 + a1) + (a2 + a3)) + ((b0 + b1) + (b2 + b3)), false, red);
  if (threadIdx.x == 0) partial[static_cast<size_t>(sp) * C + c] = acc;
}
__global__ void bias_grad_final_kernel(const float* __restrict__ partial, float* __restrict__ db, int C, int splits) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += partial[static_cast<size_t>(s) * C + c];
  db[c] = acc;
}

// ------------------------------------------------------------------------------------------------
// LRN across channels, one thread per (n,h,w) sliding along C, coalesced across w.  Same add /
// subtract sequence as the reference kernels (cuda_kernel.h:223-331) so `scale` is bit-identical to
// the restated oracle; scale and output are produced in one pass (the reference used two kernels
// and re-read bottom and scale from HBM).
// ------------------------------------------------------------------------------------------------
// scale^(-beta): beta = 0.75 (AlexNet, GoogLeNet) is rsqrt(s) * sqrt(rsqrt(s)), three SFU ops instead of
// a ~40-instruction powf; any other beta takes powf.  Both are within a few ulp of the oracle's powf.
template <bool BETA075>
__device__ __forceinline__ float pow_neg_beta(float s, float neg_beta) {
  if (BETA075) {
    float r = rsqrtf(s);
    return __fmul_rn(r, sqrtf(r));
  }
  return powf(s, neg_beta);
}

// SIZE > 0: compile-time window with the squares kept in a register ring (no re-load of the element
// leaving the window); SIZE == 0: generic window, re-loads from L1/L2.
template <int SIZE, bool BETA075>
__global__ void __launch_bounds__(kBlock) lrn_fwd_kernel(const float* __restrict__ in, float* __restrict__ scale, float* __restrict__ out,
                                                         int num, int C, size_t step, int size_rt, float alpha_over_size, float neg_beta) {
  size_t total = static_cast<size_t>(num) * step;
  const int size = SIZE > 0 ? SIZE : size_rt;
  const int pre_pad = (size - 1) / 2, post_pad = size - pre_pad - 1;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t n = t / step, p = t - n * step, off = n * C * step + p;
    const float* sin = in + off;
    float* ssc = scale + off;
    float* sout = out + off;
    float acc = 0.f;
    float sq[SIZE > 0 ? SIZE : 1];   // squares of the last SIZE inputs
    float xv[SIZE > 0 ? SIZE : 1];   // and the inputs themselves (the output channel trails the head by post_pad)
    // head runs ahead of the output channel o = head - post_pad
    for (int head0 = 0; head0 < C + post_pad; head0 += (SIZE > 0 ? SIZE : 1)) {
#pragma unroll
      for (int u = 0; u < (SIZE > 0 ? SIZE : 1); ++u) {
        int head = head0 + u;
        if (head >= C + post_pad) break;
        if (head >= size) {
          float old = SIZE > 0 ? sq[u] : 0.f;
          if (SIZE == 0) { float v = __ldg(sin + (head - size) * step); old = __fmul_rn(v, v); }
          // subtraction happens AFTER this step's addition in the reference; keep that order below
          float add = 0.f, xin = 0.f;
          if (head < C) { xin = __ldg(sin + head * step); add = __fmul_rn(xin, xin); acc = __fadd_rn(acc, add); }
          acc = __fsub_rn(acc, old);
          if (SIZE > 0) { sq[u] = add; xv[u] = xin; }
        } else if (head < C) {
          float xin = __ldg(sin + head * step);
          float add = __fmul_rn(xin, xin);
          acc = __fadd_rn(acc, add);
          if (SIZE > 0) { sq[u] = add; xv[u] = xin; }
        }
        int o = head - post_pad;
        if (o >= 0) {
          // (float)(1. + (double)t) == fadd_rn(1.0f, t): the double sum of two floats is exact, so both round once
          float sc = __fadd_rn(1.0f, __fmul_rn(acc, alpha_over_size));
          float xo;
          if (SIZE > 0) {
            // element o sits post_pad slots behind the head in the ring
            int slot = u - post_pad;
            if (slot < 0) slot += SIZE;
            xo = xv[0];
#pragma unroll
            for (int q = 1; q < SIZE; ++q) if (slot == q) xo = xv[q];
          } else {
            xo = __ldg(sin + o * step);
          }
          ssc[o * step] = sc;
          sout[o * step] = __fmul_rn(xo, pow_neg_beta<BETA075>(sc, neg_beta));
        }
      }
    }
  }
}

template <int SIZE, bool BETA075>
__global__ void __launch_bounds__(kBlock) lrn_bwd_kernel(const float* __restrict__ bottom, const float* __restrict__ top,
                                                         const float* __restrict__ scale, const float* __restrict__ top_diff,
                                                         float* __restrict__ bottom_diff, int num, int C, size_t step, int size_rt,
                                                         float neg_beta, float cache_ratio, int relu) {
  size_t total = static_cast<size_t>(num) * step;
  const int size = SIZE > 0 ? SIZE : size_rt;
  const int pre_pad = size - (size + 1) / 2, post_pad = size - pre_pad - 1;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t n = t / step, p = t - n * step, off = n * C * step + p;
    const float* b = bottom + off;
    const float* tp = top + off;
    const float* s = scale + off;
    const float* td = top_diff + off;
    float* bd = bottom_diff + off;
    float acc = 0.f;
    float ratio[SIZE > 0 ? SIZE : 1];   // td*top/scale of the last SIZE channels
    float tdv[SIZE > 0 ? SIZE : 1], scv[SIZE > 0 ? SIZE : 1];
    for (int head0 = 0; head0 < C + post_pad; head0 += (SIZE > 0 ? SIZE : 1)) {
#pragma unroll
      for (int u = 0; u < (SIZE > 0 ? SIZE : 1); ++u) {
        int head = head0 + u;
        if (head >= C + post_pad) break;
        float old = 0.f;
        if (head >= size) {
          if (SIZE > 0) old = ratio[u];
          else { size_t o = (head - size) * step; old = __fdiv_rn(__fmul_rn(__ldg(td + o), __ldg(tp + o)), __ldg(s + o)); }
        }
        if (head < C) {
          size_t o = head * step;
          float tdh = __ldg(td + o), sch = __ldg(s + o);
          float r = __fdiv_rn(__fmul_rn(tdh, __ldg(tp + o)), sch);
          acc = __fadd_rn(acc, r);
          if (SIZE > 0) { ratio[u] = r; tdv[u] = tdh; scv[u] = sch; }
        }
        if (head >= size) acc = __fsub_rn(acc, old);
        int oc = head - post_pad;
        if (oc >= 0) {
          size_t o = oc * step;
          float tdo, sco;
          if (SIZE > 0) {
            int slot = u - post_pad;
            if (slot < 0) slot += SIZE;
            tdo = tdv[0]; sco = scv[0];
#pragma unroll
            for (int q = 1; q < SIZE; ++q) if (slot == q) { tdo = tdv[q]; sco = scv[q]; }
          } else {
            tdo = __ldg(td + o); sco = __ldg(s + o);
          }
          float lhs = __fmul_rn(tdo, pow_neg_beta<BETA075>(sco, neg_beta));
          const float bv = __ldg(b + o);
          float rhs = __fmul_rn(__fmul_rn(cache_ratio, bv), acc);
          bd[o] = (relu && !(bv > 0.f)) ? 0.f : __fsub_rn(lhs, rhs);   // relu: bottom is a ReLU output, mask = ReLU backward
        }
      }
    }
  }
}

// ---- local_size == 5 fast path: 8-channel register prefetch ----------------------------------------
// The channel walk is a serial recurrence per pixel, so the only memory-level parallelism is what a
// thread keeps in flight itself: the next 8 channels are loaded while the current 8 are consumed.
// The window lives in shifted registers; adding/subtracting an exact +0.0f outside the valid range keeps
// the reference's add-then-subtract rounding sequence bit for bit.
constexpr int kLrnChunk = 8;

// scale^(-beta) on the SFU: beta = 0.75 as r * r * rsqrt(r) with r = rsqrt(s) (two MUFU.RSQ), any other beta as
// ex2(-beta * lg2(s)).  A few ulp from powf (the parity bound for LRN outputs is 1e-5 relative).
template <bool BETA075>
__device__ __forceinline__ float pow_neg_beta_sfu(float s, float neg_beta) {
  float r;
  if (BETA075) {
    float q;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(r));
    return __fmul_rn(r, __fmul_rn(r, q));
  }
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(s));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(l, neg_beta)));
  return r;
}

// Window of 5 (AlexNet, GoogLeNet).  One thread per pixel walking the channels; 32-bit element offsets (the host
// checks the tensor has < 2^31 elements), one running offset for the loads (8 channels ahead, in registers) and one
// for the two stores; same add / subtract order as the generic kernel, so `scale` stays bit-identical.
template <bool BETA075, int CHUNK, int PX>
__global__ void __launch_bounds__(kBlock) lrn_fwd5_kernel(const float* __restrict__ in, float* __restrict__ scale, float* __restrict__ out,
                                                          int num, int C, unsigned step, float alpha_over_size, float neg_beta) {
  // A thread walks the channels of PX pixels, blockDim apart, so a CTA touches PX KB of each channel plane at a time
  // (measured: the channel-marching access pattern is bound by DRAM locality, not by the SM -- see DESIGN 5.2).
  const unsigned total = static_cast<unsigned>(num) * step;
  for (unsigned t0 = blockIdx.x * (blockDim.x * PX) + threadIdx.x; t0 < total; t0 += gridDim.x * (blockDim.x * PX)) {
    unsigned ld[PX], st[PX];
    bool live[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      const unsigned t = t0 + j * blockDim.x;
      live[j] = t < total;
      const unsigned tt = live[j] ? t : t0;
      const unsigned n = tt / step, p = tt - n * step;
      ld[j] = n * static_cast<unsigned>(C) * step + p;   // offset of the next channel to load
      st[j] = ld[j];                                      // offset of the next channel to store
    }
    float cur[PX][CHUNK], nxt[PX][CHUNK];
#pragma unroll
    for (int u = 0; u < CHUNK; ++u)
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        cur[j][u] = u < C ? __ldg(in + ld[j]) : 0.f;
        ld[j] += step;
      }
    float sq0[PX], sq1[PX], sq2[PX], sq3[PX], sq4[PX], x1[PX], x2[PX], acc[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) sq0[j] = sq1[j] = sq2[j] = sq3[j] = sq4[j] = x1[j] = x2[j] = acc[j] = 0.f;
    for (int c0 = 0; c0 < C + 2; c0 += CHUNK) {
      const int left = C - (c0 + CHUNK);   // channels still to load
#pragma unroll
      for (int u = 0; u < CHUNK; ++u)
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          nxt[j][u] = u < left ? __ldg(in + ld[j]) : 0.f;
          ld[j] += step;
        }
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        const int o = c0 + u - 2;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          const float xin = cur[j][u];                // 0 past C
          const float add = __fmul_rn(xin, xin);
          acc[j] = __fadd_rn(acc[j], add);
          acc[j] = __fsub_rn(acc[j], sq4[j]);         // x[head-5]^2, +0 while the window is filling
          sq4[j] = sq3[j]; sq3[j] = sq2[j]; sq2[j] = sq1[j]; sq1[j] = sq0[j]; sq0[j] = add;
          if (o >= 0 && o < C) {
            const float sc = __fadd_rn(1.0f, __fmul_rn(acc[j], alpha_over_size));
            const float ov = __fmul_rn(x2[j], pow_neg_beta_sfu<BETA075>(sc, neg_beta));
            if (live[j]) { if (scale) scale[st[j]] = sc; out[st[j]] = ov; }
            st[j] += step;
          }
          x2[j] = x1[j]; x1[j] = xin;
        }
      }
#pragma unroll
      for (int u = 0; u < CHUNK; ++u)
#pragma unroll
        for (int j = 0; j < PX; ++j) cur[j][u] = nxt[j][u];
    }
  }
}

// The same walk with sector-aligned stores.  Planes of 55x55 / 27x27 floats start at every 4-byte phase, so a warp
// that stores "its" 32 pixels writes two partial 32-byte sectors out of five, and the write-heavy forward pass ran at
// 64 % of HBM peak where 56x56 planes reach 83 % (tools/lrn_align_probe.py).  Here a CTA owns 256 consecutive pixels
// of ONE image; each chunk of CHUNK channels is parked in shared memory and written back rotated by the plane's
// misalignment d: thread i stores element (i - d) mod 256, so every warp but one writes whole sectors.
template <bool BETA075, int CHUNK>
__global__ void __launch_bounds__(256) lrn_fwd5_aligned_kernel(const float* __restrict__ in, float* __restrict__ scale, float* __restrict__ out,
                                                               int num, int C, unsigned step, unsigned blocks_per_img, float alpha_over_size,
                                                               float neg_beta) {
  __shared__ float buf[2][2][CHUNK][256];
  const unsigned tid = threadIdx.x;
  int pp = 0;
  for (unsigned work = blockIdx.x; work < static_cast<unsigned>(num) * blocks_per_img; work += gridDim.x) {
    const unsigned n = work / blocks_per_img, pb = work - n * blocks_per_img;
    const unsigned p0 = pb * 256u, cnt = min(256u, step - p0);
    const bool live = tid < cnt;
    const unsigned base = n * static_cast<unsigned>(C) * step + p0;   // channel 0 of this CTA's pixel range
    unsigned ld = base + tid;
    float cur[CHUNK], nxt[CHUNK];
#pragma unroll
    for (int u = 0; u < CHUNK; ++u) {
      cur[u] = (live && u < C) ? __ldg(in + ld) : 0.f;
      ld += step;
    }
    float sq0 = 0.f, sq1 = 0.f, sq2 = 0.f, sq3 = 0.f, sq4 = 0.f, x1 = 0.f, x2 = 0.f, acc = 0.f;
    for (int c0 = 0; c0 < C + 2; c0 += CHUNK) {
      const int left = C - (c0 + CHUNK);   // channels still to load
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        nxt[u] = (live && u < left) ? __ldg(in + ld) : 0.f;
        ld += step;
      }
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        const float xin = cur[u];                // 0 past C
        const float add = __fmul_rn(xin, xin);
        acc = __fadd_rn(acc, add);
        acc = __fsub_rn(acc, sq4);               // x[head-5]^2, +0 while the window is filling
        sq4 = sq3; sq3 = sq2; sq2 = sq1; sq1 = sq0; sq0 = add;
        const int o = c0 + u - 2;
        if (o >= 0 && o < C) {
          const float sc = __fadd_rn(1.0f, __fmul_rn(acc, alpha_over_size));
          if (scale) buf[pp][0][u][tid] = sc;
          buf[pp][1][u][tid] = __fmul_rn(x2, pow_neg_beta_sfu<BETA075>(sc, neg_beta));
        }
        x2 = x1; x1 = xin;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) {
        const int o = c0 + u - 2;
        if (o >= 0 && o < C) {
          const unsigned g0 = base + static_cast<unsigned>(o) * step;
          const unsigned e = (tid - (g0 & 31u)) & 255u;   // whole 128-byte lines per warp
          if (e < cnt) {
            if (scale) scale[g0 + e] = buf[pp][0][u][e];
            out[g0 + e] = buf[pp][1][u][e];
          }
        }
      }
      pp ^= 1;   // the other buffer is rewritten only after the next barrier, by which time these reads are done
#pragma unroll
      for (int u = 0; u < CHUNK; ++u) cur[u] = nxt[u];
    }
  }
}

template <bool BETA075>
__global__ void __launch_bounds__(kBlock) lrn_bwd5_kernel(const float* __restrict__ bottom, const float* __restrict__ top,
                                                          const float* __restrict__ scale, const float* __restrict__ top_diff,
                                                          float* __restrict__ bottom_diff, int num, int C, unsigned step,
                                                          float neg_beta, float cache_ratio, int relu) {
  // 32-bit element offsets as in lrn_fwd5_kernel: `ld` runs 8 channels ahead of the head for top_diff / top / scale,
  // the bottom value of output channel head-2 is fetched at ld - 2*step, `st` is the output channel.
  const unsigned total = static_cast<unsigned>(num) * step;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const unsigned n = t / step, p = t - n * step;
    unsigned ld = n * static_cast<unsigned>(C) * step + p;
    unsigned st = ld;
    const unsigned two = 2u * step;
    float ctd[kLrnChunk], ctp[kLrnChunk], csc[kLrnChunk], cb[kLrnChunk];
    float ntd[kLrnChunk], ntp[kLrnChunk], nsc[kLrnChunk], nb[kLrnChunk];
#pragma unroll
    for (int u = 0; u < kLrnChunk; ++u) {
      const bool ok = u < C;
      ctd[u] = ok ? __ldg(top_diff + ld) : 0.f;
      ctp[u] = ok ? __ldg(top + ld) : 0.f;
      csc[u] = ok ? __ldg(scale + ld) : 1.f;
      cb[u] = (u >= 2 && u - 2 < C) ? __ldg(bottom + (ld - two)) : 0.f;   // bottom of the output channel head-2
      ld += step;
    }
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f;   // td*top/scale of head-1..head-5
    float td1 = 0.f, td2 = 0.f, sc1 = 1.f, sc2 = 1.f;
    float acc = 0.f;
    for (int c0 = 0; c0 < C + 2; c0 += kLrnChunk) {
      const int left = C - (c0 + kLrnChunk);   // channels still to load
#pragma unroll
      for (int u = 0; u < kLrnChunk; ++u) {
        const bool ok = u < left;
        ntd[u] = ok ? __ldg(top_diff + ld) : 0.f;
        ntp[u] = ok ? __ldg(top + ld) : 0.f;
        nsc[u] = ok ? __ldg(scale + ld) : 1.f;
        nb[u] = (u < left + 2) ? __ldg(bottom + (ld - two)) : 0.f;
        ld += step;
      }
#pragma unroll
      for (int u = 0; u < kLrnChunk; ++u) {
        const float tdh = ctd[u], sch = csc[u];
        // scale >= 1, so the approximate divide (<= 2 ulp, no special-operand slow path -- half of top_diff*top is
        // exactly 0 after ReLU/pooling and IEEE division takes its slow path on those) is safe here
        const float r = __fdividef(__fmul_rn(tdh, ctp[u]), sch);     // +0 past C (0*0/1)
        acc = __fadd_rn(acc, r);
        acc = __fsub_rn(acc, r4);
        r4 = r3; r3 = r2; r2 = r1; r1 = r0; r0 = r;
        const int o = c0 + u - 2;
        if (o >= 0 && o < C) {
          const float lhs = __fmul_rn(td2, pow_neg_beta_sfu<BETA075>(sc2, neg_beta));
          const float rhs = __fmul_rn(__fmul_rn(cache_ratio, cb[u]), acc);
          bottom_diff[st] = (relu && !(cb[u] > 0.f)) ? 0.f : __fsub_rn(lhs, rhs);   // relu: bottom is a ReLU output, mask = ReLU backward
          st += step;
        }
        td2 = td1; td1 = tdh; sc2 = sc1; sc1 = sch;
      }
#pragma unroll
      for (int u = 0; u < kLrnChunk; ++u) { ctd[u] = ntd[u]; ctp[u] = ntp[u]; csc[u] = nsc[u]; cb[u] = nb[u]; }
    }
  }
}

// LRN backward from (bottom, top_diff) alone: `scale` and `top` are recomputed in registers with exactly the forward
// pass's operations (same add / subtract order, same SFU power), so the result is bit-identical to lrn_bwd5_kernel fed
// with the forward pass's stored arrays -- 12 B per element instead of 20, and the forward pass need not store `scale`
// (8 B instead of 12).  The recurrences are chained: head h loads channel h; scale / top / ratio exist for channel
// j = h - 2; the output channel is o = h - 4.
template <bool BETA075>
__global__ void __launch_bounds__(kBlock) lrn_bwd5_lite_kernel(const float* __restrict__ bottom, const float* __restrict__ top_diff,
                                                               float* __restrict__ bottom_diff, int num, int C, unsigned step,
                                                               float alpha_over_size, float neg_beta, float cache_ratio, int relu) {
  const unsigned total = static_cast<unsigned>(num) * step;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const unsigned n = t / step, p = t - n * step;
    unsigned ld = n * static_cast<unsigned>(C) * step + p;
    unsigned st = ld;
    float cx[kLrnChunk], ctd[kLrnChunk], nx[kLrnChunk], ntd[kLrnChunk];
#pragma unroll
    for (int u = 0; u < kLrnChunk; ++u) {
      const bool ok = u < C;
      cx[u] = ok ? __ldg(bottom + ld) : 0.f;
      ctd[u] = ok ? __ldg(top_diff + ld) : 0.f;
      ld += step;
    }
    float sq0 = 0.f, sq1 = 0.f, sq2 = 0.f, sq3 = 0.f, sq4 = 0.f, acc = 0.f;      // squares of x[h-1 .. h-5]
    float xa = 0.f, xb = 0.f, xc = 0.f, xd = 0.f;                                // x[h-1 .. h-4]
    float tda = 0.f, tdb = 0.f, tdc = 0.f, tdd = 0.f;                            // top_diff[h-1 .. h-4]
    float pa = 0.f, pb = 0.f;                                                    // scale^-beta of channels h-3, h-4
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f, acc2 = 0.f;          // top_diff*top/scale of j-1 .. j-5
    for (int c0 = 0; c0 < C + 4; c0 += kLrnChunk) {
      const int left = C - (c0 + kLrnChunk);   // channels still to load
#pragma unroll
      for (int u = 0; u < kLrnChunk; ++u) {
        const bool ok = u < left;
        nx[u] = ok ? __ldg(bottom + ld) : 0.f;
        ntd[u] = ok ? __ldg(top_diff + ld) : 0.f;
        ld += step;
      }
#pragma unroll
      for (int u = 0; u < kLrnChunk; ++u) {
        const int h = c0 + u;
        const float xin = cx[u], tdin = ctd[u];      // 0 past C
        // forward recurrence (lrn_fwd5_kernel): scale of channel j = h - 2
        const float add = __fmul_rn(xin, xin);
        acc = __fadd_rn(acc, add);
        acc = __fsub_rn(acc, sq4);
        sq4 = sq3; sq3 = sq2; sq2 = sq1; sq1 = sq0; sq0 = add;
        const int j = h - 2;
        float pj = 0.f, r = 0.f;                     // past C: 0 * 0 / 1 = +0, as lrn_bwd5_kernel gets from its masked loads
        if (j >= 0 && j < C) {
          const float sc = __fadd_rn(1.0f, __fmul_rn(acc, alpha_over_size));
          pj = pow_neg_beta_sfu<BETA075>(sc, neg_beta);
          const float top = __fmul_rn(xb, pj);       // what the forward pass stored
          r = __fdividef(__fmul_rn(tdb, top), sc);
        }
        // backward recurrence (lrn_bwd5_kernel) with j as its head: output channel o = j - 2
        acc2 = __fadd_rn(acc2, r);
        acc2 = __fsub_rn(acc2, r4);
        r4 = r3; r3 = r2; r2 = r1; r1 = r0; r0 = r;
        const int o = h - 4;
        if (o >= 0 && o < C) {
          const float lhs = __fmul_rn(tdd, pb);
          const float rhs = __fmul_rn(__fmul_rn(cache_ratio, xd), acc2);
          bottom_diff[st] = (relu && !(xd > 0.f)) ? 0.f : __fsub_rn(lhs, rhs);
          st += step;
        }
        xd = xc; xc = xb; xb = xa; xa = xin;
        tdd = tdc; tdc = tdb; tdb = tda; tda = tdin;
        pb = pa; pa = pj;
      }
#pragma unroll
      for (int u = 0; u < kLrnChunk; ++u) { cx[u] = nx[u]; ctd[u] = ntd[u]; }
    }
  }
}

template <bool IS_MAX>
static int launch_pool_fwd(const float* x, float* y, size_t planes, const PoolGeom& g, cudaStream_t s, unsigned char* idx = nullptr) {
  size_t per = static_cast<size_t>(g.H) * g.W;
  int G = planes < 0x7fffffff ? pool_group(per, planes, per) : 0;
  if (G > 0) {
    size_t bytes = (per * G + 4) * sizeof(float);
    int rc = pool_smem_attr(pool_fwd_smem_kernel<IS_MAX>, bytes);
    if (rc) return rc;
    const size_t bytes2 = 2 * ((per * G + 4 + 3) & ~static_cast<size_t>(3)) * sizeof(float);   // double-buffered staging
    if (IS_MAX && g.wh == 3 && g.ww == 3 && g.sv == 2 && g.sh == 2 && g.ph == 0 && g.pw == 0 && bytes2 <= static_cast<size_t>(kPoolSmemBig)) {
      const bool fit = (g.Ho - 1) * 2 + 3 <= g.H && (g.Wo - 1) * 2 + 3 <= g.W;
      if (fit) {
        rc = pool_smem_attr(maxpool332_fwd_kernel<false>, bytes2);
        if (rc) return rc;
        maxpool332_fwd_kernel<false><<<pool_grid(planes, G), kBlock, bytes2, s>>>(x, y, static_cast<int>(planes), g, G,
                                                                                   make_fastdiv(g.Ho * g.Wo), make_fastdiv(g.Wo), idx);
      } else {
        rc = pool_smem_attr(maxpool332_fwd_kernel<true>, bytes2);
        if (rc) return rc;
        maxpool332_fwd_kernel<true><<<pool_grid(planes, G), kBlock, bytes2, s>>>(x, y, static_cast<int>(planes), g, G,
                                                                                  make_fastdiv(g.Ho * g.Wo), make_fastdiv(g.Wo), idx);
      }
      return finish_launch();
    }
    if (IS_MAX && g.wh == 3 && g.ww == 3 && g.sv == 1 && g.sh == 1 && g.ph == 1 && g.pw == 1 && g.Ho == g.H && g.Wo == g.W &&
        static_cast<size_t>(G) * g.W < (1u << 16)) {
      maxpool331_fwd_kernel<<<pool_grid(planes, G), kBlock, bytes, s>>>(x, y, static_cast<int>(planes), g.H, g.W, G, make_fastdiv(g.W), idx);
      return finish_launch();
    }
    if (idx) return MNV_EUNSUPPORTED;
    pool_fwd_smem_kernel<IS_MAX><<<pool_grid(planes, G), kBlock, bytes, s>>>(x, y, static_cast<int>(planes), g, G,
                                                                            make_fastdiv(g.Ho * g.Wo), make_fastdiv(g.Wo));
    return finish_launch();
  }
  if (idx) return MNV_EUNSUPPORTED;
  pool_fwd_kernel<IS_MAX><<<stream_grid(planes * g.Ho * g.Wo), kBlock, 0, s>>>(x, y, planes, g);
  return finish_launch();
}

}  // namespace mnv

using namespace mnv;

extern "C" {

int mnv_pooled_size(int x, int pad, int window, int stride) { return pooled_size(x, pad, window, stride); }

int mnv_instance_softmax_forward(const float* x, float* y, int N, int C, int H, int W, mnv_stream_t s) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return MNV_EINVAL;
  long long g = static_cast<long long>(C) * H * W;
  if (N == 0 || g == 0) return MNV_OK;
  if (!x || !y || g > 0x7fffffffLL) return MNV_EINVAL;
  softmax_instance_fwd_kernel<<<min(N, kNumSMs * kBlocksPerSM), kBlock, 0, as_stream(s)>>>(x, y, N, static_cast<int>(g));
  return finish_launch();
}
int mnv_instance_softmax_backward(const float* dy, const float* y, float* dx, int N, int C, int H, int W, mnv_stream_t s) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return MNV_EINVAL;
  long long g = static_cast<long long>(C) * H * W;
  if (N == 0 || g == 0) return MNV_OK;
  if (!dy || !y || !dx || g > 0x7fffffffLL) return MNV_EINVAL;
  softmax_instance_bwd_kernel<<<min(N, kNumSMs * kBlocksPerSM), kBlock, 0, as_stream(s)>>>(dy, y, dx, N, static_cast<int>(g));
  return finish_launch();
}
int mnv_channel_softmax_forward(const float* x, float* y, int N, int C, int H, int W, mnv_stream_t s) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return MNV_EINVAL;
  size_t work = static_cast<size_t>(N) * H * W;
  if (work == 0 || C == 0) return MNV_OK;
  if (!x || !y) return MNV_EINVAL;
  softmax_channel_fwd_kernel<<<stream_grid(work), kBlock, 0, as_stream(s)>>>(x, y, N, C, H * W);
  return finish_launch();
}
int mnv_channel_softmax_backward(const float* dy, const float* y, float* dx, int N, int C, int H, int W, mnv_stream_t s) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return MNV_EINVAL;
  size_t work = static_cast<size_t>(N) * H * W;
  if (work == 0 || C == 0) return MNV_OK;
  if (!dy || !y || !dx) return MNV_EINVAL;
  softmax_channel_bwd_kernel<<<stream_grid(work), kBlock, 0, as_stream(s)>>>(dy, y, dx, N, C, H * W);
  return finish_launch();
}

int mnv_max_pooling_forward(const float* x, float* y, int N, int C, int H, int W, int sv, int sh, int wh, int ww,
                            int ph, int pw, mnv_stream_t s) {
  PoolGeom g;
  int rc = make_geom(&g, N, C, H, W, sv, sh, wh, ww, ph, pw);
  if (rc) return rc;
  size_t planes = static_cast<size_t>(N) * C;
  if (planes == 0) return MNV_OK;
  if (!x || !y) return MNV_EINVAL;
  return launch_pool_fwd<true>(x, y, planes, g, as_stream(s));
}
int mnv_average_pooling_forward(const float* x, float* y, int N, int C, int H, int W, int sv, int sh, int wh,
                                int ww, int ph, int pw, mnv_stream_t s) {
  PoolGeom g;
  int rc = make_geom(&g, N, C, H, W, sv, sh, wh, ww, ph, pw);
  if (rc) return rc;
  size_t planes = static_cast<size_t>(N) * C;
  if (planes == 0) return MNV_OK;
  if (!x || !y) return MNV_EINVAL;
  return launch_pool_fwd<false>(x, y, planes, g, as_stream(s));
}
static int max_pooling_backward_impl(const float* x, const float* y, const float* dy, float* dx, int N, int C, int H,
                                     int W, int sv, int sh, int wh, int ww, int ph, int pw, mnv_stream_t s, bool* relu) {
  // *relu in: apply the ReLU-backward mask (x > 0) to dx; out: whether the kernel that ran did (else the caller does)
  const int want_relu = *relu ? 1 : 0;
  *relu = false;
  PoolGeom g;
  int rc = make_geom(&g, N, C, H, W, sv, sh, wh, ww, ph, pw);
  if (rc) return rc;
  size_t planes = static_cast<size_t>(N) * C;
  if (planes == 0) return MNV_OK;
  if (!x || !y || !dy || !dx) return MNV_EINVAL;
  {
    size_t per = static_cast<size_t>(H) * W + 2 * static_cast<size_t>(g.Ho) * g.Wo;   // x, dy, argmax
    int G = planes < 0x7fffffff ? pool_group(per, planes, static_cast<size_t>(H) * W) : 0;
    if (G > 0) {
      size_t bytes = (per * G + 16) * sizeof(float);
      const int P = static_cast<int>(planes);
      const FastDiv f1 = make_fastdiv(H * W), f2 = make_fastdiv(W), f3 = make_fastdiv(g.Ho * g.Wo), f4 = make_fastdiv(g.Wo),
                    f5 = make_fastdiv(sv), f6 = make_fastdiv(sh);
      const int grid = pool_grid(planes, G);
      cudaStream_t st = as_stream(s);
#define MNV_POOL_BWD(WH_, WW_, SV_, SH_)                                                                         \
  do {                                                                                                            \
    rc = pool_smem_attr(maxpool_bwd_smem_kernel<WH_, WW_, SV_, SH_>, bytes);                                      \
    if (rc) return rc;                                                                                            \
    maxpool_bwd_smem_kernel<WH_, WW_, SV_, SH_><<<grid, kBlock, bytes, st>>>(x, dy, dx, P, g, G, f1, f2, f3, f4, f5, f6); \
  } while (0)
      // double-buffered (x, dy) staging + one arg-max plane
      const size_t HW_ = static_cast<size_t>(H) * W, HoWo_ = static_cast<size_t>(g.Ho) * g.Wo;
      const size_t bytes2 = (2 * (((G * HW_ + 4 + 3) & ~static_cast<size_t>(3)) + ((G * HoWo_ + 4 + 3) & ~static_cast<size_t>(3))) + G * HoWo_) * sizeof(float);
      if (wh == 3 && ww == 3 && sv == 2 && sh == 2 && ph == 0 && pw == 0 && bytes2 <= static_cast<size_t>(kPoolSmemBig)) {
        const bool fit = (g.Ho - 1) * 2 + 3 <= H && (g.Wo - 1) * 2 + 3 <= W;
        const FastDiv fb = make_fastdiv(((H + 1) / 2) * ((W + 1) / 2)), fbw = make_fastdiv((W + 1) / 2);
        if (fit) {
          rc = pool_smem_attr(maxpool332_bwd_kernel<false>, bytes2);
          if (rc) return rc;
          maxpool332_bwd_kernel<false><<<grid, kBlock, bytes2, st>>>(x, dy, dx, P, g, G, f3, f4, fb, fbw, want_relu);
        } else {
          rc = pool_smem_attr(maxpool332_bwd_kernel<true>, bytes2);
          if (rc) return rc;
          maxpool332_bwd_kernel<true><<<grid, kBlock, bytes2, st>>>(x, dy, dx, P, g, G, f3, f4, fb, fbw, want_relu);
        }
        *relu = want_relu != 0;
      }
      else if (wh == 3 && ww == 3 && sv == 2 && sh == 2) MNV_POOL_BWD(3, 3, 2, 2);
      else if (wh == 2 && ww == 2 && sv == 2 && sh == 2) MNV_POOL_BWD(2, 2, 2, 2);
      else if (wh == 3 && ww == 3 && sv == 3 && sh == 3) MNV_POOL_BWD(3, 3, 3, 3);
      else if (wh == 3 && ww == 3 && sv == 1 && sh == 1) MNV_POOL_BWD(3, 3, 1, 1);
      else MNV_POOL_BWD(0, 0, 0, 0);
#undef MNV_POOL_BWD
      return finish_launch();
    }
  }
  maxpool_bwd_kernel<<<stream_grid(planes * H * W), kBlock, 0, as_stream(s)>>>(x, y, dy, dx, planes, g);
  return finish_launch();
}
int mnv_max_pooling_backward(const float* x, const float* y, const float* dy, float* dx, int N, int C, int H,
                             int W, int sv, int sh, int wh, int ww, int ph, int pw, mnv_stream_t s) {
  bool relu = false;
  return max_pooling_backward_impl(x, y, dy, dx, N, C, H, W, sv, sh, wh, ww, ph, pw, s, &relu);
}
int mnv_max_pooling_backward_relu(const float* x, const float* y, const float* dy, float* dx, int N, int C, int H,
                                  int W, int sv, int sh, int wh, int ww, int ph, int pw, mnv_stream_t s) {
  bool relu = true;
  int rc = max_pooling_backward_impl(x, y, dy, dx, N, C, H, W, sv, sh, wh, ww, ph, pw, s, &relu);
  if (rc || relu) return rc;
  // geometries without a fused kernel: the mask as a second, explicitly in-place elementwise pass
  return mnv_relu_mask_inplace(dx, x, static_cast<size_t>(N) * C * H * W, s);
}
int mnv_max_pooling_idx_supported(int N, int C, int H, int W, int sv, int sh, int wh, int ww, int ph, int pw) {
  PoolGeom g;
  if (make_geom(&g, N, C, H, W, sv, sh, wh, ww, ph, pw)) return 0;
  const size_t planes = static_cast<size_t>(N) * C, per = static_cast<size_t>(H) * W;
  if (planes == 0 || planes >= 0x7fffffff) return 0;
  if (wh == 3 && ww == 3 && sv == 1 && sh == 1 && ph == 1 && pw == 1) {   // inception pools: the column-strip pair
    const int G1 = pool_group(per, planes, per);
    return G1 > 0 && static_cast<size_t>(G1) * W < (1u << 16) ? 1 : 0;
  }
  if (!(wh == 3 && ww == 3 && sv == 2 && sh == 2 && ph == 0 && pw == 0)) return 0;
  const int G = pool_group(per, planes, per);
  if (G <= 0 || 2 * ((per * G + 4 + 3) & ~static_cast<size_t>(3)) * sizeof(float) > static_cast<size_t>(kPoolSmemBig)) return 0;   // forward staging
  const size_t blk = static_cast<size_t>((H + 1) / 2) * ((W + 1) / 2);
  return blk * blk < (1ull << 32) ? 1 : 0;
}
int mnv_max_pooling_forward_idx(const float* x, float* y, unsigned char* idx, int N, int C, int H, int W, int sv, int sh,
                                int wh, int ww, int ph, int pw, mnv_stream_t s) {
  PoolGeom g;
  int rc = make_geom(&g, N, C, H, W, sv, sh, wh, ww, ph, pw);
  if (rc) return rc;
  size_t planes = static_cast<size_t>(N) * C;
  if (planes == 0) return MNV_OK;
  if (!x || !y || !idx) return MNV_EINVAL;
  if (!mnv_max_pooling_idx_supported(N, C, H, W, sv, sh, wh, ww, ph, pw)) return MNV_EUNSUPPORTED;
  return launch_pool_fwd<true>(x, y, planes, g, as_stream(s), idx);
}
int mnv_max_pooling_backward_idx(const float* dy, const unsigned char* idx, const float* relu_top, float* dx, int N, int C,
                                 int H, int W, int sv, int sh, int wh, int ww, int ph, int pw, mnv_stream_t s) {
  PoolGeom g;
  int rc = make_geom(&g, N, C, H, W, sv, sh, wh, ww, ph, pw);
  if (rc) return rc;
  size_t planes = static_cast<size_t>(N) * C;
  if (planes == 0) return MNV_OK;
  if (!dy || !idx || !dx) return MNV_EINVAL;
  if (wh == 3 && ww == 3 && sv == 1 && sh == 1 && ph == 1 && pw == 1) {
    if (!mnv_max_pooling_idx_supported(N, C, H, W, sv, sh, wh, ww, ph, pw)) return MNV_EUNSUPPORTED;
    const size_t per = static_cast<size_t>(H) * W;
    const int G1 = pool_group(per, planes, per);
    const size_t bytes = ((per * G1 + 4 + 3) & ~static_cast<size_t>(3)) * sizeof(float) + per * G1 + 32;
    rc = pool_smem_attr(maxpool331_bwd_idx_kernel, bytes);
    if (rc) return rc;
    maxpool331_bwd_idx_kernel<<<pool_grid(planes, G1), kBlock, bytes, as_stream(s)>>>(dy, idx, relu_top, dx, static_cast<int>(planes), H, W,
                                                                                       G1, make_fastdiv(W));
    return finish_launch();
  }
  if (!(wh == 3 && ww == 3 && sv == 2 && sh == 2 && ph == 0 && pw == 0)) return MNV_EUNSUPPORTED;
  const size_t blk = static_cast<size_t>((H + 1) / 2) * ((W + 1) / 2);
  size_t G = 2048 / blk;                  // ~2K input blocks (8 per thread) per CTA pass
  if (G < 1) G = 1;
  while (G > 1 && planes / G < static_cast<size_t>(kNumSMs) * 8) --G;
  if (G * blk * blk >= (1ull << 32)) return MNV_EUNSUPPORTED;   // fast-division precondition
  const size_t groups = (planes + G - 1) / G, cap = static_cast<size_t>(kNumSMs) * kBlocksPerSM;
  maxpool332_bwd_idx_kernel<<<static_cast<unsigned>(groups < cap ? groups : cap), kBlock, 0, as_stream(s)>>>(
      dy, idx, relu_top, dx, planes, static_cast<int>(G), g, make_fastdiv(static_cast<int>(blk)), make_fastdiv((W + 1) / 2));
  return finish_launch();
}
int mnv_average_pooling_backward(const float* x, const float* y, const float* dy, float* dx, int N, int C,
                                 int H, int W, int sv, int sh, int wh, int ww, int ph, int pw, mnv_stream_t s) {
  (void)x; (void)y;
  PoolGeom g;
  int rc = make_geom(&g, N, C, H, W, sv, sh, wh, ww, ph, pw);
  if (rc) return rc;
  size_t planes = static_cast<size_t>(N) * C;
  if (planes == 0) return MNV_OK;
  if (!dy || !dx) return MNV_EINVAL;
  {
    size_t per = static_cast<size_t>(g.Ho) * g.Wo;
    int G = planes < 0x7fffffff ? pool_group(per, planes, static_cast<size_t>(H) * W) : 0;
    if (G > 0) {
      size_t bytes = (per * G + 4) * sizeof(float);
      rc = pool_smem_attr(avgpool_bwd_smem_kernel, bytes);
      if (rc) return rc;
      avgpool_bwd_smem_kernel<<<pool_grid(planes, G), kBlock, bytes, as_stream(s)>>>(dy, dx, static_cast<int>(planes), g, G,
                                                                                  make_fastdiv(H * W), make_fastdiv(W), make_fastdiv(sv), make_fastdiv(sh));
      return finish_launch();
    }
  }
  avgpool_bwd_kernel<<<stream_grid(planes * H * W), kBlock, 0, as_stream(s)>>>(dy, dx, planes, g);
  return finish_launch();
}

int mnv_conv_backward_bias(const float* dy, float* db, int N, int C, int H, int W, void* workspace,
                           size_t workspace_bytes, mnv_stream_t s) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return MNV_EINVAL;
  if (C == 0) return MNV_OK;
  if (!dy || !db) return MNV_EINVAL;
  int hw = H * W;
  // 16 images per CTA (8 warps x 2 planes, one pass), but no fewer CTAs than ~8 per SM; bounded by the workspace
  int splits = (N + 15) / 16;
  const int want = (kNumSMs * 8 + C - 1) / C;
  if (splits < want) splits = want;
  if (splits > N) splits = N > 0 ? N : 1;
  size_t max_splits = workspace ? workspace_bytes / (sizeof(float) * C) : 0;
  if (static_cast<size_t>(splits) > max_splits) splits = static_cast<int>(max_splits);
  if (splits <= 1) {
    bias_grad_partial_kernel<<<dim3(C, 1), kBlock, 0, as_stream(s)>>>(dy, db, N, C, hw, 1);
    return finish_launch();
  }
  float* partial = static_cast<float*>(workspace);
  bias_grad_partial_kernel<<<dim3(C, splits), kBlock, 0, as_stream(s)>>>(dy, partial, N, C, hw, splits);
  int rc = finish_launch();
  if (rc) return rc;
  bias_grad_final_kernel<<<(C + 127) / 128, 128, 0, as_stream(s)>>>(partial, db, C, splits);
  return finish_launch();
}

static bool lrn_fast5(int local_size, int channel, size_t work, int lookahead) {
  // the size-5 kernels keep 32-bit element offsets (their loads run up to `lookahead` channels past the tensor's end, unissued)
  return local_size == 5 && channel >= 5 && (static_cast<unsigned long long>(channel) + lookahead) * work < (1ull << 31);
}
static int lrn_forward_impl(const float* bottom, float* scale, float* res, int local_size, float alpha, float beta,
                            int num_img, int channel, int width, int height, mnv_stream_t s) {
  if (num_img < 0 || channel < 0 || width < 0 || height < 0 || local_size <= 0) return MNV_EINVAL;
  size_t step = static_cast<size_t>(width) * height, work = step * num_img;
  if (work == 0 || channel == 0) return MNV_OK;
  if (!bottom || !res) return MNV_EINVAL;
  if (!scale && !lrn_fast5(local_size, channel, work, 32)) return MNV_EUNSUPPORTED;   // scale-less form: window 5 only
  const float aos = alpha / local_size;
  const bool b075 = beta == 0.75f;
  const int grid = stream_grid(work);
  // the size-5 kernel keeps 32-bit element offsets (its loads run up to 16 channels past the tensor's end, unissued)
  if (local_size == 5 && channel >= 5 && (static_cast<unsigned long long>(channel) + 32) * work < (1ull << 31)) {
    const unsigned step32 = static_cast<unsigned>(step);
    // 16 channels of register look-ahead: measured 3960 GB/s on 256x96x55x55 against 3600 with 8 (the kernel is bound by
    // DRAM latency under a 1-read : 2-write mix; evict-first / write-through stores made no difference)
    // (the scale-less form writes one array, not two: there the plain walk is faster, 5.1 vs 4.1 TB/s on 55x55)
    if ((step32 & 7u) != 0 && step32 >= 128 && scale) {
      // planes that do not start on sector boundaries (odd sizes): stores rotated through shared memory (4400 vs 3940 GB/s
      // on 55x55, 4060 vs 3760 on 27x27)
      const unsigned bpi = (step32 + 255u) / 256u;
      const size_t items = static_cast<size_t>(num_img) * bpi;
      const int g2 = static_cast<int>(items < static_cast<size_t>(kNumSMs) * 6 ? items : static_cast<size_t>(kNumSMs) * 6);
      if (b075) lrn_fwd5_aligned_kernel<true, 8><<<g2, 256, 0, as_stream(s)>>>(bottom, scale, res, num_img, channel, step32, bpi, aos, -beta);
      else lrn_fwd5_aligned_kernel<false, 8><<<g2, 256, 0, as_stream(s)>>>(bottom, scale, res, num_img, channel, step32, bpi, aos, -beta);
      return finish_launch();
    }
    // sector-aligned planes: 16 channels of register look-ahead, one pixel per thread (5.2-5.6 TB/s on 28x28 / 32x32 / 56x56)
    if (b075) lrn_fwd5_kernel<true, 16, 1><<<grid, kBlock, 0, as_stream(s)>>>(bottom, scale, res, num_img, channel, step32, aos, -beta);
    else lrn_fwd5_kernel<false, 16, 1><<<grid, kBlock, 0, as_stream(s)>>>(bottom, scale, res, num_img, channel, step32, aos, -beta);
  } else {
    if (b075) lrn_fwd_kernel<0, true><<<grid, kBlock, 0, as_stream(s)>>>(bottom, scale, res, num_img, channel, step, local_size, aos, -beta);
    else lrn_fwd_kernel<0, false><<<grid, kBlock, 0, as_stream(s)>>>(bottom, scale, res, num_img, channel, step, local_size, aos, -beta);
  }
  return finish_launch();
}
int mnv_lrn_forward(const float* bottom, float* scale, float* res, int local_size, float alpha, float beta,
                    int num_img, int channel, int width, int height, mnv_stream_t s) {
  if (!scale && static_cast<size_t>(width) * height * num_img != 0 && channel != 0) return MNV_EINVAL;
  return lrn_forward_impl(bottom, scale, res, local_size, alpha, beta, num_img, channel, width, height, s);
}
int mnv_lrn_forward_lite(const float* bottom, float* res, int local_size, float alpha, float beta,
                         int num_img, int channel, int width, int height, mnv_stream_t s) {
  return lrn_forward_impl(bottom, nullptr, res, local_size, alpha, beta, num_img, channel, width, height, s);
}
int mnv_lrn_backward_lite(const float* bottom_data, const float* top_diff, float* bottom_diff, int local_size, float alpha,
                          float beta, int num_img, int channel, int width, int height, int relu, mnv_stream_t s) {
  if (num_img < 0 || channel < 0 || width < 0 || height < 0 || local_size <= 0) return MNV_EINVAL;
  size_t step = static_cast<size_t>(width) * height, work = step * num_img;
  if (work == 0 || channel == 0) return MNV_OK;
  if (!bottom_data || !top_diff || !bottom_diff) return MNV_EINVAL;
  if (!lrn_fast5(local_size, channel, work, 32)) return MNV_EUNSUPPORTED;
  const float cache_ratio = static_cast<float>(2. * alpha * beta / local_size);  // cuda_perform.cu:665
  const float aos = alpha / local_size;
  const unsigned step32 = static_cast<unsigned>(step);
  const int grid = stream_grid(work);
  if (beta == 0.75f) lrn_bwd5_lite_kernel<true><<<grid, kBlock, 0, as_stream(s)>>>(bottom_data, top_diff, bottom_diff, num_img, channel, step32, aos, -beta, cache_ratio, relu);
  else lrn_bwd5_lite_kernel<false><<<grid, kBlock, 0, as_stream(s)>>>(bottom_data, top_diff, bottom_diff, num_img, channel, step32, aos, -beta, cache_ratio, relu);
  return finish_launch();
}
static int lrn_backward_impl(const float* bottom_data, const float* top_data, const float* scale, const float* top_diff,
                             float* bottom_diff, int local_size, float alpha, float beta, int num_img, int channel,
                             int width, int height, mnv_stream_t s, int relu) {
  if (num_img < 0 || channel < 0 || width < 0 || height < 0 || local_size <= 0) return MNV_EINVAL;
  size_t step = static_cast<size_t>(width) * height, work = step * num_img;
  if (work == 0 || channel == 0) return MNV_OK;
  if (!bottom_data || !top_data || !scale || !top_diff || !bottom_diff) return MNV_EINVAL;
  float cache_ratio = static_cast<float>(2. * alpha * beta / local_size);  // cuda_perform.cu:665
  const bool b075 = beta == 0.75f;
  const int grid = stream_grid(work);
  if (local_size == 5 && channel >= 5 && (static_cast<unsigned long long>(channel) + 16) * work < (1ull << 31)) {
    const unsigned step32 = static_cast<unsigned>(step);
    if (b075) lrn_bwd5_kernel<true><<<grid, kBlock, 0, as_stream(s)>>>(bottom_data, top_data, scale, top_diff, bottom_diff, num_img, channel, step32, -beta, cache_ratio, relu);
    else lrn_bwd5_kernel<false><<<grid, kBlock, 0, as_stream(s)>>>(bottom_data, top_data, scale, top_diff, bottom_diff, num_img, channel, step32, -beta, cache_ratio, relu);
  } else {
    if (b075) lrn_bwd_kernel<0, true><<<grid, kBlock, 0, as_stream(s)>>>(bottom_data, top_data, scale, top_diff, bottom_diff, num_img, channel, step, local_size, -beta, cache_ratio, relu);
    else lrn_bwd_kernel<0, false><<<grid, kBlock, 0, as_stream(s)>>>(bottom_data, top_data, scale, top_diff, bottom_diff, num_img, channel, step, local_size, -beta, cache_ratio, relu);
  }
  return finish_launch();
}

int mnv_lrn_backward(const float* bottom_data, const float* top_data, const float* scale, const float* top_diff,
                     float* bottom_diff, int local_size, float alpha, float beta, int num_img, int channel,
                     int width, int height, mnv_stream_t s) {
  return lrn_backward_impl(bottom_data, top_data, scale, top_diff, bottom_diff, local_size, alpha, beta, num_img, channel, width, height, s, 0);
}
int mnv_lrn_backward_relu(const float* bottom_data, const float* top_data, const float* scale, const float* top_diff,
                          float* bottom_diff, int local_size, float alpha, float beta, int num_img, int channel,
                          int width, int height, mnv_stream_t s) {
  return lrn_backward_impl(bottom_data, top_data, scale, top_diff, bottom_diff, local_size, alpha, beta, num_img, channel, width, height, s, 1);
}

}  // extern "C"
