// nn_ops.cu -- a8/a9 Softmax, a12/a13 Pooling, a17 ConvBackwardBias, a22 LRN.
// These replace the cuDNN-v2 calls L10, L11, L13 and the Caffe-derived LRN kernels K12-K14 of the
// reference.  Every wrapper there created/destroyed descriptors and synchronised
// the stream per call (minerva/op/impl/cuda/cuda_perform.cu:339-615); here each op is one
// enqueue-only launch (two for bias-grad with a workspace).
#include <math_constants.h>
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
  int c = blockIdx.x, sp = blockIdx.y;
  int n_per = (N + splits - 1) / splits;
  int n0 = sp * n_per, n1 = min(N, n0 + n_per);
  float acc = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float* p = dy + (static_cast<size_t>(n) * C + c) * hw;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) acc += __ldg(p + i);
  }
  acc = block_reduce(acc, false, red);
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
__global__ void __launch_bounds__(kBlock) lrn_fwd_kernel(const float* __restrict__ in, float* __restrict__ scale, float* __restrict__ out,
                                                         int num, int C, size_t step, int size, float alpha_over_size, float neg_beta) {
  size_t total = static_cast<size_t>(num) * step;
  const int pre_pad = (size - 1) / 2, post_pad = size - pre_pad - 1;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t n = t / step, p = t - n * step, off = n * C * step + p;
    const float* sin = in + off;
    float* ssc = scale + off;
    float* sout = out + off;
    float acc = 0.f;
    // head runs ahead of the output channel o = head - post_pad
    for (int head = 0; head < C + post_pad; ++head) {
      if (head < C) { float v = __ldg(sin + head * step); acc = __fadd_rn(acc, __fmul_rn(v, v)); }
      if (head >= size) { float v = __ldg(sin + (head - size) * step); acc = __fsub_rn(acc, __fmul_rn(v, v)); }
      int o = head - post_pad;
      if (o >= 0) {
        float sc = static_cast<float>(1.0 + static_cast<double>(__fmul_rn(acc, alpha_over_size)));
        ssc[o * step] = sc;
        sout[o * step] = __fmul_rn(__ldg(sin + o * step), powf(sc, neg_beta));
      }
    }
  }
}

__global__ void __launch_bounds__(kBlock) lrn_bwd_kernel(const float* __restrict__ bottom, const float* __restrict__ top,
                                                         const float* __restrict__ scale, const float* __restrict__ top_diff,
                                                         float* __restrict__ bottom_diff, int num, int C, size_t step, int size,
                                                         float neg_beta, float cache_ratio) {
  size_t total = static_cast<size_t>(num) * step;
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
    for (int head = 0; head < C + post_pad; ++head) {
      if (head < C) {
        size_t o = head * step;
        acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(__ldg(td + o), __ldg(tp + o)), __ldg(s + o)));
      }
      if (head >= size) {
        size_t o = (head - size) * step;
        acc = __fsub_rn(acc, __fdiv_rn(__fmul_rn(__ldg(td + o), __ldg(tp + o)), __ldg(s + o)));
      }
      int oc = head - post_pad;
      if (oc >= 0) {
        size_t o = oc * step;
        float lhs = __fmul_rn(__ldg(td + o), powf(__ldg(s + o), neg_beta));
        float rhs = __fmul_rn(__fmul_rn(cache_ratio, __ldg(b + o)), acc);
        bd[o] = __fsub_rn(lhs, rhs);
      }
    }
  }
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
  pool_fwd_kernel<true><<<stream_grid(planes * g.Ho * g.Wo), kBlock, 0, as_stream(s)>>>(x, y, planes, g);
  return finish_launch();
}
int mnv_average_pooling_forward(const float* x, float* y, int N, int C, int H, int W, int sv, int sh, int wh,
                                int ww, int ph, int pw, mnv_stream_t s) {
  PoolGeom g;
  int rc = make_geom(&g, N, C, H, W, sv, sh, wh, ww, ph, pw);
  if (rc) return rc;
  size_t planes = static_cast<size_t>(N) * C;
  if (planes == 0) return MNV_OK;
  if (!x || !y) return MNV_EINVAL;
  pool_fwd_kernel<false><<<stream_grid(planes * g.Ho * g.Wo), kBlock, 0, as_stream(s)>>>(x, y, planes, g);
  return finish_launch();
}
int mnv_max_pooling_backward(const float* x, const float* y, const float* dy, float* dx, int N, int C, int H,
                             int W, int sv, int sh, int wh, int ww, int ph, int pw, mnv_stream_t s) {
  PoolGeom g;
  int rc = make_geom(&g, N, C, H, W, sv, sh, wh, ww, ph, pw);
  if (rc) return rc;
  size_t planes = static_cast<size_t>(N) * C;
  if (planes == 0) return MNV_OK;
  if (!x || !y || !dy || !dx) return MNV_EINVAL;
  maxpool_bwd_kernel<<<stream_grid(planes * H * W), kBlock, 0, as_stream(s)>>>(x, y, dy, dx, planes, g);
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
  avgpool_bwd_kernel<<<stream_grid(planes * H * W), kBlock, 0, as_stream(s)>>>(dy, dx, planes, g);
  return finish_launch();
}

int mnv_conv_backward_bias(const float* dy, float* db, int N, int C, int H, int W, void* workspace,
                           size_t workspace_bytes, mnv_stream_t s) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return MNV_EINVAL;
  if (C == 0) return MNV_OK;
  if (!dy || !db) return MNV_EINVAL;
  int hw = H * W;
  // enough CTAs for ~4 per SM, bounded by the images available and the workspace
  int splits = (kNumSMs * 4 + C - 1) / C;
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

int mnv_lrn_forward(const float* bottom, float* scale, float* res, int local_size, float alpha, float beta,
                    int num_img, int channel, int width, int height, mnv_stream_t s) {
  if (num_img < 0 || channel < 0 || width < 0 || height < 0 || local_size <= 0) return MNV_EINVAL;
  size_t step = static_cast<size_t>(width) * height, work = step * num_img;
  if (work == 0 || channel == 0) return MNV_OK;
  if (!bottom || !scale || !res) return MNV_EINVAL;
  lrn_fwd_kernel<<<stream_grid(work), kBlock, 0, as_stream(s)>>>(bottom, scale, res, num_img, channel, step, local_size,
                                                               alpha / local_size, -beta);
  return finish_launch();
}
int mnv_lrn_backward(const float* bottom_data, const float* top_data, const float* scale, const float* top_diff,
                     float* bottom_diff, int local_size, float alpha, float beta, int num_img, int channel,
                     int width, int height, mnv_stream_t s) {
  if (num_img < 0 || channel < 0 || width < 0 || height < 0 || local_size <= 0) return MNV_EINVAL;
  size_t step = static_cast<size_t>(width) * height, work = step * num_img;
  if (work == 0 || channel == 0) return MNV_OK;
  if (!bottom_data || !top_data || !scale || !top_diff || !bottom_diff) return MNV_EINVAL;
  float cache_ratio = static_cast<float>(2. * alpha * beta / local_size);  // cuda_perform.cu:665
  lrn_bwd_kernel<<<stream_grid(work), kBlock, 0, as_stream(s)>>>(bottom_data, top_data, scale, top_diff, bottom_diff, num_img,
                                                               channel, step, local_size, -beta, cache_ratio);
  return finish_launch();
}

}  // extern "C"
