/*
 * mnv_oracle.c -- CPU restatement of Minerva's physical-op hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under minerva_b200/ may import, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use
 * it, and there only as the checker or the reported CPU baseline.
 *
 * Two kinds of function live here:
 *  (1) restatements of reference CPU code in minerva/op/impl/basic.cpp -- each cites the lines
 *      it follows; these are pinned bit-for-bit against the reference itself compiled into
 *      oracle/_ref (tests/test_oracle_vs_ref.py) and against the reference's golden vectors
 *      (tests/unittest_reduction.cpp etc., see tests/golden/);
 *  (2) restatements of ops that are NO_IMPL on the reference's CPU device
 *      (minerva/op/impl/bundle.h:29-44,51-54): conv x4, pooling x4, activation backward x3,
 *      softmax backward + channel mode, LRN x2, concat/slice.  Semantics follow the cuDNN-v2
 *      descriptors the reference configures (minerva/op/impl/cuda/cuda_perform.cu:227-615) and
 *      the LRN kernel text (cuda_kernel.h:223-331).  Pinned by goldens where the reference has
 *      them: conv forward (tests/unittest_conv_forward.cpp:7-68) and max-pool forward
 *      (tests/unittest_pooling_forward.cpp:7-81).  PARITY UNPINNED (no reference test, no
 *      reference CPU code) for: conv backward data/filter/bias, pooling backward, average
 *      pooling, softmax backward and channel mode, activation backward, LRN, concat/slice,
 *      transpose on non-zero data, max-index; those are cross-checked against torch CPU
 *      autograd in tests/test_oracle_vs_torch.py.
 *
 * Style: the reference's -- naive loops, fp32 everywhere, sequential accumulation in index
 * order.  Build with -ffp-contract=off so no FMA contraction changes rounding.  OpenMP
 * pragmas only split independent output elements across threads (each output is still
 * accumulated sequentially by one thread), so results do not depend on the thread count.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* a1 Arithmetic -- basic.cpp:23-52                                                            */
ORC_API void orc_add(const float* a, const float* b, float* c, size_t n) {
  for (size_t i = 0; i < n; ++i) c[i] = a[i] + b[i];
}
ORC_API void orc_sub(const float* a, const float* b, float* c, size_t n) {
  for (size_t i = 0; i < n; ++i) c[i] = a[i] - b[i];
}
ORC_API void orc_dot_mult(const float* a, const float* b, float* c, size_t n) {
  for (size_t i = 0; i < n; ++i) c[i] = a[i] * b[i];
}
ORC_API void orc_dot_div(const float* a, const float* b, float* c, size_t n) {
  for (size_t i = 0; i < n; ++i) c[i] = a[i] / b[i];
}

/* a2/a3 ArithmeticConst -- basic.cpp:54-107 (side 0 = const on the left) */
ORC_API void orc_const_add(const float* in, float* out, float v, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = in[i] + v;
}
ORC_API void orc_const_sub(const float* in, float* out, float v, size_t n) { /* in - v */
  for (size_t i = 0; i < n; ++i) out[i] = in[i] - v;
}
ORC_API void orc_left_const_sub(const float* in, float* out, float v, size_t n) { /* v - in */
  for (size_t i = 0; i < n; ++i) out[i] = v - in[i];
}
ORC_API void orc_scale(const float* in, float* out, size_t n, float v) {
  for (size_t i = 0; i < n; ++i) out[i] = in[i] * v;
}
ORC_API void orc_const_div(const float* in, float* out, float v, size_t n) { /* in / v */
  for (size_t i = 0; i < n; ++i) out[i] = in[i] / v;
}
ORC_API void orc_left_const_div(const float* in, float* out, float v, size_t n) { /* v / in */
  for (size_t i = 0; i < n; ++i) out[i] = v / in[i];
}

/* a4 Elewise -- basic.cpp:125-148.  `exp(float)` / `log(float)` under <cmath> resolve to the
 * float overloads, i.e. expf / logf. */
ORC_API void orc_elewise_exp(const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = expf(in[i]);
}
ORC_API void orc_elewise_ln(const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = logf(in[i]);
}
ORC_API void orc_elewise_negative(const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = -in[i];
}

/* a7 NormArithmetic, 2-D column-major {m,n} -- basic.cpp:319-367 restricted to the two cases
 * the CUDA path supports (cuda.cpp:222-257).  op: 0 add, 1 sub, 2 mult, 3 div. */
static inline float orc_apply(int op, float x, float y) {
  switch (op) {
    case 0: return x + y;
    case 1: return x - y;
    case 2: return x * y;
    default: return x / y;
  }
}
ORC_API void orc_norm_on_col(int op, const float* mat, const float* vec, float* res, int m, int n) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) res[i + (size_t)j * m] = orc_apply(op, mat[i + (size_t)j * m], vec[j]);
}
ORC_API void orc_norm_on_row(int op, const float* mat, const float* vec, float* res, int m, int n) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) res[i + (size_t)j * m] = orc_apply(op, mat[i + (size_t)j * m], vec[i]);
}

/* a5 Reduction, 2-D -- basic.cpp:189-217: first element, then fold the rest in index order;
 * max uses `if (tmp < tmp2) tmp = tmp2`. */
ORC_API void orc_reduction_on_col(int is_max, const float* in, float* out, int m, int n) {
  for (int j = 0; j < n; ++j) {
    float t = in[(size_t)j * m];
    for (int i = 1; i < m; ++i) {
      float t2 = in[i + (size_t)j * m];
      if (is_max) { if (t < t2) t = t2; } else { t += t2; }
    }
    out[j] = t;
  }
}
ORC_API void orc_reduction_on_row(int is_max, const float* in, float* out, int m, int n) {
  for (int i = 0; i < m; ++i) {
    float t = in[i];
    for (int j = 1; j < n; ++j) {
      float t2 = in[i + (size_t)j * m];
      if (is_max) { if (t < t2) t = t2; } else { t += t2; }
    }
    out[i] = t;
  }
}

/* a6 MaxIndex, 2-D -- basic.cpp:369-398: strict `<`, first maximum wins, index stored as float */
ORC_API void orc_max_index_on_col(const float* in, float* out, int m, int n) {
  for (int j = 0; j < n; ++j) {
    float best = in[(size_t)j * m];
    int idx = 0;
    for (int i = 0; i < m; ++i)
      if (best < in[i + (size_t)j * m]) { best = in[i + (size_t)j * m]; idx = i; }
    out[j] = (float)idx;
  }
}
ORC_API void orc_max_index_on_row(const float* in, float* out, int m, int n) {
  for (int i = 0; i < m; ++i) {
    float best = in[i];
    int idx = 0;
    for (int j = 0; j < n; ++j)
      if (best < in[i + (size_t)j * m]) { best = in[i + (size_t)j * m]; idx = j; }
    out[i] = (float)idx;
  }
}

/* a18 MatMult -- basic.cpp:150-173 (the non-CBLAS branch): column-major, fp32, sequential k */
ORC_API void orc_matmult(const float* a, const float* b, float* c, int m, int n, int k) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < m; ++i) {
      float acc = 0;
      for (int l = 0; l < k; ++l) acc += a[i + (size_t)l * m] * b[l + (size_t)j * k];
      c[i + (size_t)j * m] = acc;
    }
  }
}

/* a19 Transpose -- basic.cpp:175-187.  a is {m,n} column-major, c is {n,m}. */
ORC_API void orc_transpose(const float* a, float* c, int m, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) c[j + (size_t)i * n] = a[i + (size_t)j * m];
}

/* a20 Reshape -- basic.cpp:400-404 */
ORC_API void orc_copy(const float* src, float* dst, size_t n) { memcpy(dst, src, n * sizeof(float)); }

/* a23 Concat/Slice -- the per-image copies of cuda.cpp:80-155 written as one strided copy */
ORC_API void orc_copy_strided(const float* src, float* dst, size_t inner, size_t outer,
                              size_t src_stride, size_t dst_stride) {
  for (size_t b = 0; b < outer; ++b)
    memcpy(dst + b * dst_stride, src + b * src_stride, inner * sizeof(float));
}
/* Select -- intended semantics of cuda_kernel.h:333-342 */
ORC_API void orc_select(float* dst, const float* src, const int* indices, size_t n_idx,
                        size_t cols, size_t rows) {
  (void)cols;
  for (size_t j = 0; j < n_idx; ++j)
    for (size_t r = 0; r < rows; ++r) dst[r + j * rows] = src[r + (size_t)indices[j] * rows];
}

/* a21 Fill -- basic.cpp:292-299 */
ORC_API void orc_fill(float* dst, size_t n, float v) {
  for (size_t i = 0; i < n; ++i) dst[i] = v;
}

/* a10 Activation forward -- basic.cpp:406-446.  Sigmoid is evaluated in double around a
 * float expf (basic.cpp:416). */
ORC_API void orc_sigmoid_forward(const float* x, float* y, size_t n) {
  for (size_t i = 0; i < n; ++i) y[i] = (float)(1.0 / (1.0 + (double)expf(-x[i])));
}
ORC_API void orc_relu_forward(const float* x, float* y, size_t n) {
  for (size_t i = 0; i < n; ++i) y[i] = x[i] > 0 ? x[i] : 0;
}
ORC_API void orc_tanh_forward(const float* x, float* y, size_t n) {
  for (size_t i = 0; i < n; ++i) y[i] = tanhf(x[i]);
}

/* a11 Activation backward -- NO_IMPL on the reference CPU (bundle.h:29,32,34,42).  cuDNN v2
 * formulas for the modes configured at cuda_perform.cu:444-487; argument order of the
 * reference's flat functions (bottom, top, top_diff, bottom_diff). */
ORC_API void orc_sigmoid_backward(const float* x, const float* y, const float* dy, float* dx, size_t n) {
  (void)x;
  for (size_t i = 0; i < n; ++i) dx[i] = dy[i] * y[i] * (1.0f - y[i]);
}
ORC_API void orc_relu_backward(const float* x, const float* y, const float* dy, float* dx, size_t n) {
  (void)y;
  for (size_t i = 0; i < n; ++i) dx[i] = x[i] > 0 ? dy[i] : 0;
}
ORC_API void orc_tanh_backward(const float* x, const float* y, const float* dy, float* dx, size_t n) {
  (void)x;
  for (size_t i = 0; i < n; ++i) dx[i] = dy[i] * (1.0f - y[i] * y[i]);
}

/* a8 Softmax forward.  Instance mode follows basic.cpp:219-261 (max with strict `<`,
 * e = expf(x - max) stored, sequential fp32 sum, divide) over the C*H*W contiguous values of
 * each image (CUDNN_SOFTMAX_MODE_INSTANCE, cuda_perform.cu:348).  Channel mode
 * (cuda_perform.cu:363) is the same recipe over C with stride H*W -- no reference CPU code. */
static void orc_softmax_group(const float* x, float* y, int count, size_t stride) {
  float mx = x[0];
  for (int i = 1; i < count; ++i) if (mx < x[i * stride]) mx = x[i * stride];
  float sum = 0;
  for (int i = 0; i < count; ++i) { y[i * stride] = expf(x[i * stride] - mx); sum += y[i * stride]; }
  for (int i = 0; i < count; ++i) y[i * stride] /= sum;
}
ORC_API void orc_instance_softmax_forward(const float* x, float* y, int N, int C, int H, int W) {
  size_t g = (size_t)C * H * W;
  for (int n = 0; n < N; ++n) orc_softmax_group(x + n * g, y + n * g, (int)g, 1);
}
ORC_API void orc_channel_softmax_forward(const float* x, float* y, int N, int C, int H, int W) {
  size_t hw = (size_t)H * W;
  for (int n = 0; n < N; ++n)
    for (size_t p = 0; p < hw; ++p) orc_softmax_group(x + n * C * hw + p, y + n * C * hw + p, C, hw);
}
/* a9 Softmax backward: dx = y * (dy - sum_group(dy*y)) (cuDNN; cuda_perform.cu:369-397) */
static void orc_softmax_back_group(const float* dy, const float* y, float* dx, int count, size_t stride) {
  float dot = 0;
  for (int i = 0; i < count; ++i) dot += dy[i * stride] * y[i * stride];
  for (int i = 0; i < count; ++i) dx[i * stride] = y[i * stride] * (dy[i * stride] - dot);
}
ORC_API void orc_instance_softmax_backward(const float* dy, const float* y, float* dx, int N, int C, int H, int W) {
  size_t g = (size_t)C * H * W;
  for (int n = 0; n < N; ++n) orc_softmax_back_group(dy + n * g, y + n * g, dx + n * g, (int)g, 1);
}
ORC_API void orc_channel_softmax_backward(const float* dy, const float* y, float* dx, int N, int C, int H, int W) {
  size_t hw = (size_t)H * W;
  for (int n = 0; n < N; ++n)
    for (size_t p = 0; p < hw; ++p)
      orc_softmax_back_group(dy + n * C * hw + p, y + n * C * hw + p, dx + n * C * hw + p, C, hw);
}

/* ------------------------------------------------------------------------------------------ */
/* Convolution -- NO_IMPL on the reference CPU (bundle.h:35-38).  CUDNN_CONVOLUTION mode
 * (cuda_perform.cu:243): the filter is rotated by 180 degrees.  NCHW / KCRS. */
ORC_API int orc_conv_out(int x, int pad, int f, int stride) { return (x + 2 * pad - f) / stride + 1; }

ORC_API void orc_conv_forward(const float* x, const float* w, const float* b, float* y, int N, int Ci,
                              int Co, int H, int W, int ph, int pw, int sv, int sh, int fh, int fw) {
  int Ho = orc_conv_out(H, ph, fh, sv), Wo = orc_conv_out(W, pw, fw, sh);
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int co = 0; co < Co; ++co)
      for (int i = 0; i < Ho; ++i)
        for (int j = 0; j < Wo; ++j) {
          float acc = 0;
          for (int ci = 0; ci < Ci; ++ci)
            for (int kh = 0; kh < fh; ++kh) {
              int ih = i * sv - ph + kh;
              if (ih < 0 || ih >= H) continue;
              for (int kw = 0; kw < fw; ++kw) {
                int iw = j * sh - pw + kw;
                if (iw < 0 || iw >= W) continue;
                acc += x[(((size_t)n * Ci + ci) * H + ih) * W + iw] *
                       w[(((size_t)co * Ci + ci) * fh + (fh - 1 - kh)) * fw + (fw - 1 - kw)];
              }
            }
          y[(((size_t)n * Co + co) * Ho + i) * Wo + j] = acc + b[co];
        }
}

/* dx[n,ci,h,w] = sum_{co,kh,kw : h = i*sv-ph+kh, w = j*sh-pw+kw} dy[n,co,i,j] * wflip[co,ci,kh,kw].
 * Gather form so every dx element is accumulated by one thread in (co,kh,kw) order. */
ORC_API void orc_conv_backward_data(const float* dy, const float* w, float* dx, int N, int Ci, int Co,
                                    int H, int W, int ph, int pw, int sv, int sh, int fh, int fw) {
  int Ho = orc_conv_out(H, ph, fh, sv), Wo = orc_conv_out(W, pw, fw, sh);
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < N; ++n)
    for (int ci = 0; ci < Ci; ++ci)
      for (int h = 0; h < H; ++h)
        for (int x0 = 0; x0 < W; ++x0) {
          float acc = 0;
          for (int co = 0; co < Co; ++co)
            for (int kh = 0; kh < fh; ++kh) {
              int t = h + ph - kh;
              if (t < 0 || t % sv) continue;
              int i = t / sv;
              if (i >= Ho) continue;
              for (int kw = 0; kw < fw; ++kw) {
                int u = x0 + pw - kw;
                if (u < 0 || u % sh) continue;
                int j = u / sh;
                if (j >= Wo) continue;
                acc += dy[(((size_t)n * Co + co) * Ho + i) * Wo + j] *
                       w[(((size_t)co * Ci + ci) * fh + (fh - 1 - kh)) * fw + (fw - 1 - kw)];
              }
            }
          dx[(((size_t)n * Ci + ci) * H + h) * W + x0] = acc;
        }
}

/* dw[co,ci,fh-1-kh,fw-1-kw] = sum_{n,i,j} dy[n,co,i,j] * x[n,ci,i*sv-ph+kh,j*sh-pw+kw] */
ORC_API void orc_conv_backward_filter(const float* x, const float* dy, float* dw, int N, int Ci, int Co,
                                      int H, int W, int ph, int pw, int sv, int sh, int fh, int fw) {
  int Ho = orc_conv_out(H, ph, fh, sv), Wo = orc_conv_out(W, pw, fw, sh);
#pragma omp parallel for collapse(2) schedule(static)
  for (int co = 0; co < Co; ++co)
    for (int ci = 0; ci < Ci; ++ci)
      for (int kh = 0; kh < fh; ++kh)
        for (int kw = 0; kw < fw; ++kw) {
          float acc = 0;
          for (int n = 0; n < N; ++n)
            for (int i = 0; i < Ho; ++i) {
              int ih = i * sv - ph + kh;
              if (ih < 0 || ih >= H) continue;
              for (int j = 0; j < Wo; ++j) {
                int iw = j * sh - pw + kw;
                if (iw < 0 || iw >= W) continue;
                acc += dy[(((size_t)n * Co + co) * Ho + i) * Wo + j] *
                       x[(((size_t)n * Ci + ci) * H + ih) * W + iw];
              }
            }
          dw[(((size_t)co * Ci + ci) * fh + (fh - 1 - kh)) * fw + (fw - 1 - kw)] = acc;
        }
}

/* db[c] = sum_{n,h,w} dy (cudnnConvolutionBackwardBias, cuda_perform.cu:320-337).  Accumulated
 * in double and rounded once: with up to 7.7e5 terms a sequential fp32 sum is itself off by
 * ~1e-4 relative, which would make the oracle the noisier side of the comparison. */
ORC_API void orc_conv_backward_bias(const float* dy, float* db, int N, int C, int H, int W) {
  size_t hw = (size_t)H * W;
  for (int c = 0; c < C; ++c) {
    double acc = 0;
    for (int n = 0; n < N; ++n)
      for (size_t p = 0; p < hw; ++p) acc += dy[((size_t)n * C + c) * hw + p];
    db[c] = (float)acc;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Pooling -- NO_IMPL on the reference CPU (bundle.h:43-44).  Geometry: convolution.cpp:107-114
 * (== cuda_perform.cu:491-498). */
ORC_API int orc_pooled_size(int x, int pad, int window, int stride) {
  int p = (x + 2 * pad - window + stride - 1) / stride + 1;
  if (0 <= (p - 1) * stride - x - pad) --p;
  return p;
}

/* max: padding counts as -inf; first maximum in (kh-major, kw-minor) scan wins (strict >). */
ORC_API void orc_max_pooling_forward(const float* x, float* y, int N, int C, int H, int W, int sv, int sh,
                                     int wh, int ww, int ph, int pw) {
  int Ho = orc_pooled_size(H, ph, wh, sv), Wo = orc_pooled_size(W, pw, ww, sh);
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < N * C; ++nc) {
    const float* xp = x + (size_t)nc * H * W;
    float* yp = y + (size_t)nc * Ho * Wo;
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        float best = -INFINITY;
        for (int kh = 0; kh < wh; ++kh) {
          int ih = i * sv - ph + kh;
          if (ih < 0 || ih >= H) continue;
          for (int kw = 0; kw < ww; ++kw) {
            int iw = j * sh - pw + kw;
            if (iw < 0 || iw >= W) continue;
            float v = xp[ih * W + iw];
            if (v > best) best = v;
          }
        }
        yp[i * Wo + j] = best;
      }
  }
}

/* avg: CUDNN_POOLING_AVERAGE_COUNT_INCLUDE_PADDING (cuda_perform.cu:540): sum of in-range
 * values in scan order, divided by wh*ww. */
ORC_API void orc_average_pooling_forward(const float* x, float* y, int N, int C, int H, int W, int sv,
                                         int sh, int wh, int ww, int ph, int pw) {
  int Ho = orc_pooled_size(H, ph, wh, sv), Wo = orc_pooled_size(W, pw, ww, sh);
  float div = (float)(wh * ww);
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < N * C; ++nc) {
    const float* xp = x + (size_t)nc * H * W;
    float* yp = y + (size_t)nc * Ho * Wo;
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        float acc = 0;
        for (int kh = 0; kh < wh; ++kh) {
          int ih = i * sv - ph + kh;
          if (ih < 0 || ih >= H) continue;
          for (int kw = 0; kw < ww; ++kw) {
            int iw = j * sh - pw + kw;
            if (iw < 0 || iw >= W) continue;
            acc += xp[ih * W + iw];
          }
        }
        yp[i * Wo + j] = acc / div;
      }
  }
}

/* max backward: each window sends its dy to the first in-range position holding the window
 * maximum (same scan as forward); windows are visited in (i-major, j-minor) order, so a
 * bottom element that is the argmax of several windows accumulates them in that order.
 * `top` is accepted for signature parity and is not needed (the maximum is recomputed). */
ORC_API void orc_max_pooling_backward(const float* x, const float* y, const float* dy, float* dx, int N,
                                      int C, int H, int W, int sv, int sh, int wh, int ww, int ph,
                                      int pw) {
  (void)y;
  int Ho = orc_pooled_size(H, ph, wh, sv), Wo = orc_pooled_size(W, pw, ww, sh);
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < N * C; ++nc) {
    const float* xp = x + (size_t)nc * H * W;
    const float* dyp = dy + (size_t)nc * Ho * Wo;
    float* dxp = dx + (size_t)nc * H * W;
    for (int p = 0; p < H * W; ++p) dxp[p] = 0;
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        float best = -INFINITY;
        int arg = -1;
        for (int kh = 0; kh < wh; ++kh) {
          int ih = i * sv - ph + kh;
          if (ih < 0 || ih >= H) continue;
          for (int kw = 0; kw < ww; ++kw) {
            int iw = j * sh - pw + kw;
            if (iw < 0 || iw >= W) continue;
            float v = xp[ih * W + iw];
            if (v > best) { best = v; arg = ih * W + iw; }
          }
        }
        if (arg >= 0) dxp[arg] += dyp[i * Wo + j];
      }
  }
}

/* avg backward: every in-range position of a window receives dy/(wh*ww) */
ORC_API void orc_average_pooling_backward(const float* x, const float* y, const float* dy, float* dx,
                                          int N, int C, int H, int W, int sv, int sh, int wh, int ww,
                                          int ph, int pw) {
  (void)x; (void)y;
  int Ho = orc_pooled_size(H, ph, wh, sv), Wo = orc_pooled_size(W, pw, ww, sh);
  float div = (float)(wh * ww);
#pragma omp parallel for schedule(static)
  for (int nc = 0; nc < N * C; ++nc) {
    const float* dyp = dy + (size_t)nc * Ho * Wo;
    float* dxp = dx + (size_t)nc * H * W;
    for (int p = 0; p < H * W; ++p) dxp[p] = 0;
    for (int i = 0; i < Ho; ++i)
      for (int j = 0; j < Wo; ++j) {
        float g = dyp[i * Wo + j] / div;
        for (int kh = 0; kh < wh; ++kh) {
          int ih = i * sv - ph + kh;
          if (ih < 0 || ih >= H) continue;
          for (int kw = 0; kw < ww; ++kw) {
            int iw = j * sh - pw + kw;
            if (iw < 0 || iw >= W) continue;
            dxp[ih * W + iw] += g;
          }
        }
      }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* LRN across channels -- NO_IMPL on the reference CPU (bundle.h:51-52).  Restated from the
 * kernel text: LRNFillScale cuda_kernel.h:223-263 (sliding add/subtract window, `1.` is a
 * double literal so the add is done in double and rounded on store), LRNComputeOutput
 * :265-270 (`pow(float,float)` -> powf), LRNComputeDiff :272-331. */
ORC_API void orc_lrn_forward(const float* in, float* scale, float* out, int size, float alpha, float beta,
                             int N, int C, int Wd, int Ht) {
  float alpha_over_size = alpha / size;
  size_t step = (size_t)Ht * Wd;
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n)
    for (size_t p = 0; p < step; ++p) {
      const float* sin = in + (size_t)n * C * step + p;
      float* ssc = scale + (size_t)n * C * step + p;
      int head = 0;
      int pre_pad = (size - 1) / 2;
      int post_pad = size - pre_pad - 1;
      float acc = 0;
      while (head < post_pad && head < C) { acc += sin[head * step] * sin[head * step]; ++head; }
      while (head < size) {
        if (head < C) acc += sin[head * step] * sin[head * step];
        if (head - post_pad >= 0 && head - post_pad < C)
          ssc[(head - post_pad) * step] = (float)(1. + acc * alpha_over_size);
        ++head;
      }
      while (head < C) {
        acc += sin[head * step] * sin[head * step];
        acc -= sin[(head - size) * step] * sin[(head - size) * step];
        ssc[(head - post_pad) * step] = (float)(1. + acc * alpha_over_size);
        ++head;
      }
      while (head < C + post_pad) {
        if (head - size >= 0 && head - size < C) acc -= sin[(head - size) * step] * sin[(head - size) * step];
        if (head - post_pad >= 0 && head - post_pad < C)
          ssc[(head - post_pad) * step] = (float)(1. + acc * alpha_over_size);
        ++head;
      }
    }
  size_t total = (size_t)N * C * step;
  float nb = -beta;
  for (size_t i = 0; i < total; ++i) out[i] = in[i] * powf(scale[i], nb);
}

ORC_API void orc_lrn_backward(const float* bottom, const float* top, const float* scale,
                              const float* top_diff, float* bottom_diff, int size, float alpha,
                              float beta, int N, int C, int Wd, int Ht) {
  float negative_beta = -beta;
  float cache_ratio = (float)(2. * alpha * beta / size);
  size_t step = (size_t)Ht * Wd;
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n)
    for (size_t p = 0; p < step; ++p) {
      size_t off = (size_t)n * C * step + p;
      const float* b = bottom + off;
      const float* t = top + off;
      const float* s = scale + off;
      const float* td = top_diff + off;
      float* bd = bottom_diff + off;
      int head = 0;
      int pre_pad = size - (size + 1) / 2;
      int post_pad = size - pre_pad - 1;
      float acc = 0;
      while (head < post_pad && head < C) { acc += td[head * step] * t[head * step] / s[head * step]; ++head; }
      while (head < size) {
        if (head < C) acc += td[head * step] * t[head * step] / s[head * step];
        int o = head - post_pad;
        if (o >= 0 && o < C)
          bd[o * step] = td[o * step] * powf(s[o * step], negative_beta) - cache_ratio * b[o * step] * acc;
        ++head;
      }
      while (head < C) {
        acc += td[head * step] * t[head * step] / s[head * step];
        acc -= td[(head - size) * step] * t[(head - size) * step] / s[(head - size) * step];
        int o = head - post_pad;
        bd[o * step] = td[o * step] * powf(s[o * step], negative_beta) - cache_ratio * b[o * step] * acc;
        ++head;
      }
      while (head < C + post_pad) {
        int q = head - size;
        if (q >= 0 && q < C) acc -= td[q * step] * t[q * step] / s[q * step];
        int o = head - post_pad;
        if (o >= 0 && o < C)
          bd[o * step] = td[o * step] * powf(s[o * step], negative_beta) - cache_ratio * b[o * step] * acc;
        ++head;
      }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Generators.  The reference seeds std::default_random_engine / cuRAND from the wall clock
 * (basic.cpp:274,285; cuda.cpp:601,606), so only the distribution is a contract.  The CUDA
 * path uses Philox4x32-10 keyed by the seed; this is the same generator so Bernoulli masks can
 * be compared bit-for-bit and normals to float tolerance. */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  uint32_t n1 = (uint32_t)p1;
  uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
ORC_API void orc_philox4x32(uint32_t seed, uint64_t block, uint32_t stream_id, uint32_t out[4]) {
  uint32_t c[4] = {(uint32_t)block, (uint32_t)(block >> 32), stream_id, 0u};
  uint32_t k[2] = {seed, 0x0B200B20u};
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(c));
}
/* element i uses word i%4 of block i/4; u = top 24 bits * 2^-24 in [0,1); 1 if u < p */
ORC_API void orc_rand_bernoulli(float* dst, size_t n, unsigned int seed, float p) {
  for (size_t blk = 0; blk * 4 < n; ++blk) {
    uint32_t r[4];
    orc_philox4x32(seed, blk, 1u, r);
    for (int j = 0; j < 4 && blk * 4 + j < n; ++j) {
      float u = (float)(r[j] >> 8) * (1.0f / 16777216.0f);
      dst[blk * 4 + j] = u < p ? 1.0f : 0.0f;
    }
  }
}
/* Box-Muller on word pairs: u1 in (0,1], u2 in [0,1) */
ORC_API void orc_randn(float* dst, size_t n, unsigned int seed, float mean, float sd) {
  for (size_t blk = 0; blk * 4 < n; ++blk) {
    uint32_t r[4];
    float z[4];
    orc_philox4x32(seed, blk, 2u, r);
    for (int h = 0; h < 2; ++h) {
      float u1 = (float)((r[2 * h] >> 8) + 1u) * (1.0f / 16777216.0f);
      float u2 = (float)(r[2 * h + 1] >> 8) * (1.0f / 16777216.0f);
      float rad = sqrtf(-2.0f * logf(u1));
      float ang = 6.283185307179586f * u2;
      z[2 * h] = rad * cosf(ang);
      z[2 * h + 1] = rad * sinf(ang);
    }
    for (int j = 0; j < 4 && blk * 4 + j < n; ++j) dst[blk * 4 + j] = mean + sd * z[j];
  }
}

/* 8(f) rank 2 -- momentum SGD as owl/net/net.py:252-256 writes it, one tensor at a time:
 * delta = mom*delta - (lr/B)*grad - (lr*wd)*w ; w += delta */
ORC_API void orc_sgd_momentum_update(float* w, float* delta, const float* grad, size_t n, float mom,
                                     float lr_over_batch, float lr_times_wd) {
  for (size_t i = 0; i < n; ++i) {
    float d = mom * delta[i] - lr_over_batch * grad[i] - lr_times_wd * w[i];
    delta[i] = d;
    w[i] = w[i] + d;
  }
}
