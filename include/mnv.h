/*
 * mnv.h -- C ABI of the B200-native replacement for Minerva's physical-op kernel layer.
 *
 * This is the drop-in boundary: one entry point per row of the reference's inner function
 * table (reference: minerva/op/impl/cuda/cuda_perform.h:12-76, 53 free functions in
 * namespace minerva::cuda).  The reference's host shims (minerva/op/impl/cuda.cpp) unpack
 * DataList/closure into flat pointers + ints and call that table; a maintainer re-points
 * those shims at the functions below.
 *
 * Conventions (same as the reference unless stated):
 *   - every pointer is a DEVICE pointer to dense fp32, resident on the current device;
 *   - Scale lists the fastest dim first: a {W,H,C,N} image batch is byte-identical to a
 *     C-order NCHW array, a {fw,fh,Cin,Cout} filter to KCRS, a {m,n} matrix is column-major
 *     (reference: minerva/common/scale.cpp:182-191);
 *   - `stream` is a cudaStream_t passed as void*.  Every call is enqueue-only on that stream:
 *     no device synchronisation, no cudaMalloc/cudaFree, no cudaSetDevice, thread-safe and
 *     re-entrant (the reference calls from 4 worker threads per GpuDevice,
 *     minerva/device/device.cpp:138,214-222);
 *   - outputs never alias inputs; inputs may alias each other;
 *   - return value: 0 on success, a positive cudaError_t from the launch, or a negative
 *     MNV_E* code for an argument error.  The reference's `void` + CHECK-fatal behaviour is
 *     kept by the host shim, which CHECKs the return code.
 *   - no cuBLAS / cuDNN / cuRAND / Thrust behind any of these.
 *
 * Workspace: kernels that need scratch memory (split-K partials, bias-grad partials) take an
 * explicit (workspace, workspace_bytes) pair owned by the device layer, one per stream, as the
 * reference's PooledDataStore would hand out.  Passing NULL/0 is legal and selects the
 * scratch-free schedule (every convolution / MatMult entry has one; the only exception is
 * mnv_matmult_ex, which returns MNV_EWORKSPACE when an operand's alignment forces it to
 * materialise a transpose and no workspace was given).  mnv_workspace_bytes_hint() returns a size that lets every schedule
 * run for the AlexNet / GoogLeNet shapes.
 */
#ifndef MNV_H_
#define MNV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mnv_stream_t; /* cudaStream_t */

#define MNV_OK 0
#define MNV_EINVAL (-1)      /* bad argument (null pointer, negative dim, ...) */
#define MNV_EUNSUPPORTED (-2) /* shape outside what the kernel handles */
#define MNV_EWORKSPACE (-3)  /* workspace too small for a schedule that needs it */

/* ---- library info ------------------------------------------------------------------------- */
int mnv_abi_version(void);            /* bumps when a signature below changes */
const char* mnv_build_info(void);     /* "sm_100a nvcc 12.9 ..." */
size_t mnv_workspace_bytes_hint(void);
/* number of kernels this library has launched in this process (all threads); bench.py reads
 * it before/after the timed region to report "gpu_launches". */
uint64_t mnv_launch_count(void);
/* Programmatic dependent launch (on by default): the library's kernels let their successor in the stream start its
 * prologue while they drain, and wait for their predecessor's completion before their first global access -- the
 * stream order of results is unchanged.  0 = plain launches (for A/B timing).  -> the previous setting. */
int mnv_set_dependent_launch(int enabled);

/* ---- a1 Arithmetic: c = a o b  (cuda_perform.h:12-14,16; cuda_perform.cu:32-64) ------------ */
int mnv_add(const float* a, const float* b, float* c, size_t n, mnv_stream_t stream);
int mnv_sub(const float* a, const float* b, float* c, size_t n, mnv_stream_t stream);
/* dst = ((srcs[0] + srcs[1]) + srcs[2]) + ... : the bits of count - 1 chained mnv_add calls, in one pass.  `srcs` is a HOST
 * array of 1..8 device pointers; dst may be srcs[0] (the one sanctioned alias: an accumulation that continues a sum), no
 * other overlap.  owl.net sums the sensitivities of a blob with several consumers with it (net.py:1102-1114). */
int mnv_add_n(const float* const* srcs, int count, float* dst, size_t n, mnv_stream_t stream);
int mnv_dot_mult(const float* a, const float* b, float* c, size_t n, mnv_stream_t stream);
int mnv_dot_div(const float* a, const float* b, float* c, size_t n, mnv_stream_t stream);

/* ---- a2/a3 ArithmeticConst (cuda_perform.h:18,21-23; cuda.cpp:170-200) --------------------- */
int mnv_const_add(const float* in, float* out, float val, size_t n, mnv_stream_t stream);      /* in + val */
int mnv_left_const_sub(const float* in, float* out, float val, size_t n, mnv_stream_t stream); /* val - in */
int mnv_left_const_div(const float* in, float* out, float val, size_t n, mnv_stream_t stream); /* val / in */
int mnv_scale(const float* in, float* out, size_t n, float val, mnv_stream_t stream);          /* in * val */
/* New: the reference CUDA path computes in/val as in*(1/val), which is not
 * bit-equal to basic.cpp:99-103.  This entry does the IEEE division. */
int mnv_const_div(const float* in, float* out, float val, size_t n, mnv_stream_t stream);      /* in / val */

/* ---- a4 Elewise (cuda_perform.h:46-48) ----------------------------------------------------- */
int mnv_elewise_exp(const float* in, float* out, size_t n, mnv_stream_t stream);
int mnv_elewise_ln(const float* in, float* out, size_t n, mnv_stream_t stream);
int mnv_elewise_negative(const float* in, float* out, size_t n, mnv_stream_t stream);

/* Exact mode: the same three ops returning the CPU reference's bits -- glibc 2.39's expf / logf restated in
 * double precision (minerva_b200/csrc/glibc_math.h, pinned against libm on all 2^32 inputs).  The default entries above
 * use CUDA's expf / logf (<= 2 ulp from glibc, ~1.3x faster). */
int mnv_elewise_exp_exact(const float* in, float* out, size_t n, mnv_stream_t stream);
int mnv_elewise_ln_exact(const float* in, float* out, size_t n, mnv_stream_t stream);

/* ---- a7 NormArithmetic on a column-major {m,n} matrix (cuda_perform.h:25-33) ---------------
 * "OnCol": vec has n entries, res[i + j*m] = matrix[i + j*m] o vec[j]   (dims_to_replicate {0})
 * "OnRow": vec has m entries, res[i + j*m] = matrix[i + j*m] o vec[i]   (dims_to_replicate {1}) */
int mnv_norm_add_on_col(const float* matrix, const float* vec, float* res, int m, int n, mnv_stream_t stream);
int mnv_norm_sub_on_col(const float* matrix, const float* vec, float* res, int m, int n, mnv_stream_t stream);
int mnv_norm_mult_on_col(const float* matrix, const float* vec, float* res, int m, int n, mnv_stream_t stream);
int mnv_norm_div_on_col(const float* matrix, const float* vec, float* res, int m, int n, mnv_stream_t stream);
int mnv_norm_add_on_row(const float* matrix, const float* vec, float* res, int m, int n, mnv_stream_t stream);
int mnv_norm_sub_on_row(const float* matrix, const float* vec, float* res, int m, int n, mnv_stream_t stream);
int mnv_norm_mult_on_row(const float* matrix, const float* vec, float* res, int m, int n, mnv_stream_t stream);
int mnv_norm_div_on_row(const float* matrix, const float* vec, float* res, int m, int n, mnv_stream_t stream);

/* ---- a5 Reduction / a6 MaxIndex on a column-major {m,n} matrix (cuda_perform.h:35-41) -------
 * "OnCol": out has n entries, out[j] = reduce_i in[i + j*m]      (dims_to_reduce {0})
 * "OnRow": out has m entries, out[i] = reduce_j in[i + j*m]      (dims_to_reduce {1})
 * max: first operand kept on ties / NaN in a later slot is skipped, as basic.cpp:206-209.
 * max-index: first index of the maximum, strict `<` (basic.cpp:391), written as float. */
int mnv_reduction_sum_on_col(const float* in, float* out, int m, int n, mnv_stream_t stream);
int mnv_reduction_max_on_col(const float* in, float* out, int m, int n, mnv_stream_t stream);
int mnv_reduction_sum_on_row(const float* in, float* out, int m, int n, mnv_stream_t stream);
int mnv_reduction_max_on_row(const float* in, float* out, int m, int n, mnv_stream_t stream);
int mnv_max_index_on_col(const float* in, float* out, int m, int n, mnv_stream_t stream);
int mnv_max_index_on_row(const float* in, float* out, int m, int n, mnv_stream_t stream);

/* ---- a19/a20 copies (cuda_perform.h:15,19,43) ---------------------------------------------- */
int mnv_copy(const float* src, float* dst, size_t n, mnv_stream_t stream);
int mnv_reshape(const float* in, float* out, size_t bytes, mnv_stream_t stream);
/* c{n,m} = transpose(a{m,n}); a column-major m x n. */
int mnv_transpose(const float* a, float* c, int m, int n, mnv_stream_t stream);
/* a23 Concat/Slice: the reference issues one cublasScopy per image (cuda.cpp:80-155).  One
 * launch here: copies `outer` blocks of `inner` contiguous floats, block b going from
 * src + b*src_stride to dst + b*dst_stride (strides in floats). */
int mnv_copy_strided(const float* src, float* dst, size_t inner, size_t outer,
                     size_t src_stride, size_t dst_stride, mnv_stream_t stream);
/* up to 8 such copies that share `outer`, in one launch: Concat of up to 8 arrays (cuda.cpp:75-108 issues one copy per
 * input) and the Slices that take a concatenated gradient apart (host array of segments). */
typedef struct mnv_copy_seg_t {
  const float* src;
  float* dst;
  size_t inner, src_stride, dst_stride;
} mnv_copy_seg_t;
int mnv_copy_strided_n(const mnv_copy_seg_t* segs, int count, size_t outer, mnv_stream_t stream);
/* Select (cuda_perform.h:74): dst{rows,n_idx} = columns `indices` of src{rows,cols}.  The
 * reference passes a host pointer to the device (cuda_perform.cu:676); here `indices` is a
 * device array of n_idx ints. */
int mnv_select(float* dst, const float* src, const int* indices, size_t n_idx, size_t cols,
               size_t rows, mnv_stream_t stream);

/* ---- a18 MatMult (cuda_perform.h:17; cuda_perform.cu:66-70) --------------------------------
 * c{m,n} = a{m,k} * b{k,n}, all column-major, alpha=1 beta=0.  TF32 tcgen05 tensor-core
 * kernel with fp32 accumulation in TMEM.  workspace (optional) enables split-K. */
int mnv_matmult(const float* a, const float* b, float* c, int m, int n, int k,
                void* workspace, size_t workspace_bytes, mnv_stream_t stream);
/* Extension (host-path): c{m,n} = op(a) * op(b) with op = transpose when trans_x != 0 (a stored {k,m},
 * b stored {n,k}).  The reference materialises `x.trans()` with cublasSgeam before every such product
 * (narray.cpp:116-123 + cuda_perform.cu:72-76; owl FullyConnected.bp, owl/owl/net/net.py:612-614); both operand
 * orders are native UMMA layouts, so owl's lazy `trans()` feeds them here without the copy. */
int mnv_matmult_ex(const float* a, const float* b, float* c, int m, int n, int k, int trans_a, int trans_b,
                   void* workspace, size_t workspace_bytes, mnv_stream_t stream);

/* ---- a14-a17 Convolution (cuda_perform.h:50-53; cuda_perform.cu:227-337) -------------------
 * NCHW fp32, CUDNN_CONVOLUTION mode (filter rotated 180 degrees):
 *   top[n,co,i,j] = bias[co] + sum_{ci,kh,kw} bottom[n,ci,i*sv-ph+kh,j*sh-pw+kw]
 *                                              * filter[co,ci,fh-1-kh,fw-1-kw]
 *   top_h = (bottom_h + 2*ph - fh)/sv + 1 (floor), likewise width (convolution.cpp:13-18).
 * Implicit-GEMM tcgen05 kernels (TF32 inputs, fp32 accumulate). */
int mnv_conv_forward(const float* bottom, const float* filter, const float* bias, float* top,
                     int num_images, int bottom_num_channels, int top_num_channels,
                     int bottom_height, int bottom_width, int pad_height, int pad_width,
                     int stride_vertical, int stride_horizontal, int filter_height,
                     int filter_width, void* workspace, size_t workspace_bytes,
                     mnv_stream_t stream);
/* Extension (fusion): the same convolution with max(x, 0) applied in the epilogue -- bit-identical to
 * mnv_conv_forward followed by mnv_relu_forward.  owl.net uses it when a ConvConnection's only consumer is a ReluUnit
 * (owl/owl/net/net.py:621-716 + :281-296 run them as two ops and re-read the activation from HBM). */
int mnv_conv_forward_relu(const float* bottom, const float* filter, const float* bias, float* top,
                     int num_images, int bottom_num_channels, int top_num_channels,
                     int bottom_height, int bottom_width, int pad_height, int pad_width,
                     int stride_vertical, int stride_horizontal, int filter_height,
                     int filter_width, void* workspace, size_t workspace_bytes,
                     mnv_stream_t stream);
/* Adjoint w.r.t. bottom.  The reference derives bottom_h from top_h (cuda_perform.cu:277),
 * which is wrong when (H+2p-f)%s != 0; NArray passes the true bottom shape
 * (convolution.cpp:29-48), so it is explicit here. */
int mnv_conv_backward_data(const float* top_diff, const float* filter, float* bottom_diff,
                           int num_images, int bottom_num_channels, int top_num_channels,
                           int bottom_height, int bottom_width, int pad_height, int pad_width,
                           int stride_vertical, int stride_horizontal, int filter_height,
                           int filter_width, void* workspace, size_t workspace_bytes,
                           mnv_stream_t stream);
int mnv_conv_backward_filter(const float* bottom, const float* top_diff, float* filter_diff,
                             int num_images, int bottom_num_channels, int top_num_channels,
                             int bottom_height, int bottom_width, int pad_height, int pad_width,
                             int stride_vertical, int stride_horizontal, int filter_height,
                             int filter_width, void* workspace, size_t workspace_bytes,
                             mnv_stream_t stream);
/* bias_diff[c] = sum_{n,h,w} top_diff[n,c,h,w] */
int mnv_conv_backward_bias(const float* top_diff, float* bias_diff, int num_images,
                           int top_num_channels, int top_height, int top_width,
                           void* workspace, size_t workspace_bytes, mnv_stream_t stream);

/* ---- a8/a9 Softmax (cuda_perform.h:54-57) --------------------------------------------------
 * instance: over C*H*W per image; channel: over C per (n,h,w).  Max-subtracted, expf, fp32. */
int mnv_instance_softmax_forward(const float* bottom, float* top, int num_images, int num_channels,
                                 int height, int width, mnv_stream_t stream);
int mnv_channel_softmax_forward(const float* bottom, float* top, int num_images, int num_channels,
                                int height, int width, mnv_stream_t stream);
/* bottom_diff = top * (top_diff - sum_group(top_diff * top)) */
int mnv_instance_softmax_backward(const float* top_diff, const float* top, float* bottom_diff,
                                  int num_images, int num_channels, int height, int width,
                                  mnv_stream_t stream);
int mnv_channel_softmax_backward(const float* top_diff, const float* top, float* bottom_diff,
                                 int num_images, int num_channels, int height, int width,
                                 mnv_stream_t stream);

/* ---- a10/a11 Activation (cuda_perform.h:58-63) ---------------------------------------------
 * Element count is num_images*num_channels*height*width (the reference's Sigmoid/Relu/Tanh
 * shims pass 1,1,1,Prod; cuda.cpp:345,363,381). */
int mnv_sigmoid_forward(const float* bottom, float* top, int num_images, int num_channels,
                        int height, int width, mnv_stream_t stream);
int mnv_relu_forward(const float* bottom, float* top, int num_images, int num_channels,
                     int height, int width, mnv_stream_t stream);
int mnv_tanh_forward(const float* bottom, float* top, int num_images, int num_channels,
                     int height, int width, mnv_stream_t stream);
/* Exact mode: sigmoid as (float)(1.0 / (1.0 + (double)expf(-x))) with glibc's expf (basic.cpp:416), tanh as glibc's tanhf
 * (fdlibm over expm1f, basic.cpp:444) -- bit-identical to the reference's CPU ops; relu is exact in the default entry. */
int mnv_sigmoid_forward_exact(const float* bottom, float* top, int num_images, int num_channels,
                              int height, int width, mnv_stream_t stream);
int mnv_tanh_forward_exact(const float* bottom, float* top, int num_images, int num_channels,
                           int height, int width, mnv_stream_t stream);
/* sigmoid: dx = dy*y*(1-y); relu: dx = x>0 ? dy : 0; tanh: dx = dy*(1-y*y) */
int mnv_sigmoid_backward(const float* bottom, const float* top, const float* top_diff,
                         float* bottom_diff, int num_images, int num_channels, int height,
                         int width, mnv_stream_t stream);
int mnv_relu_backward(const float* bottom, const float* top, const float* top_diff,
                      float* bottom_diff, int num_images, int num_channels, int height,
                      int width, mnv_stream_t stream);
int mnv_tanh_backward(const float* bottom, const float* top, const float* top_diff,
                      float* bottom_diff, int num_images, int num_channels, int height,
                      int width, mnv_stream_t stream);

/* ---- a12/a13 Pooling (cuda_perform.h:64-67; cuda_perform.cu:489-615) -----------------------
 * Output size: P=(X+2p-k+s-1)/s+1; if ((P-1)*s >= X+p) --P  (convolution.cpp:107-114).
 * max: padding is -inf, first maximum in (h-major, w-minor) window scan wins;
 * avg: CUDNN_POOLING_AVERAGE_COUNT_INCLUDE_PADDING, divisor is always wh*ww. */
int mnv_max_pooling_forward(const float* bottom, float* top, int num_images, int num_channels,
                            int bottom_height, int bottom_width, int stride_vertical,
                            int stride_horizontal, int window_height, int window_width,
                            int pad_height, int pad_width, mnv_stream_t stream);
int mnv_average_pooling_forward(const float* bottom, float* top, int num_images, int num_channels,
                                int bottom_height, int bottom_width, int stride_vertical,
                                int stride_horizontal, int window_height, int window_width,
                                int pad_height, int pad_width, mnv_stream_t stream);
int mnv_max_pooling_backward(const float* bottom, const float* top, const float* top_diff,
                             float* bottom_diff, int num_images, int num_channels,
                             int bottom_height, int bottom_width, int stride_vertical,
                             int stride_horizontal, int window_height, int window_width,
                             int pad_height, int pad_width, mnv_stream_t stream);
int mnv_average_pooling_backward(const float* bottom, const float* top, const float* top_diff,
                                 float* bottom_diff, int num_images, int num_channels,
                                 int bottom_height, int bottom_width, int stride_vertical,
                                 int stride_horizontal, int window_height, int window_width,
                                 int pad_height, int pad_width, mnv_stream_t stream);
/* helper exported for the host shims: the ceil-mode size rule above */
int mnv_pooled_size(int x, int pad, int window, int stride);

/* ---- a21 generators (cuda_perform.h:69-71) -------------------------------------------------
 * Philox4x32-10 counter-based streams keyed by `seed` (the reference uses cuRAND XORWOW seeded
 * from the wall clock, cuda.cpp:601,606 -- only the distribution is a contract).
 * randn: N(mean, var^2) -- `var` is used as the standard deviation on both reference paths
 * (basic.cpp:275, cuda_perform.cu:621).  Runs on `stream` (the reference's ran off-stream). */
int mnv_randn(float* dst, size_t n, unsigned int seed, float mean, float var, mnv_stream_t stream);
int mnv_rand_bernoulli(float* dst, size_t n, unsigned int seed, float p, mnv_stream_t stream);
/* The same Bernoulli stream with its key read on the device: seed = (*seed_base + seed_add) ^ seed_xor.  For callers that
 * record a training step into a CUDA graph (owl.net.NetTrainer(graph=True)): the launch is frozen, the mask is not -- the
 * caller stores the step's base word before every replay.  Bit-identical to mnv_rand_bernoulli with that seed. */
int mnv_rand_bernoulli_ds(float* dst, size_t n, const unsigned int* seed_base, unsigned int seed_add,
                          unsigned int seed_xor, float p, mnv_stream_t stream);
int mnv_fill(float* dst, size_t n, float val, mnv_stream_t stream);

/* ---- a22 LRN across channels (cuda_perform.h:72-73; cuda_kernel.h:223-331) ------------------
 * scale = 1 + (alpha/local_size) * sum_window x^2 (written to `scale`, an in/out buffer as in
 * the reference); res = bottom * scale^-beta.  One fused pass. */
int mnv_lrn_forward(const float* bottom, float* scale, float* res, int local_size, float alpha,
                    float beta, int num_img, int channel, int width, int height,
                    mnv_stream_t stream);
int mnv_lrn_backward(const float* bottom_data, const float* top_data, const float* scale,
                     const float* top_diff, float* bottom_diff, int local_size, float alpha,
                     float beta, int num_img, int channel, int width, int height,
                     mnv_stream_t stream);
/* Extensions (fusion): the same backward passes followed by ReLU backward, for a bottom that is a ReLU
 * output (mask = bottom > 0; both kernels read the bottom anyway) -- bit-identical to the unfused pair
 * mnv_relu_backward(bottom, bottom, mnv_xxx_backward(...)).  owl.net uses them when a ReluUnit's only consumer is an
 * LRNUnit / max PoolingUnit (owl/owl/net/net.py:281-296 runs the mask as its own 12 B/element pass). */
int mnv_lrn_backward_relu(const float* bottom_data, const float* top_data, const float* scale,
                          const float* top_diff, float* bottom_diff, int local_size, float alpha,
                          float beta, int num_img, int channel, int width, int height,
                          mnv_stream_t stream);
/* Extension (fusion): ConvBackwardFilter + ConvBackwardBias in one call.  filter_diff is what
 * mnv_conv_backward_filter writes (same kernels, bit-identical); bias_diff[c] = sum of top_diff over images and pixels
 * is accumulated by the pre-pass that re-pitches top_diff for the tensor maps (odd plane sizes: every AlexNet /
 * GoogLeNet layer), so the separate 4 B-per-element reduction disappears; with 16-byte-pitched planes or without a
 * workspace it falls back to mnv_conv_backward_bias.  owl.net's ConvConnection.bp asks for both gradients at once
 * (owl/owl/net/net.py:700-716 issues two ops). */
int mnv_conv_backward_filter_bias(const float* bottom, const float* top_diff, float* filter_diff,
                                  float* bias_diff, int num_images, int bottom_num_channels,
                                  int top_num_channels, int bottom_height, int bottom_width,
                                  int pad_height, int pad_width, int stride_vertical,
                                  int stride_horizontal, int filter_height, int filter_width,
                                  void* workspace, size_t workspace_bytes, mnv_stream_t stream);
/* Extension (recompute instead of re-read): LRN without the `scale` array.  The forward pass writes only
 * `res` (8 B per element instead of 12); the backward pass reads only bottom and top_diff (12 B instead of 20) and
 * recomputes scale and top in registers with the forward pass's exact operation sequence, so bottom_diff is
 * bit-identical to mnv_lrn_backward fed with the stored arrays.  relu != 0 additionally applies the ReLU-backward mask
 * (bottom > 0).  Window 5 only (AlexNet, GoogLeNet); other windows return MNV_EUNSUPPORTED and the caller keeps the
 * three-array form.  owl.net's LRNUnit uses the pair when both passes run on this backend. */
int mnv_lrn_forward_lite(const float* bottom, float* res, int local_size, float alpha, float beta,
                         int num_img, int channel, int width, int height, mnv_stream_t stream);
int mnv_lrn_backward_lite(const float* bottom_data, const float* top_diff, float* bottom_diff,
                          int local_size, float alpha, float beta, int num_img, int channel,
                          int width, int height, int relu, mnv_stream_t stream);
int mnv_max_pooling_backward_relu(const float* bottom, const float* top, const float* top_diff,
                                  float* bottom_diff, int num_images, int num_channels,
                                  int bottom_height, int bottom_width, int stride_vertical,
                                  int stride_horizontal, int window_height, int window_width,
                                  int pad_height, int pad_width, mnv_stream_t stream);

/* Extension (remember instead of recompute): 3x3 / stride 2 / pad 0 max pooling (AlexNet, GoogLeNet) and
 * 3x3 / stride 1 / pad 1 max pooling (GoogLeNet's inception pools; mnv_max_pooling_idx_supported says which) whose
 * forward pass also writes one byte per pooled element -- the window position kh*3+kw of the first maximum in scan
 * order, 255 for a window of NaNs -- so that the backward pass reads (top_diff, idx) = 5 B per pooled element instead of
 * re-reading the whole bottom (4 B per INPUT element) and recomputing the arg-max.  bottom_diff is bit-identical to
 * mnv_max_pooling_backward.  relu_top != NULL folds ReLU backward in for a bottom that is a ReLU output: pass the
 * pooled top (the arg-max element passes iff its value, the window maximum, is > 0).  Other geometries return
 * MNV_EUNSUPPORTED and the caller keeps the recomputing pair. */
int mnv_max_pooling_idx_supported(int num_images, int num_channels, int bottom_height, int bottom_width,
                                  int stride_vertical, int stride_horizontal, int window_height,
                                  int window_width, int pad_height, int pad_width);   /* 1 / 0, no launch */
int mnv_max_pooling_forward_idx(const float* bottom, float* top, unsigned char* idx, int num_images,
                                int num_channels, int bottom_height, int bottom_width,
                                int stride_vertical, int stride_horizontal, int window_height,
                                int window_width, int pad_height, int pad_width, mnv_stream_t stream);
int mnv_max_pooling_backward_idx(const float* top_diff, const unsigned char* idx, const float* relu_top,
                                 float* bottom_diff, int num_images, int num_channels,
                                 int bottom_height, int bottom_width, int stride_vertical,
                                 int stride_horizontal, int window_height, int window_width,
                                 int pad_height, int pad_width, mnv_stream_t stream);

/* ---- the data layer's transform on the device -----------------------------
 * The reference converts every minibatch on the HOST (owl/owl/net/netio.py:300-311: uint8 Datum - mean image, random
 * crop, optional mirror, astype(float32)) and uploads 4 bytes per pixel through ArrayLoader on the default stream
 * (op/impl/cuda.cpp:592-597).  Here the stored uint8 images are uploaded (1 byte per pixel) and transformed by one
 * kernel:  dst[n][c][y][x] = (float(src[n][c][oy+y][ox+xs]) - mean[c][oy+y][ox+xs]) * scale,  xs = mirror ? crop_w-1-x : x.
 *   src   uint8 [N][C][src_h][src_w];   mean fp32 [C][src_h][src_w] or NULL (no subtraction);
 *   crop_mirror  int32 [N][3] = (oy, ox, mirror) per image, or NULL = (0, 0, 0);   dst fp32 [N][C][crop_h][crop_w].
 * Bit-exact against the host computation (one subtraction, one multiplication, both rounded to nearest). */
int mnv_image_transform_u8(const unsigned char* src, const float* mean, const int* crop_mirror, float* dst,
                           int num_images, int num_channels, int src_height, int src_width, int crop_height,
                           int crop_width, float scale, mnv_stream_t stream);

/* ---- channels-last twins shared by the convolution calls of one training step -------------------------
 * The im2col tensor maps that feed the tensor core read a channels-last copy of an NCHW activation.  One training step
 * wants each copy several times: forward and backward-filter read the bottom, backward-data and backward-filter read
 * top_diff.  A caller that keeps the arrays alive between those calls (owl.net does) may own the copy -- the "twin",
 * mnv_conv_twin_bytes() bytes, 256-byte aligned -- and pass it with an in/out state word: *state == 0: the call fills the
 * twin (instead of a workspace copy) and sets *state = 1 if it did; *state != 0: the twin is current, the pre-pass is
 * skipped.  twin == NULL: exactly the reference-shaped entry (copy in the workspace, once per call).  Results are
 * bit-identical either way.  mnv_conv_twin_wanted(): bit 0 = some direction of this convolution can use a bottom twin,
 * bit 1 = a top_diff twin (0: do not allocate; no launch).  mnv_conv_backward_filter_tw: bias_diff may be NULL; when it
 * is not, the bias sums ride on the pass that fills the top_diff twin (so ask for them in the call that fills it).
 * State word: bit 0 = the copy is current; bit 1 = the buffer's tail also holds per-(image, 128-pixel tile, channel) sums
 * of the array (left by mnv_relu_backward_tw), from which backward-filter folds bias_diff without reading top_diff. */
size_t mnv_conv_twin_bytes(int num_images, int num_channels, int height, int width);
int mnv_conv_twin_wanted(int num_images, int bottom_num_channels, int top_num_channels, int bottom_height,
                         int bottom_width, int pad_height, int pad_width, int stride_vertical, int stride_horizontal,
                         int filter_height, int filter_width);
int mnv_conv_forward_tw(const float* bottom, const float* filter, const float* bias, float* top, int num_images,
                        int bottom_num_channels, int top_num_channels, int bottom_height, int bottom_width,
                        int pad_height, int pad_width, int stride_vertical, int stride_horizontal, int filter_height,
                        int filter_width, int relu, float* bottom_twin, int* bottom_twin_state, void* workspace,
                        size_t workspace_bytes, mnv_stream_t stream);
int mnv_conv_backward_data_tw(const float* top_diff, const float* filter, float* bottom_diff, int num_images,
                              int bottom_num_channels, int top_num_channels, int bottom_height, int bottom_width,
                              int pad_height, int pad_width, int stride_vertical, int stride_horizontal,
                              int filter_height, int filter_width, float* top_diff_twin, int* top_diff_twin_state,
                              void* workspace, size_t workspace_bytes, mnv_stream_t stream);
int mnv_conv_backward_filter_tw(const float* bottom, const float* top_diff, float* filter_diff, float* bias_diff,
                                int num_images, int bottom_num_channels, int top_num_channels, int bottom_height,
                                int bottom_width, int pad_height, int pad_width, int stride_vertical,
                                int stride_horizontal, int filter_height, int filter_width, float* bottom_twin,
                                int* bottom_twin_state, float* top_diff_twin, int* top_diff_twin_state, void* workspace,
                                size_t workspace_bytes, mnv_stream_t stream);

/* ReLU backward (bottom_diff = top > 0 ? top_diff : 0, as mnv_relu_backward) that also fills the channels-last twin of
 * bottom_diff -- the top_diff twin of the convolution whose output the ReLU rectified -- and the channel sums behind it
 * (*twin_state = 3).  One pass of 16 B per element replaces ReLU backward (12 B) + the conversion (8 B).  twin == NULL:
 * the plain op. */
int mnv_relu_backward_tw(const float* top, const float* top_diff, float* bottom_diff, int num_images, int num_channels,
                         int height, int width, float* twin, int* twin_state, mnv_stream_t stream);

/* ---- explicit in-place forms ------------------------------------------------------------------
 * The entries above never alias an output with an input.  Two callers need to: the data-parallel gradient merge
 * (owl/net/merge.py: shard += peer's shard; the reference's `wgrad[upd_gpu] += wgrad[gid]`, owl/owl/net/trainer.py:131-135)
 * and the max-pooling-backward ReLU fallback.  These read the aliased buffer with ordinary (coherent) loads. */
int mnv_accumulate(float* acc, const float* x, size_t n, mnv_stream_t stream);          /* acc[i] += x[i] */
int mnv_relu_mask_inplace(float* dx, const float* x, size_t n, mnv_stream_t stream);    /* dx[i] = x[i] > 0 ? dx[i] : 0 */

/* ---- fused momentum-SGD update (owl/net/net.py:252-256) ----------------
 * delta = mom*delta - (lr/batch)*grad - (lr*wd)*w ; w += delta    (20 B/param instead of the
 * reference's ten-op chain).  In place on w and delta. */
int mnv_sgd_momentum_update(float* w, float* delta, const float* grad, size_t n, float momentum,
                            float lr_over_batch, float lr_times_wd, mnv_stream_t stream);

/* The same update for `count` tensors in one launch (host array of descriptors, copied into the kernel's parameters):
 * a net's many small parameter tensors make one launch per tensor launch-bound.  Bit-identical to `count` single calls. */
typedef struct {
  float* w;
  float* delta;
  const float* grad;
  size_t n;
  float lr_over_batch;
  float lr_times_wd;
} mnv_sgd_tensor_t;
int mnv_sgd_momentum_update_multi(const mnv_sgd_tensor_t* tensors, int count, float momentum, mnv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MNV_H_ */
