// Host shims: unpack DataList / closure into the flat C ABI of include/mnv.h.  Same role, argument order
// and arity checks as the reference's minerva/op/impl/cuda.cpp:22-655; what changed underneath is that
// every call is one enqueue-only launch of an sm_100a kernel on ctx.stream (no cuBLAS / cuDNN handles,
// no per-call descriptor churn, cudaMalloc or stream sync).  A non-zero return code is fatal, like the
// reference's CHECK / CUDA_CALL macros (minerva/common/cuda_utils.h:70-98).
#include <chrono>
#include <sstream>
#include <stdexcept>
#include "mnv.h"
#include "op/hotpath.h"

namespace minerva {
namespace cuda {
namespace {

[[noreturn]] void Fatal(const std::string& msg) { throw std::runtime_error(msg); }
#define MNV_REQUIRE(cond, msg) do { if (!(cond)) Fatal(std::string("Check failed: " #cond " ") + (msg)); } while (0)
void Ok(int rc, const char* what) {
  if (rc != MNV_OK) {
    std::ostringstream os;
    os << what << " failed with code " << rc;
    Fatal(os.str());
  }
}
unsigned WallSeed() {   // the reference seeds its generators from the wall clock (cuda.cpp:601,606)
  return static_cast<unsigned>(std::chrono::system_clock::now().time_since_epoch().count());
}
size_t Len(const DataShard& d) { return static_cast<size_t>(d.size_.Prod()); }

}  // namespace

void Run(const DataList& in, const DataList& out, ArithmeticClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 2, "Arithmetic takes 2 inputs");
  MNV_REQUIRE(out.size() == 1, "Arithmetic takes 1 output");
  float *l = in[0].data_, *r = in[1].data_, *res = out[0].data_;
  size_t n = Len(out[0]);
  switch (c.type) {
    case ArithmeticType::kAdd: Ok(mnv_add(l, r, res, n, ctx.stream), "mnv_add"); break;
    case ArithmeticType::kSub: Ok(mnv_sub(l, r, res, n, ctx.stream), "mnv_sub"); break;
    case ArithmeticType::kMult: Ok(mnv_dot_mult(l, r, res, n, ctx.stream), "mnv_dot_mult"); break;
    case ArithmeticType::kDiv: Ok(mnv_dot_div(l, r, res, n, ctx.stream), "mnv_dot_div"); break;
  }
}

void Run(const DataList& in, const DataList& out, ArithmeticConstClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1, "(arithmetic const) #inputs is wrong!");
  MNV_REQUIRE(out.size() == 1, "(arithmetic const) #outputs is wrong!");
  float *x = in[0].data_, *y = out[0].data_;
  size_t n = Len(in[0]);
  switch (c.type) {
    case ArithmeticType::kAdd: Ok(mnv_const_add(x, y, c.val, n, ctx.stream), "mnv_const_add"); break;
    case ArithmeticType::kSub:
      if (c.side == 0) Ok(mnv_left_const_sub(x, y, c.val, n, ctx.stream), "mnv_left_const_sub");
      else Ok(mnv_const_add(x, y, -c.val, n, ctx.stream), "mnv_const_add");
      break;
    case ArithmeticType::kMult: Ok(mnv_scale(x, y, n, c.val, ctx.stream), "mnv_scale"); break;
    case ArithmeticType::kDiv:
      if (c.side == 0) Ok(mnv_left_const_div(x, y, c.val, n, ctx.stream), "mnv_left_const_div");
      else Ok(mnv_const_div(x, y, c.val, n, ctx.stream), "mnv_const_div");   // IEEE division, bit-equal to basic.cpp:99-103
      break;
  }
}

void Run(const DataList& in, const DataList& out, ElewiseClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "(elewise) #inputs/#outputs is wrong!");
  size_t n = Len(out[0]);
  switch (c.type) {
    case ElewiseType::kExp: Ok(mnv_elewise_exp(in[0].data_, out[0].data_, n, ctx.stream), "mnv_elewise_exp"); break;
    case ElewiseType::kLn: Ok(mnv_elewise_ln(in[0].data_, out[0].data_, n, ctx.stream), "mnv_elewise_ln"); break;
    case ElewiseType::kNegative: Ok(mnv_elewise_negative(in[0].data_, out[0].data_, n, ctx.stream), "mnv_elewise_negative"); break;
  }
}

void Run(const DataList& in, const DataList& out, MatMultClosure&, const Context& ctx) {
  MNV_REQUIRE(in.size() == 2 && out.size() == 1, "(matmult) #inputs/#outputs is wrong!");
  int m = in[0].size_[0], k = in[0].size_[1], n = out[0].size_[1];
  Ok(mnv_matmult(in[0].data_, in[1].data_, out[0].data_, m, n, k, ctx.workspace, ctx.workspace_bytes, ctx.stream), "mnv_matmult");
}

void Run(const DataList& in, const DataList& out, TransposeClosure&, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "(transpose) #inputs/#outputs is wrong!");
  Ok(mnv_transpose(in[0].data_, out[0].data_, in[0].size_[0], in[0].size_[1], ctx.stream), "mnv_transpose");
}

void Run(const DataList& in, const DataList& out, ReshapeClosure&, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "(reshape) #inputs/#outputs is wrong!");
  Ok(mnv_reshape(in[0].data_, out[0].data_, Len(in[0]) * sizeof(float), ctx.stream), "mnv_reshape");
}

void Run(const DataList& in, const DataList& out, NormArithmeticClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 2 && out.size() == 1, "NormArithmetic kernel wrong #input/#output");
  MNV_REQUIRE(in[0].size_.NumDims() == 2, "currently support 2D normalizee matrix only");
  MNV_REQUIRE(c.dims_to_replicate.NumDims() == 1, "currently do norm on one dimension only");
  int m = in[0].size_[0], n = in[0].size_[1];
  float *mat = in[0].data_, *vec = in[1].data_, *res = out[0].data_;
  typedef int (*Fn)(const float*, const float*, float*, int, int, mnv_stream_t);
  static const Fn kOnCol[4] = {mnv_norm_add_on_col, mnv_norm_sub_on_col, mnv_norm_mult_on_col, mnv_norm_div_on_col};
  static const Fn kOnRow[4] = {mnv_norm_add_on_row, mnv_norm_sub_on_row, mnv_norm_mult_on_row, mnv_norm_div_on_row};
  const Fn fn = (c.dims_to_replicate[0] == 0 ? kOnCol : kOnRow)[static_cast<int>(c.type)];
  Ok(fn(mat, vec, res, m, n, ctx.stream), "mnv_norm_*");
}

void Run(const DataList& in, const DataList& out, ReductionClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "Reduction kernel wrong #input/#output");
  MNV_REQUIRE(in[0].size_.NumDims() == 2, "currently support 2D reduction matrix only");
  MNV_REQUIRE(c.dims_to_reduce.NumDims() == 1, "currently do reduction on one dimension only");
  int m = in[0].size_[0], n = in[0].size_[1];
  bool col = c.dims_to_reduce[0] == 0, sum = c.type == ReductionType::kSum;
  int rc = col ? (sum ? mnv_reduction_sum_on_col : mnv_reduction_max_on_col)(in[0].data_, out[0].data_, m, n, ctx.stream)
               : (sum ? mnv_reduction_sum_on_row : mnv_reduction_max_on_row)(in[0].data_, out[0].data_, m, n, ctx.stream);
  Ok(rc, "mnv_reduction_*");
}

void Run(const DataList& in, const DataList& out, MaxIndexClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "MaxIndex kernel wrong #input/#output");
  MNV_REQUIRE(in[0].size_.NumDims() == 2, "currently support 2D MaxIndex matrix only");
  int m = in[0].size_[0], n = in[0].size_[1];
  Ok((c.dim == 0 ? mnv_max_index_on_col : mnv_max_index_on_row)(in[0].data_, out[0].data_, m, n, ctx.stream), "mnv_max_index_*");
}

// ---- activations: argument order (diff, top, bottom) as in narray_elewise.cpp:51-82 ------------------
#define MNV_ACT_FWD(Closure, fn)                                                                          \
  void Run(const DataList& in, const DataList& out, Closure&, const Context& ctx) {                       \
    MNV_REQUIRE(in.size() == 1 && out.size() == 1, #fn " #inputs/#outputs wrong");                        \
    Ok(fn(in[0].data_, out[0].data_, 1, 1, 1, in[0].size_.Prod(), ctx.stream), #fn);                      \
  }
#define MNV_ACT_BWD(Closure, fn)                                                                          \
  void Run(const DataList& in, const DataList& out, Closure&, const Context& ctx) {                       \
    MNV_REQUIRE(in.size() == 3 && out.size() == 1, #fn " #inputs/#outputs wrong");                        \
    Ok(fn(in[2].data_, in[1].data_, in[0].data_, out[0].data_, 1, 1, 1, in[0].size_.Prod(), ctx.stream), #fn); \
  }
MNV_ACT_FWD(SigmoidForwardClosure, mnv_sigmoid_forward)
MNV_ACT_FWD(ReluForwardClosure, mnv_relu_forward)
MNV_ACT_FWD(TanhForwardClosure, mnv_tanh_forward)
MNV_ACT_BWD(SigmoidBackwardClosure, mnv_sigmoid_backward)
MNV_ACT_BWD(ReluBackwardClosure, mnv_relu_backward)
MNV_ACT_BWD(TanhBackwardClosure, mnv_tanh_backward)

void Run(const DataList& in, const DataList& out, ActivationForwardClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "(activation forward) #inputs/#outputs wrong");
  const Scale& s = in[0].size_;
  auto fn = c.algorithm == ActivationAlgorithm::kSigmoid ? mnv_sigmoid_forward
            : c.algorithm == ActivationAlgorithm::kRelu ? mnv_relu_forward : mnv_tanh_forward;
  Ok(fn(in[0].data_, out[0].data_, s[3], s[2], s[1], s[0], ctx.stream), "mnv_*_forward");
}
void Run(const DataList& in, const DataList& out, ActivationBackwardClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 3 && out.size() == 1, "(activation backward) #inputs/#outputs wrong");
  const Scale& s = in[0].size_;
  auto fn = c.algorithm == ActivationAlgorithm::kSigmoid ? mnv_sigmoid_backward
            : c.algorithm == ActivationAlgorithm::kRelu ? mnv_relu_backward : mnv_tanh_backward;
  Ok(fn(in[2].data_, in[1].data_, in[0].data_, out[0].data_, s[3], s[2], s[1], s[0], ctx.stream), "mnv_*_backward");
}

// ---- convolution: inputs (bottom, filter, bias) / (top_diff, filter) / (top_diff, bottom) ---------------
void Run(const DataList& in, const DataList& out, ConvForwardClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 3 && out.size() == 1, "(conv forward) #inputs/#outputs wrong");
  const Scale &b = in[0].size_, &f = in[1].size_, &t = out[0].size_;
  Ok(mnv_conv_forward(in[0].data_, in[1].data_, in[2].data_, out[0].data_, b[3], b[2], t[2], b[1], b[0], c.pad_height,
                      c.pad_width, c.stride_vertical, c.stride_horizontal, f[1], f[0], ctx.workspace, ctx.workspace_bytes,
                      ctx.stream), "mnv_conv_forward");
}
void Run(const DataList& in, const DataList& out, ConvBackwardDataClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 2 && out.size() == 1, "(conv backward data) #inputs/#outputs wrong");
  const Scale &td = in[0].size_, &f = in[1].size_, &bd = out[0].size_;   // output has the bottom's shape (convolution.cpp:47)
  Ok(mnv_conv_backward_data(in[0].data_, in[1].data_, out[0].data_, td[3], bd[2], td[2], bd[1], bd[0], c.pad_height,
                            c.pad_width, c.stride_vertical, c.stride_horizontal, f[1], f[0], ctx.workspace,
                            ctx.workspace_bytes, ctx.stream), "mnv_conv_backward_data");
}
void Run(const DataList& in, const DataList& out, ConvBackwardFilterClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 2 && out.size() == 1, "(conv backward filter) #inputs/#outputs wrong");
  const Scale &td = in[0].size_, &b = in[1].size_, &fd = out[0].size_;
  Ok(mnv_conv_backward_filter(in[1].data_, in[0].data_, out[0].data_, td[3], b[2], td[2], b[1], b[0], c.pad_height,
                              c.pad_width, c.stride_vertical, c.stride_horizontal, fd[1], fd[0], ctx.workspace,
                              ctx.workspace_bytes, ctx.stream), "mnv_conv_backward_filter");
}
void Run(const DataList& in, const DataList& out, ConvBackwardBiasClosure&, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "(conv backward bias) #inputs/#outputs wrong");
  const Scale& td = in[0].size_;
  Ok(mnv_conv_backward_bias(in[0].data_, out[0].data_, td[3], td[2], td[1], td[0], ctx.workspace, ctx.workspace_bytes,
                            ctx.stream), "mnv_conv_backward_bias");
}

void Run(const DataList& in, const DataList& out, SoftmaxForwardClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "(softmax forward) #inputs/#outputs wrong");
  const Scale& s = in[0].size_;
  auto fn = c.algorithm == SoftmaxAlgorithm::kInstance ? mnv_instance_softmax_forward : mnv_channel_softmax_forward;
  Ok(fn(in[0].data_, out[0].data_, s[3], s[2], s[1], s[0], ctx.stream), "mnv_*_softmax_forward");
}
void Run(const DataList& in, const DataList& out, SoftmaxBackwardClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 2 && out.size() == 1, "(softmax backward) #inputs/#outputs wrong");
  const Scale& s = in[0].size_;
  auto fn = c.algorithm == SoftmaxAlgorithm::kInstance ? mnv_instance_softmax_backward : mnv_channel_softmax_backward;
  Ok(fn(in[0].data_, in[1].data_, out[0].data_, s[3], s[2], s[1], s[0], ctx.stream), "mnv_*_softmax_backward");
}

void Run(const DataList& in, const DataList& out, PoolingForwardClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "(pooling forward) #inputs/#outputs wrong");
  const Scale& b = in[0].size_;
  auto fn = c.algorithm == PoolingInfo::Algorithm::kMax ? mnv_max_pooling_forward : mnv_average_pooling_forward;
  Ok(fn(in[0].data_, out[0].data_, b[3], b[2], b[1], b[0], c.stride_vertical, c.stride_horizontal, c.height, c.width,
        c.pad_height, c.pad_width, ctx.stream), "mnv_*_pooling_forward");
}
void Run(const DataList& in, const DataList& out, PoolingBackwardClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 3 && out.size() == 1, "(pooling backward) #inputs/#outputs wrong");
  const Scale& b = in[2].size_;   // (diff, top, bottom)
  auto fn = c.algorithm == PoolingInfo::Algorithm::kMax ? mnv_max_pooling_backward : mnv_average_pooling_backward;
  Ok(fn(in[2].data_, in[1].data_, in[0].data_, out[0].data_, b[3], b[2], b[1], b[0], c.stride_vertical, c.stride_horizontal,
        c.height, c.width, c.pad_height, c.pad_width, ctx.stream), "mnv_*_pooling_backward");
}

// LRN: the closure carries the data shape; forward writes its `scale` INPUT in place (cuda.cpp:45-59)
void Run(const DataList& in, const DataList& out, LRNForwardClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 2 && out.size() == 1, "(LRNForward) #inputs/#outputs is wrong!");
  const Scale& s = c.data_shape;
  Ok(mnv_lrn_forward(in[0].data_, in[1].data_, out[0].data_, c.local_size, c.alpha, c.beta, s[3], s[2], s[1], s[0], ctx.stream),
     "mnv_lrn_forward");
}
void Run(const DataList& in, const DataList& out, LRNBackwardClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 4 && out.size() == 1, "(LRNBackward) #inputs/#outputs is wrong!");
  const Scale& s = c.data_shape;
  Ok(mnv_lrn_backward(in[0].data_, in[1].data_, in[2].data_, in[3].data_, out[0].data_, c.local_size, c.alpha, c.beta, s[3],
                      s[2], s[1], s[0], ctx.stream), "mnv_lrn_backward");
}

// Concat / Slice on the last or second-to-last dimension: one strided-copy launch per input instead of one
// cuBLAS copy per image (cuda.cpp:80-155)
void Run(const DataList& in, const DataList& out, ConcatClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() > 1 && out.size() == 1, "(Concat) #inputs/#outputs is wrong!");
  size_t nd = in[0].size_.NumDims(), dim = static_cast<size_t>(c.catdim);
  MNV_REQUIRE(nd - dim <= 2, "(Concat) Currently only support concat on the last two dims!");
  size_t inner_unit = 1, outer = 1;
  for (size_t i = 0; i < dim; ++i) inner_unit *= out[0].size_[i];
  for (size_t i = dim + 1; i < nd; ++i) outer *= out[0].size_[i];
  size_t dst_stride = inner_unit * out[0].size_[dim], off = 0;
  for (const DataShard& d : in) {
    size_t inner = inner_unit * d.size_[dim];
    Ok(mnv_copy_strided(d.data_, out[0].data_ + off, inner, outer, inner, dst_stride, ctx.stream), "mnv_copy_strided");
    off += inner;
  }
}
void Run(const DataList& in, const DataList& out, SliceClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "(Slice) #inputs/#outputs is wrong!");
  size_t nd = in[0].size_.NumDims(), dim = static_cast<size_t>(c.slice_dim);
  MNV_REQUIRE(nd - dim <= 2, "(Slice) Currently only support slice on the last two dims!");
  size_t inner_unit = 1, outer = 1;
  for (size_t i = 0; i < dim; ++i) inner_unit *= in[0].size_[i];
  for (size_t i = dim + 1; i < nd; ++i) outer *= in[0].size_[i];
  Ok(mnv_copy_strided(in[0].data_ + inner_unit * c.st_off, out[0].data_, inner_unit * c.slice_count, outer,
                      inner_unit * in[0].size_[dim], inner_unit * c.slice_count, ctx.stream), "mnv_copy_strided");
}

void Run(const DataList& in, const DataList& out, SelectClosure& c, const Context& ctx) {
  MNV_REQUIRE(in.size() == 1 && out.size() == 1, "(Select) #inputs/#outputs is wrong!");
  MNV_REQUIRE(out[0].size_[1] == static_cast<int>(c.indices.size()), "(Select) index count mismatch");
  // indices are staged through the workspace (the reference passed a HOST pointer to its kernel, cuda_perform.cu:676)
  MNV_REQUIRE(ctx.workspace && ctx.workspace_bytes >= c.indices.size() * sizeof(int), "(Select) workspace too small");
  cudaError_t e = cudaMemcpyAsync(ctx.workspace, c.indices.data(), c.indices.size() * sizeof(int), cudaMemcpyHostToDevice, ctx.stream);
  MNV_REQUIRE(e == cudaSuccess, "(Select) index upload failed");
  Ok(mnv_select(out[0].data_, in[0].data_, static_cast<const int*>(ctx.workspace), c.indices.size(), in[0].size_[1],
                in[0].size_[0], ctx.stream), "mnv_select");
}

// ---- data generators (outputs only) ------------------------------------------------------------------
void Run(const DataList& out, ArrayLoaderClosure& c, const Context& ctx) {
  MNV_REQUIRE(out.size() == 1, "(array loader) #outputs wrong");
  MNV_REQUIRE(static_cast<bool>(c.data), "probably already executed");
  // on the op's stream (the reference used the default stream, cuda.cpp:595)
  cudaError_t e = cudaMemcpyAsync(out[0].data_, c.data.get(), Len(out[0]) * sizeof(float), cudaMemcpyDefault, ctx.stream);
  MNV_REQUIRE(e == cudaSuccess, "(array loader) copy failed");
  e = cudaStreamSynchronize(ctx.stream);   // the host buffer is released below
  MNV_REQUIRE(e == cudaSuccess, "(array loader) sync failed");
  c.data.reset();
}
void Run(const DataList& out, RandnClosure& c, const Context& ctx) {
  MNV_REQUIRE(out.size() == 1, "(normal) #outputs wrong");
  Ok(mnv_randn(out[0].data_, Len(out[0]), WallSeed(), c.mu, c.var, ctx.stream), "mnv_randn");
}
void Run(const DataList& out, RandBernoulliClosure& c, const Context& ctx) {
  MNV_REQUIRE(out.size() == 1, "(bernoulli) #outputs wrong");
  Ok(mnv_rand_bernoulli(out[0].data_, Len(out[0]), WallSeed(), c.p, ctx.stream), "mnv_rand_bernoulli");
}
void Run(const DataList& out, FillClosure& c, const Context& ctx) {
  MNV_REQUIRE(out.size() == 1, "(fill) #outputs wrong");
  Ok(mnv_fill(out[0].data_, Len(out[0]), c.val, ctx.stream), "mnv_fill");
}

}  // namespace cuda
}  // namespace minerva
