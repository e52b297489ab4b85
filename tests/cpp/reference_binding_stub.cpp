// reference_binding_stub.cpp -- the reference-side binding of INTEGRATION.md, compiled (syntax only) against the REFERENCE'S
// OWN headers under /root/reference/minerva (with HAS_CUDA, as the reference's GPU build sees them): each function below is
// the replacement body of a shim in minerva/op/impl/cuda.cpp -- same name, same signature as declared in the reference's
// minerva/op/impl/cuda.h:10-43 -- calling the C ABI of include/mnv.h instead of CudaPerform*.  tests/test_abi_cpu.py runs
//   g++ -std=c++11 -fsyntax-only -DHAS_CUDA -I/root/reference/minerva -Ioracle/shim -Iinclude -I/usr/local/cuda/include ...
// when /root/reference is present (TEST INFRASTRUCTURE; nothing here is linked into the product).
//
// The reference's Context carries a stream plus cuBLAS / cuDNN handles (op/context.h:28-37); the handles are simply unused.
// The kernel workspace, which the reference does not have, rides on a derived context that GpuDevice::DoExecute creates
// (device.cpp:214-220) -- Context is polymorphic, so no reference header changes at all.
#include <dmlc/logging.h>
#include "op/impl/cuda.h"
#include "mnv.h"

namespace minerva {

struct B200Context : public Context {        // what the re-pointed GpuDevice::DoExecute passes as `ctx`
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
};

namespace cuda {
namespace {
inline void Workspace(const Context& c, void** ws, size_t* bytes) {
  const B200Context* b = dynamic_cast<const B200Context*>(&c);
  *ws = b ? b->workspace : nullptr;           // NULL / 0 is legal: every entry has a scratch-free schedule
  *bytes = b ? b->workspace_bytes : 0;
}
}  // namespace
#define MNV_CALL(expr) CHECK_EQ((expr), MNV_OK) << #expr   /* keeps the reference's fatal-on-error behaviour */

void Arithmetic(const DataList& inputs, const DataList& outputs, ArithmeticClosure& closure, const Context& context) {
  CHECK_EQ(inputs.size(), 2) << "Arithmetic takes 2 inputs";
  CHECK_EQ(outputs.size(), 1) << "Arithmetic takes 1 output";
  float *left = inputs[0].data_, *right = inputs[1].data_, *res = outputs[0].data_;
  size_t size = outputs[0].size_.Prod();
  switch (closure.type) {
    case ArithmeticType::kAdd: MNV_CALL(mnv_add(left, right, res, size, context.stream)); break;
    case ArithmeticType::kSub: MNV_CALL(mnv_sub(left, right, res, size, context.stream)); break;
    case ArithmeticType::kMult: MNV_CALL(mnv_dot_mult(left, right, res, size, context.stream)); break;
    case ArithmeticType::kDiv: MNV_CALL(mnv_dot_div(left, right, res, size, context.stream)); break;
  }
}

void MatMult(const DataList& inputs, const DataList& outputs, MatMultClosure&, const Context& context) {
  CHECK_EQ(inputs.size(), 2) << "(matmult) #inputs is wrong!";
  CHECK_EQ(outputs.size(), 1) << "(matmult) #outputs is wrong!";
  void* ws; size_t wsb;
  Workspace(context, &ws, &wsb);
  int m = inputs[0].size_[0], k = inputs[0].size_[1], n = outputs[0].size_[1];
  MNV_CALL(mnv_matmult(inputs[0].data_, inputs[1].data_, outputs[0].data_, m, n, k, ws, wsb, context.stream));
}

void ConvForward(const DataList& inputs, const DataList& outputs, ConvForwardClosure& closure, const Context& context) {
  CHECK_EQ(inputs.size(), 3) << "(conv forward) #inputs wrong";
  CHECK_EQ(outputs.size(), 1) << "(conv forward) #outputs wrong";
  auto &bottom = inputs[0], &filter = inputs[1], &bias = inputs[2], &top = outputs[0];
  void* ws; size_t wsb;
  Workspace(context, &ws, &wsb);
  MNV_CALL(mnv_conv_forward(bottom.data_, filter.data_, bias.data_, top.data_, bottom.size_[3], bottom.size_[2], top.size_[2],
                            bottom.size_[1], bottom.size_[0], closure.pad_height, closure.pad_width, closure.stride_vertical,
                            closure.stride_horizontal, filter.size_[1], filter.size_[0], ws, wsb, context.stream));
}

void ConvBackwardData(const DataList& inputs, const DataList& outputs, ConvBackwardDataClosure& closure, const Context& context) {
  CHECK_EQ(inputs.size(), 2) << "(conv backward data) #inputs wrong";
  auto &top_diff = inputs[0], &filter = inputs[1], &bottom_diff = outputs[0];
  void* ws; size_t wsb;
  Workspace(context, &ws, &wsb);
  MNV_CALL(mnv_conv_backward_data(top_diff.data_, filter.data_, bottom_diff.data_, top_diff.size_[3], bottom_diff.size_[2],
                                  top_diff.size_[2], bottom_diff.size_[1], bottom_diff.size_[0], closure.pad_height,
                                  closure.pad_width, closure.stride_vertical, closure.stride_horizontal, filter.size_[1],
                                  filter.size_[0], ws, wsb, context.stream));
}

void PoolingForward(const DataList& inputs, const DataList& outputs, PoolingForwardClosure& closure, const Context& context) {
  CHECK_EQ(inputs.size(), 1) << "(pooling forward) #inputs wrong";
  auto &bottom = inputs[0], &top = outputs[0];
  auto fn = closure.algorithm == PoolingInfo::Algorithm::kMax ? mnv_max_pooling_forward : mnv_average_pooling_forward;
  MNV_CALL(fn(bottom.data_, top.data_, bottom.size_[3], bottom.size_[2], bottom.size_[1], bottom.size_[0], closure.stride_vertical,
              closure.stride_horizontal, closure.height, closure.width, closure.pad_height, closure.pad_width, context.stream));
}

void LRNForward(const DataList& inputs, const DataList& outputs, LRNForwardClosure& closure, const Context& context) {
  CHECK_EQ(inputs.size(), 2) << "(LRNForward) #inputs is wrong!";
  const Scale& s = closure.data_shape;
  MNV_CALL(mnv_lrn_forward(inputs[0].data_, inputs[1].data_, outputs[0].data_, closure.local_size, closure.alpha, closure.beta,
                           s[3], s[2], s[1], s[0], context.stream));
}

void Fill(const DataList& outputs, FillClosure& closure, const Context& context) {
  CHECK_EQ(outputs.size(), 1) << "(fill) #outputs wrong";
  MNV_CALL(mnv_fill(outputs[0].data_, outputs[0].size_.Prod(), closure.val, context.stream));
}

// the definitions above ARE the functions the reference declares (same types, or these initialisations do not compile)
void (*const kCheckArithmetic)(const DataList&, const DataList&, ArithmeticClosure&, const Context&) = &Arithmetic;
void (*const kCheckMatMult)(const DataList&, const DataList&, MatMultClosure&, const Context&) = &MatMult;
void (*const kCheckConvForward)(const DataList&, const DataList&, ConvForwardClosure&, const Context&) = &ConvForward;
void (*const kCheckConvBackwardData)(const DataList&, const DataList&, ConvBackwardDataClosure&, const Context&) = &ConvBackwardData;
void (*const kCheckPoolingForward)(const DataList&, const DataList&, PoolingForwardClosure&, const Context&) = &PoolingForward;
void (*const kCheckLRNForward)(const DataList&, const DataList&, LRNForwardClosure&, const Context&) = &LRNForward;
void (*const kCheckFill)(const DataList&, FillClosure&, const Context&) = &Fill;

// ... and they are the ONLY functions of these names: decltype(&f) is ill-formed for an overload set, so a definition whose
// signature drifted from the reference's declaration (creating a second overload) fails here
static_assert(sizeof(decltype(&Arithmetic)) && sizeof(decltype(&MatMult)) && sizeof(decltype(&ConvForward)) &&
              sizeof(decltype(&ConvBackwardData)) && sizeof(decltype(&PoolingForward)) && sizeof(decltype(&LRNForward)) &&
              sizeof(decltype(&Fill)), "reference declarations and replacement definitions agree");

}  // namespace cuda
}  // namespace minerva
