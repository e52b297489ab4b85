// The PhysicalOp / ComputeFn plug-in surface of Minerva's op layer, kept API-compatible so the NArray
// front end and the DAG scheduler dispatch to the B200 kernels unchanged:
//   ComputeFn::Execute(DataList const& in, DataList const& out, Context const&)   op/compute_fn.h:9-12
//   DataShard{float* data_, Scale const& size_}                                    op/data_shard.h:7-14
//   Context{impl_type, stream, ...}                                                op/context.h:11-37
//   closures (field names / order are used by aggregate-init in narray code)       op/closure.h:8-176
//   ComputeFnWithClosure / PhyDataGenFnWithClosure, FnBundle<Closure>::Call         op/physical_fn.h, impl/impl.h
//   the XxxOp classes with Name()                                                  op/physical_op.h:14-382
// Differences: Context carries a per-stream kernel workspace instead of cuBLAS/cuDNN handles; FnBundle
// is a template over an overload set instead of INSTALL_COMPUTE_FN macros; on ImplType::kBasic the
// product throws "no implementation" (north_star: no CPU fallback on the GPU path -- the CPU
// restatement lives in oracle/ as test infrastructure).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>
#include "common/scale.h"
#include "narray/convolution_info.h"

namespace minerva {

// ---- context / data ---------------------------------------------------------------------------
enum class ImplType { kNA = 0, kBasic, kMkl, kCuda };

struct Context {
  ImplType impl_type = ImplType::kNA;
  cudaStream_t stream = nullptr;
  void* workspace = nullptr;        // per-(device, stream) scratch owned by the device layer
  std::size_t workspace_bytes = 0;
  virtual ~Context() {}
};

struct DataShard {
  DataShard(float* data, Scale const& size) : data_(data), size_(size) {}
  float* const data_;
  Scale const& size_;
};
using DataList = std::vector<DataShard>;

class BasicFn {
 public:
  virtual std::string Name() const = 0;
  virtual ~BasicFn() {}
};
template <class T> struct ClosureTrait { T closure; };

class ComputeFn : public BasicFn {
 public:
  virtual void Execute(DataList const&, DataList const&, Context const&) = 0;
};

struct PhysicalOp {
  std::shared_ptr<ComputeFn> compute_fn;
  uint64_t device_id;
};

// ---- closures -----------------------------------------------------------------------------------
enum class ArithmeticType { kAdd = 0, kSub, kMult, kDiv };
enum class ElewiseType { kExp = 0, kLn, kNegative };
enum class ReductionType { kSum = 0, kMax };

struct ArrayLoaderClosure { std::shared_ptr<float> data; };
struct RandnClosure { float mu, var; };
struct RandBernoulliClosure { float p; };
struct FillClosure { float val; };
struct MatMultClosure {};
struct TransposeClosure {};
struct ReshapeClosure {};
struct ReductionClosure { ReductionType type; Scale dims_to_reduce; };
struct MaxIndexClosure { int dim; };
struct ElewiseClosure { ElewiseType type; };
struct SigmoidForwardClosure {};
struct SigmoidBackwardClosure {};
struct ReluForwardClosure {};
struct ReluBackwardClosure {};
struct TanhForwardClosure {};
struct TanhBackwardClosure {};
struct ArithmeticClosure { ArithmeticType type; };
struct ArithmeticConstClosure { ArithmeticType type; float val; int side; /* 0: const on the left */ };
struct NormArithmeticClosure { ArithmeticType type; Scale dims_to_replicate; };
template <int i> struct ConvClosure { int pad_height, pad_width, stride_vertical, stride_horizontal; };
typedef ConvClosure<0> ConvForwardClosure;
typedef ConvClosure<1> ConvBackwardDataClosure;
typedef ConvClosure<2> ConvBackwardFilterClosure;
struct ConvBackwardBiasClosure {};
template <int i> struct SoftmaxClosure { SoftmaxAlgorithm algorithm; };
typedef SoftmaxClosure<0> SoftmaxForwardClosure;
typedef SoftmaxClosure<1> SoftmaxBackwardClosure;
template <int i> struct ActivationClosure { ActivationAlgorithm algorithm; };
typedef ActivationClosure<0> ActivationForwardClosure;
typedef ActivationClosure<1> ActivationBackwardClosure;
template <int i> struct PoolingClosure {
  PoolingInfo::Algorithm algorithm;
  int height, width, stride_vertical, stride_horizontal, pad_height, pad_width;
};
typedef PoolingClosure<0> PoolingForwardClosure;
typedef PoolingClosure<1> PoolingBackwardClosure;
template <int i> struct LRNClosure { int local_size; float alpha, beta; Scale data_shape; };
typedef LRNClosure<0> LRNForwardClosure;
typedef LRNClosure<1> LRNBackwardClosure;
struct ConcatClosure { int catdim; };
struct SliceClosure { int slice_dim, st_off, slice_count; };
struct SelectClosure { std::vector<int> indices; };

// ---- dispatch -----------------------------------------------------------------------------------
[[noreturn]] inline void NoImplementation(const char* closure_name, ImplType t) {
  std::ostringstream os;
  os << "no implementation for " << closure_name << " on impl type " << static_cast<int>(t)
     << " (this build provides ImplType::kCuda only)";
  throw std::runtime_error(os.str());   // the reference LOG(FATAL)s (op/impl/bundle.h:14-17)
}

namespace cuda {   // host shims over the C ABI, op/impl/cuda.cpp
#define MNV_SHIM(C) void Run(const DataList&, const DataList&, C&, const Context&);
MNV_SHIM(ArithmeticClosure) MNV_SHIM(ArithmeticConstClosure) MNV_SHIM(MatMultClosure) MNV_SHIM(TransposeClosure)
MNV_SHIM(ReductionClosure) MNV_SHIM(NormArithmeticClosure) MNV_SHIM(MaxIndexClosure) MNV_SHIM(ReshapeClosure)
MNV_SHIM(ElewiseClosure) MNV_SHIM(SigmoidForwardClosure) MNV_SHIM(SigmoidBackwardClosure) MNV_SHIM(ReluForwardClosure)
MNV_SHIM(ReluBackwardClosure) MNV_SHIM(TanhForwardClosure) MNV_SHIM(TanhBackwardClosure) MNV_SHIM(ConvForwardClosure)
MNV_SHIM(ConvBackwardDataClosure) MNV_SHIM(ConvBackwardFilterClosure) MNV_SHIM(ConvBackwardBiasClosure)
MNV_SHIM(SoftmaxForwardClosure) MNV_SHIM(SoftmaxBackwardClosure) MNV_SHIM(ActivationForwardClosure)
MNV_SHIM(ActivationBackwardClosure) MNV_SHIM(PoolingForwardClosure) MNV_SHIM(PoolingBackwardClosure)
MNV_SHIM(LRNForwardClosure) MNV_SHIM(LRNBackwardClosure) MNV_SHIM(ConcatClosure) MNV_SHIM(SliceClosure)
MNV_SHIM(SelectClosure)
#undef MNV_SHIM
// data generators take outputs only
void Run(const DataList&, ArrayLoaderClosure&, const Context&);
void Run(const DataList&, RandnClosure&, const Context&);
void Run(const DataList&, RandBernoulliClosure&, const Context&);
void Run(const DataList&, FillClosure&, const Context&);
}  // namespace cuda

template <typename C> class FnBundle {
 public:
  static void Call(const DataList& in, const DataList& out, C& c, const Context& ctx) {
    if (ctx.impl_type == ImplType::kCuda) cuda::Run(in, out, c, ctx);
    else NoImplementation(typeid(C).name(), ctx.impl_type);
  }
  static void Call(const DataList& out, C& c, const Context& ctx) {
    if (ctx.impl_type == ImplType::kCuda) cuda::Run(out, c, ctx);
    else NoImplementation(typeid(C).name(), ctx.impl_type);
  }
};

template <typename Closure> class ComputeFnWithClosure : public ComputeFn, public ClosureTrait<Closure> {
 public:
  void Execute(const DataList& inputs, const DataList& outputs, const Context& context) {
    FnBundle<Closure>::Call(inputs, outputs, ClosureTrait<Closure>::closure, context);
  }
};
template <typename Closure> class PhyDataGenFnWithClosure : public ComputeFn, public ClosureTrait<Closure> {
 public:
  void Execute(const DataList&, const DataList& outputs, const Context& context) {
    FnBundle<Closure>::Call(outputs, ClosureTrait<Closure>::closure, context);
  }
};

// ---- the op classes (names as the reference prints them, op/physical_op.h) ------------------------
#define MNV_OP(Cls, Base, Closure, NameExpr) \
  class Cls : public Base<Closure> { public: std::string Name() const { return NameExpr; } };
MNV_OP(ArrayLoaderOp, PhyDataGenFnWithClosure, ArrayLoaderClosure, ":array loader")
MNV_OP(RandnOp, PhyDataGenFnWithClosure, RandnClosure, ":normal")
MNV_OP(RandBernoulliOp, PhyDataGenFnWithClosure, RandBernoulliClosure, ":bernoulli")
MNV_OP(FillOp, PhyDataGenFnWithClosure, FillClosure, ":const")
MNV_OP(MatMultOp, ComputeFnWithClosure, MatMultClosure, "*")
MNV_OP(TransOp, ComputeFnWithClosure, TransposeClosure, "trans")
MNV_OP(ReductionOp, ComputeFnWithClosure, ReductionClosure, closure.type == ReductionType::kSum ? "sum" : "max")
MNV_OP(MaxIndexOp, ComputeFnWithClosure, MaxIndexClosure, "max index")
MNV_OP(ReshapeOp, ComputeFnWithClosure, ReshapeClosure, "reshape")
MNV_OP(ElewiseOp, ComputeFnWithClosure, ElewiseClosure,
       closure.type == ElewiseType::kExp ? "exp" : closure.type == ElewiseType::kLn ? "ln" : "-")
MNV_OP(ArithmeticOp, ComputeFnWithClosure, ArithmeticClosure,
       closure.type == ArithmeticType::kAdd ? "+" : closure.type == ArithmeticType::kSub ? "-"
       : closure.type == ArithmeticType::kMult ? ".*" : "./")
MNV_OP(ArithmeticConstOp, ComputeFnWithClosure, ArithmeticConstClosure, "arithmetic const")
MNV_OP(NormArithmeticOp, ComputeFnWithClosure, NormArithmeticClosure, "norm arithmetic")
MNV_OP(SigmoidForwardOp, ComputeFnWithClosure, SigmoidForwardClosure, "sigmoid forward")
MNV_OP(SigmoidBackwardOp, ComputeFnWithClosure, SigmoidBackwardClosure, "sigmoid backward")
MNV_OP(ReluForwardOp, ComputeFnWithClosure, ReluForwardClosure, "relu forward")
MNV_OP(ReluBackwardOp, ComputeFnWithClosure, ReluBackwardClosure, "relu backward")
MNV_OP(TanhForwardOp, ComputeFnWithClosure, TanhForwardClosure, "tanh forward")
MNV_OP(TanhBackwardOp, ComputeFnWithClosure, TanhBackwardClosure, "tanh backward")
MNV_OP(ConvForwardOp, ComputeFnWithClosure, ConvForwardClosure, "conv ff")
MNV_OP(ConvBackwardDataOp, ComputeFnWithClosure, ConvBackwardDataClosure, "conv bp data")
MNV_OP(ConvBackwardFilterOp, ComputeFnWithClosure, ConvBackwardFilterClosure, "conv bp filter")
MNV_OP(ConvBackwardBiasOp, ComputeFnWithClosure, ConvBackwardBiasClosure, "conv bp bias")
MNV_OP(SoftmaxForwardOp, ComputeFnWithClosure, SoftmaxForwardClosure,
       closure.algorithm == SoftmaxAlgorithm::kInstance ? "instance softmax ff" : "channel softmax ff")
MNV_OP(SoftmaxBackwardOp, ComputeFnWithClosure, SoftmaxBackwardClosure,
       closure.algorithm == SoftmaxAlgorithm::kInstance ? "instance softmax bp" : "channel softmax bp")
MNV_OP(ActivationForwardOp, ComputeFnWithClosure, ActivationForwardClosure,
       closure.algorithm == ActivationAlgorithm::kSigmoid ? "sigmoid ff"
       : closure.algorithm == ActivationAlgorithm::kRelu ? "relu ff" : "tanh ff")
MNV_OP(ActivationBackwardOp, ComputeFnWithClosure, ActivationBackwardClosure,
       closure.algorithm == ActivationAlgorithm::kSigmoid ? "sigmoid bp"
       : closure.algorithm == ActivationAlgorithm::kRelu ? "relu bp" : "tanh bp")
MNV_OP(PoolingForwardOp, ComputeFnWithClosure, PoolingForwardClosure,
       closure.algorithm == PoolingInfo::Algorithm::kMax ? "max pooling ff" : "average pooling ff")
MNV_OP(PoolingBackwardOp, ComputeFnWithClosure, PoolingBackwardClosure,
       closure.algorithm == PoolingInfo::Algorithm::kMax ? "max pooling bp" : "average pooling bp")
MNV_OP(LRNForwardOp, ComputeFnWithClosure, LRNForwardClosure, "LRN Forward")
MNV_OP(LRNBackwardOp, ComputeFnWithClosure, LRNBackwardClosure, "LRN Backward")
MNV_OP(ConcatOp, ComputeFnWithClosure, ConcatClosure, "Concat")
MNV_OP(SliceOp, ComputeFnWithClosure, SliceClosure, "Slice")
MNV_OP(SelectOp, ComputeFnWithClosure, SelectClosure, "Select")
#undef MNV_OP

}  // namespace minerva
