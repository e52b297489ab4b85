// ref_shim.cpp -- C entry points over the REFERENCE's own CPU op implementations
// (minerva/op/impl/basic.cpp), compiled from the sources where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/libminerva_ref.so.  Used to (a) pin oracle/mnv_oracle.c
// bit-for-bit and (b) time the reference CPU path in bench.py.  TEST INFRASTRUCTURE ONLY.
// No reference source is copied into this repository; this file only calls its public functions.
#include <vector>
#include "op/impl/basic.h"
#include "op/closure.h"

using namespace minerva;

namespace {
DataList One(float* p, const Scale& s) { return DataList{DataShard(p, s)}; }
}

extern "C" {

// type: 0 add 1 sub 2 mult 3 div (ArithmeticType, op/closure.h)
void ref_arithmetic(int type, float* a, float* b, float* c, int n) {
  Scale s{n};
  ArithmeticClosure cl{static_cast<ArithmeticType>(type)};
  DataList in{DataShard(a, s), DataShard(b, s)};
  basic::Arithmetic(in, One(c, s), cl);
}
void ref_arithmetic_const(int type, int side, float val, float* in, float* out, int n) {
  Scale s{n};
  ArithmeticConstClosure cl{static_cast<ArithmeticType>(type), val, side};
  basic::ArithmeticConst(One(in, s), One(out, s), cl);
}
// type: 0 exp 1 ln 2 negative
void ref_elewise(int type, float* in, float* out, int n) {
  Scale s{n};
  ElewiseClosure cl{static_cast<ElewiseType>(type)};
  basic::Elewise(One(in, s), One(out, s), cl);
}
void ref_matmult(float* a, float* b, float* c, int m, int n, int k) {
  Scale sa{m, k}, sb{k, n}, sc{m, n};
  MatMultClosure cl;
  DataList in{DataShard(a, sa), DataShard(b, sb)};
  basic::MatMult(in, One(c, sc), cl);
}
// a is {m,n}; c is {n,m}
void ref_transpose(float* a, float* c, int m, int n) {
  Scale sa{m, n}, sc{n, m};
  TransposeClosure cl;
  basic::Transpose(One(a, sa), One(c, sc), cl);
}
// reduce a {m,n} matrix over dimension `dim` (0 or 1); is_max selects kMax
void ref_reduction(int is_max, int dim, float* in, float* out, int m, int n) {
  Scale si{m, n};
  Scale so = dim == 0 ? Scale{1, n} : Scale{m, 1};
  ReductionClosure cl{is_max ? ReductionType::kMax : ReductionType::kSum, Scale{dim}};
  basic::Reduction(One(in, si), One(out, so), cl);
}
void ref_max_index(int dim, float* in, float* out, int m, int n) {
  Scale si{m, n};
  Scale so = dim == 0 ? Scale{1, n} : Scale{m, 1};
  MaxIndexClosure cl{dim};
  basic::MaxIndex(One(in, si), One(out, so), cl);
}
// dims_to_replicate = {dim}: dim 0 -> vec is {1,n}; dim 1 -> vec is {m,1}
void ref_norm_arithmetic(int type, int dim, float* mat, float* vec, float* res, int m, int n) {
  Scale sm{m, n};
  Scale sv = dim == 0 ? Scale{1, n} : Scale{m, 1};
  NormArithmeticClosure cl{static_cast<ArithmeticType>(type), Scale{dim}};
  DataList in{DataShard(mat, sm), DataShard(vec, sv)};
  basic::NormArithmetic(in, One(res, sm), cl);
}
void ref_sigmoid_forward(float* x, float* y, int n) {
  Scale s{n};
  SigmoidForwardClosure cl;
  basic::SigmoidForward(One(x, s), One(y, s), cl);
}
void ref_relu_forward(float* x, float* y, int n) {
  Scale s{n};
  ReluForwardClosure cl;
  basic::ReluForward(One(x, s), One(y, s), cl);
}
void ref_tanh_forward(float* x, float* y, int n) {
  Scale s{n};
  TanhForwardClosure cl;
  basic::TanhForward(One(x, s), One(y, s), cl);
}
// {W,H,C,N}; the reference CPU path normalises over dimension 0 only (basic.cpp:231-232)
void ref_softmax_forward(float* x, float* y, int w, int h, int c, int n) {
  Scale s{w, h, c, n};
  SoftmaxForwardClosure cl{SoftmaxAlgorithm::kInstance};
  basic::SoftmaxForward(One(x, s), One(y, s), cl);
}
void ref_fill(float* dst, int n, float val) {
  Scale s{n};
  FillClosure cl{val};
  basic::Fill(One(dst, s), cl);
}
// ScaleRange::Flatten (common/scale.cpp:182-191) for the unittest_scale goldens
long ref_flatten(const int* dims, const int* idx, int nd) {
  Scale d(std::vector<int>(dims, dims + nd));
  Scale i(std::vector<int>(idx, idx + nd));
  return static_cast<long>(ScaleRange::MakeRangeFromOrigin(d).Flatten(i));
}

}  // extern "C"
