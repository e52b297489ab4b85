// C++ test of the drop-in surface, written like the reference's own gtest cases (tests/unittest_*.cpp) but
// at the plug-in boundary: ops are built as `XxxOp` objects with closures, executed through
// ComputeFn::Execute(DataList, DataList, Context) on a GpuDevice (streams + pooled store + workspace), and
// compared with the CPU oracle (linked as test infrastructure).  Prints "PASS <n> checks" and exits 0.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <vector>
#include "device/gpu_device.h"
#include "op/physical_op.h"

extern "C" {   // oracle/mnv_oracle.c
void orc_add(const float*, const float*, float*, size_t);
void orc_dot_div(const float*, const float*, float*, size_t);
void orc_const_div(const float*, float*, float, size_t);
void orc_relu_forward(const float*, float*, size_t);
void orc_relu_backward(const float*, const float*, const float*, float*, size_t);
void orc_matmult(const float*, const float*, float*, int, int, int);
void orc_norm_on_row(int, const float*, const float*, float*, int, int);
void orc_reduction_on_row(int, const float*, float*, int, int);
void orc_max_index_on_col(const float*, float*, int, int);
void orc_conv_forward(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int);
void orc_conv_backward_data(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int);
void orc_conv_backward_filter(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int);
void orc_max_pooling_forward(const float*, float*, int, int, int, int, int, int, int, int, int, int);
void orc_instance_softmax_forward(const float*, float*, int, int, int, int);
}

using namespace minerva;

static int g_checks = 0, g_fail = 0;
static std::mt19937 g_rng(0x5EED);

struct Arr {   // a device buffer from the pooled store + its Scale, like a PhysicalData + its storage
  Scale size;
  float* dev;
  uint64_t id;
};
static uint64_t g_next_id = 1;

static Arr Make(GpuDevice& d, Scale s, const std::vector<float>* host = nullptr) {
  Arr a{s, nullptr, g_next_id++};
  a.dev = d.data_store().CreateData(a.id, static_cast<size_t>(s.Prod()) * sizeof(float));
  if (host) cudaMemcpy(a.dev, host->data(), host->size() * sizeof(float), cudaMemcpyHostToDevice);
  return a;
}
static std::vector<float> Get(const Arr& a) {
  std::vector<float> h(a.size.Prod());
  cudaMemcpy(h.data(), a.dev, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
  return h;
}
static std::vector<float> Randn(size_t n, float sd = 1.f) {
  std::normal_distribution<float> dist(0.f, sd);
  std::vector<float> v(n);
  for (auto& x : v) x = dist(g_rng);
  return v;
}
static void ExpectBits(const std::vector<float>& got, const std::vector<float>& want, const char* what) {
  ++g_checks;
  if (got.size() != want.size() || std::memcmp(got.data(), want.data(), got.size() * sizeof(float)) != 0) {
    ++g_fail;
    std::printf("FAIL (bit-exact) %s\n", what);
  }
}
static void ExpectNormRel(const std::vector<float>& got, const std::vector<float>& want, double tol, const char* what) {
  ++g_checks;
  double num = 0, den = 0;
  for (size_t i = 0; i < got.size(); ++i) { double d = double(got[i]) - want[i]; num += d * d; den += double(want[i]) * want[i]; }
  double err = std::sqrt(num / (den > 0 ? den : 1));
  if (!(err < tol)) { ++g_fail; std::printf("FAIL %s: norm-rel %.3e >= %.1e\n", what, err, tol); }
}
template <class Op> static PhysicalOp Wrap(Op* op) { return PhysicalOp{std::shared_ptr<ComputeFn>(op), 0}; }

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { std::printf("no CUDA device\n"); return 2; }
  GpuDevice dev(0);
  Scale s5{2, 3, 4, 5, 6};   // the reference tests' shape (tests/unittest_arithmetic.cpp:10)
  const size_t n = s5.Prod();

  {  // Arithmetic + and ./ -- bit-exact (unittest_arithmetic.cpp:7-145)
    auto ha = Randn(n), hb = Randn(n, 5.f);
    Arr a = Make(dev, s5, &ha), b = Make(dev, s5, &hb), c = Make(dev, s5);
    std::vector<float> want(n);
    auto* op = new ArithmeticOp();
    op->closure = {ArithmeticType::kAdd};
    PhysicalOp po = Wrap(op);
    dev.DoExecute({DataShard(a.dev, a.size), DataShard(b.dev, b.size)}, {DataShard(c.dev, c.size)}, po, 0);
    orc_add(ha.data(), hb.data(), want.data(), n);
    ExpectBits(Get(c), want, "ArithmeticOp +");
    if (po.compute_fn->Name() != "+") { ++g_fail; std::printf("FAIL Name()\n"); }
    op->closure = {ArithmeticType::kDiv};
    dev.DoExecute({DataShard(a.dev, a.size), DataShard(b.dev, b.size)}, {DataShard(c.dev, c.size)}, po, 1);
    orc_dot_div(ha.data(), hb.data(), want.data(), n);
    ExpectBits(Get(c), want, "ArithmeticOp ./");
    // ArithmeticConst x / v with side = 1 (right const): true division
    auto* cop = new ArithmeticConstOp();
    cop->closure = {ArithmeticType::kDiv, 3.7f, 1};
    PhysicalOp pc = Wrap(cop);
    dev.DoExecute({DataShard(a.dev, a.size)}, {DataShard(c.dev, c.size)}, pc, 2);
    orc_const_div(ha.data(), want.data(), 3.7f, n);
    ExpectBits(Get(c), want, "ArithmeticConstOp x / v");
    // relu forward / backward with aliased inputs (top == bottom), as owl calls it
    auto* rf = new ReluForwardOp();
    PhysicalOp prf = Wrap(rf);
    dev.DoExecute({DataShard(a.dev, a.size)}, {DataShard(c.dev, c.size)}, prf, 3);
    std::vector<float> y(n);
    orc_relu_forward(ha.data(), y.data(), n);
    ExpectBits(Get(c), y, "ReluForwardOp");
    Arr d = Make(dev, s5);
    auto* rb = new ReluBackwardOp();
    PhysicalOp prb = Wrap(rb);
    dev.DoExecute({DataShard(b.dev, b.size), DataShard(c.dev, c.size), DataShard(c.dev, c.size)}, {DataShard(d.dev, d.size)}, prb, 0);
    orc_relu_backward(y.data(), y.data(), hb.data(), want.data(), n);
    ExpectBits(Get(d), want, "ReluBackwardOp(diff, top, top)");
    for (Arr* p : {&a, &b, &c, &d}) dev.data_store().FreeData(p->id);
  }

  {  // MatMult + bias (NormArithmetic) + Reduction + MaxIndex: the FC layer of apps/mnist_common.h
    int m = 64, k = 100, nn = 33;
    auto hw = Randn(size_t(m) * k), hx = Randn(size_t(k) * nn), hb = Randn(m);
    Arr w = Make(dev, Scale{m, k}, &hw), x = Make(dev, Scale{k, nn}, &hx), bias = Make(dev, Scale{m, 1}, &hb);
    Arr y = Make(dev, Scale{m, nn}), z = Make(dev, Scale{m, nn}), r = Make(dev, Scale{m, 1}), am = Make(dev, Scale{1, nn});
    PhysicalOp mm = Wrap(new MatMultOp());
    dev.DoExecute({DataShard(w.dev, w.size), DataShard(x.dev, x.size)}, {DataShard(y.dev, y.size)}, mm, 0);
    std::vector<float> wy(size_t(m) * nn), wz(wy.size()), wr(m), wam(nn);
    orc_matmult(hw.data(), hx.data(), wy.data(), m, nn, k);
    ExpectNormRel(Get(y), wy, 5e-3, "MatMultOp (TF32)");
    auto* na = new NormArithmeticOp();
    na->closure = {ArithmeticType::kAdd, Scale{1}};
    PhysicalOp pna = Wrap(na);
    auto hy = Get(y);
    dev.DoExecute({DataShard(y.dev, y.size), DataShard(bias.dev, bias.size)}, {DataShard(z.dev, z.size)}, pna, 1);
    orc_norm_on_row(0, hy.data(), hb.data(), wz.data(), m, nn);
    ExpectBits(Get(z), wz, "NormArithmeticOp + on dim 1");
    auto* red = new ReductionOp();
    red->closure = {ReductionType::kMax, Scale{1}};
    PhysicalOp pred = Wrap(red);
    dev.DoExecute({DataShard(z.dev, z.size)}, {DataShard(r.dev, r.size)}, pred, 2);
    orc_reduction_on_row(1, wz.data(), wr.data(), m, nn);
    ExpectBits(Get(r), wr, "ReductionOp max dim 1");
    auto* mi = new MaxIndexOp();
    mi->closure = {0};
    PhysicalOp pmi = Wrap(mi);
    dev.DoExecute({DataShard(z.dev, z.size)}, {DataShard(am.dev, am.size)}, pmi, 3);
    orc_max_index_on_col(wz.data(), wam.data(), m, nn);
    ExpectBits(Get(am), wam, "MaxIndexOp dim 0");
  }

  {  // Convolution forward / backward-data / backward-filter + max pooling + instance softmax (LeNet-like)
    int N = 3, Ci = 4, Co = 6, H = 9, W = 9, ph = 1, pw = 1, sv = 1, sh = 1, fh = 3, fw = 3;
    int Ho = (H + 2 * ph - fh) / sv + 1, Wo = (W + 2 * pw - fw) / sh + 1;
    auto hx = Randn(size_t(N) * Ci * H * W), hw = Randn(size_t(Co) * Ci * fh * fw), hb = Randn(Co), hdy = Randn(size_t(N) * Co * Ho * Wo);
    Arr x = Make(dev, Scale{W, H, Ci, N}, &hx), w = Make(dev, Scale{fw, fh, Ci, Co}, &hw), b = Make(dev, Scale{Co}, &hb);
    Arr y = Make(dev, Scale{Wo, Ho, Co, N}), dy = Make(dev, Scale{Wo, Ho, Co, N}, &hdy), dx = Make(dev, x.size), dw = Make(dev, w.size);
    auto* cf = new ConvForwardOp();
    cf->closure = {ph, pw, sv, sh};
    PhysicalOp pcf = Wrap(cf);
    dev.DoExecute({DataShard(x.dev, x.size), DataShard(w.dev, w.size), DataShard(b.dev, b.size)}, {DataShard(y.dev, y.size)}, pcf, 0);
    std::vector<float> wy(hdy.size()), wdx(hx.size()), wdw(hw.size());
    orc_conv_forward(hx.data(), hw.data(), hb.data(), wy.data(), N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw);
    ExpectNormRel(Get(y), wy, 5e-3, "ConvForwardOp");
    auto* cbd = new ConvBackwardDataOp();
    cbd->closure = {ph, pw, sv, sh};
    PhysicalOp pcbd = Wrap(cbd);
    dev.DoExecute({DataShard(dy.dev, dy.size), DataShard(w.dev, w.size)}, {DataShard(dx.dev, dx.size)}, pcbd, 1);
    orc_conv_backward_data(hdy.data(), hw.data(), wdx.data(), N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw);
    ExpectNormRel(Get(dx), wdx, 5e-3, "ConvBackwardDataOp");
    auto* cbf = new ConvBackwardFilterOp();
    cbf->closure = {ph, pw, sv, sh};
    PhysicalOp pcbf = Wrap(cbf);
    dev.DoExecute({DataShard(dy.dev, dy.size), DataShard(x.dev, x.size)}, {DataShard(dw.dev, dw.size)}, pcbf, 2);
    orc_conv_backward_filter(hx.data(), hdy.data(), wdw.data(), N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw);
    ExpectNormRel(Get(dw), wdw, 5e-3, "ConvBackwardFilterOp");
    // max pooling 3x3/2 on the conv output (ceil-mode size rule)
    int Hp = 4, Wp = 4;   // (9 - 3 + 1)/2 + 1 = 4
    Arr pl = Make(dev, Scale{Wp, Hp, Co, N});
    auto* pf = new PoolingForwardOp();
    pf->closure = {PoolingInfo::Algorithm::kMax, 3, 3, 2, 2, 0, 0};
    PhysicalOp ppf = Wrap(pf);
    dev.DoExecute({DataShard(y.dev, y.size)}, {DataShard(pl.dev, pl.size)}, ppf, 3);
    auto hy = Get(y);
    std::vector<float> wpl(size_t(N) * Co * Hp * Wp);
    orc_max_pooling_forward(hy.data(), wpl.data(), N, Co, Ho, Wo, 2, 2, 3, 3, 0, 0);
    ExpectBits(Get(pl), wpl, "PoolingForwardOp max");
    // softmax over {10,1,1,8}
    auto hs = Randn(80, 3.f);
    Arr sx = Make(dev, Scale{10, 1, 1, 8}, &hs), sy = Make(dev, Scale{10, 1, 1, 8});
    auto* sf = new SoftmaxForwardOp();
    sf->closure = {SoftmaxAlgorithm::kInstance};
    PhysicalOp psf = Wrap(sf);
    dev.DoExecute({DataShard(sx.dev, sx.size)}, {DataShard(sy.dev, sy.size)}, psf, 0);
    std::vector<float> wsy(80);
    orc_instance_softmax_forward(hs.data(), wsy.data(), 8, 1, 1, 10);
    ExpectNormRel(Get(sy), wsy, 1e-5, "SoftmaxForwardOp instance");
  }

  {  // data generators and the no-CPU-fallback rule
    Arr f = Make(dev, Scale{7, 3});
    auto* fo = new FillOp();
    fo->closure = {0.25f};
    PhysicalOp pfo = Wrap(fo);
    dev.DoExecute({}, {DataShard(f.dev, f.size)}, pfo, 0);
    ExpectBits(Get(f), std::vector<float>(21, 0.25f), "FillOp");
    ++g_checks;
    bool threw = false;
    try {
      Context cpu;
      cpu.impl_type = ImplType::kBasic;
      fo->Execute({}, {DataShard(f.dev, f.size)}, cpu);
    } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { ++g_fail; std::printf("FAIL kBasic must have no implementation\n"); }
    // pooled store: an exact-size block is reused after FreeData
    float* p0 = f.dev;
    dev.data_store().FreeData(f.id);
    Arr g2 = Make(dev, Scale{7, 3});
    ++g_checks;
    if (g2.dev != p0) { ++g_fail; std::printf("FAIL pool did not reuse the block\n"); }
  }

  if (g_fail) { std::printf("FAILED %d of %d checks\n", g_fail, g_checks); return 1; }
  std::printf("PASS %d checks\n", g_checks);
  return 0;
}
