// mnist_apps -- BASELINE.json configs[0..1] (apps/mnist_mlp, apps/mnist_cnn) as C++ programs over the plug-in surface:
// every op is a ComputeFn built from its closure, wrapped in a Task and pushed to the StreamDevice, which calls
// ComputeFn::Execute(inputs, outputs, Context{kCuda, stream, workspace}) -- the path the reference's DAG scheduler drives
// (minerva/device/device.cpp:68-119).  Op sequences, shapes, fillers and the update rule follow the reference apps
// (apps/mnist_common.h:123-222 MnistCnnAlgo, :224-288 MnistMlpAlgo); data is synthetic (MNIST-shaped).
//
//   mnist_apps --net lenet|mlp [--mb 256] [--steps 50] [--warmup 5] [--completion enqueue|event|blocking]
//              [--alpha 0.01] [--seed 1] [--dump-dir DIR]
// prints one JSON line: images/s (device-timed with CUDA events around the timed steps), losses, pool statistics.
// --dump-dir writes the initial parameters, the batch and the final parameters as raw fp32 files so that the Python
// owl.net twin of the same net can be checked against this program bit for bit (tests/test_gpu_d_cpp_plugin.py).
#include <cuda_runtime.h>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>
#include "device/stream_device.h"

using namespace minerva;

namespace {

struct Arr {            // a handle, like an NArray: data id + shape
  uint64_t id = 0;
  Scale size;
  bool valid() const { return id != 0; }
};

class Listener : public DeviceListener {
 public:
  void OnOperationComplete(Task* t) override { ++done; delete t; }
  std::atomic<uint64_t> done{0};
};

// Program-order graph builder: what NArray::Compute + the scheduler do, minus laziness (out of scope).
class G {
 public:
  G(StreamDevice& d) : dev(d) {}
  template <class Op, class Closure>
  Arr Run(const Closure& c, std::initializer_list<Arr> in, const Scale& out_size) {
    auto op = std::make_shared<Op>();
    op->closure = c;
    return Push(op, in, out_size);
  }
  Arr Push(std::shared_ptr<ComputeFn> fn, std::initializer_list<Arr> in, const Scale& out_size) {
    Task* t = new Task();
    for (const Arr& a : in) t->inputs.emplace_back(PhysicalData(a.size, dev.device_id(), a.id), a.id);
    Arr out;
    out.id = next_id++;
    out.size = out_size;
    t->outputs.emplace_back(PhysicalData(out_size, dev.device_id(), out.id), out.id);
    t->op.compute_fn = fn;
    t->op.device_id = dev.device_id();
    t->id = out.id;
    ++ops;
    dev.PushTask(t);
    return out;
  }
  void Free(Arr& a) { if (a.valid()) dev.FreeDataIfExist(a.id); a.id = 0; }
  // ---- the NArray operators the apps use ------------------------------------------------------------------------
  Arr Load(const std::vector<float>& host, const Scale& s) {
    std::shared_ptr<float> p(new float[host.size()], [](float* q) { delete[] q; });
    std::memcpy(p.get(), host.data(), host.size() * sizeof(float));
    return Run<ArrayLoaderOp>(ArrayLoaderClosure{p}, {}, s);
  }
  Arr Zeros(const Scale& s) { return Run<FillOp>(FillClosure{0.f}, {}, s); }
  Arr Arith(ArithmeticType t, Arr a, Arr b) { return Run<ArithmeticOp>(ArithmeticClosure{t}, {a, b}, a.size); }
  Arr MulConst(Arr a, float v) { return Run<ArithmeticConstOp>(ArithmeticConstClosure{ArithmeticType::kMult, v, 0}, {a}, a.size); }
  Arr MatMult(Arr a, Arr b) { return Run<MatMultOp>(MatMultClosure{}, {a, b}, Scale{a.size[0], b.size[1]}); }
  Arr Trans(Arr a) { return Run<TransOp>(TransposeClosure{}, {a}, Scale{a.size[1], a.size[0]}); }
  Arr Reshape(Arr a, const Scale& s) { return Run<ReshapeOp>(ReshapeClosure{}, {a}, s); }
  Arr AddBias(Arr m, Arr b) { return Run<NormArithmeticOp>(NormArithmeticClosure{ArithmeticType::kAdd, Scale{1}}, {m, b}, m.size); }
  Arr SumDim1(Arr m) { return Run<ReductionOp>(ReductionClosure{ReductionType::kSum, Scale{1}}, {m}, Scale{m.size[0], 1}); }
  Arr SumDim0(Arr m) { return Run<ReductionOp>(ReductionClosure{ReductionType::kSum, Scale{0}}, {m}, Scale{1, m.size[1]}); }
  Arr Ln(Arr a) { return Run<ElewiseOp>(ElewiseClosure{ElewiseType::kLn}, {a}, a.size); }
  Arr Relu(Arr a) { return Run<ReluForwardOp>(ReluForwardClosure{}, {a}, a.size); }
  Arr ReluBack(Arr diff, Arr top, Arr bottom) { return Run<ReluBackwardOp>(ReluBackwardClosure{}, {diff, top, bottom}, diff.size); }
  Arr ActFwd(Arr a) { return Run<ActivationForwardOp>(ActivationForwardClosure{ActivationAlgorithm::kRelu}, {a}, a.size); }
  Arr ActBwd(Arr diff, Arr top, Arr bottom) {
    return Run<ActivationBackwardOp>(ActivationBackwardClosure{ActivationAlgorithm::kRelu}, {diff, top, bottom}, diff.size);
  }
  Arr Softmax(Arr a) { return Run<SoftmaxForwardOp>(SoftmaxForwardClosure{SoftmaxAlgorithm::kInstance}, {a}, a.size); }
  Arr Conv(Arr x, Arr w, Arr b, const ConvInfo& ci) {
    const int wo = (x.size[0] + 2 * ci.pad_width - w.size[0]) / ci.stride_horizontal + 1;
    const int ho = (x.size[1] + 2 * ci.pad_height - w.size[1]) / ci.stride_vertical + 1;
    return Run<ConvForwardOp>(ConvForwardClosure{ci.pad_height, ci.pad_width, ci.stride_vertical, ci.stride_horizontal}, {x, w, b},
                              Scale{wo, ho, w.size[3], x.size[3]});
  }
  Arr ConvBwdData(Arr diff, Arr bottom, Arr w, const ConvInfo& ci) {
    return Run<ConvBackwardDataOp>(ConvBackwardDataClosure{ci.pad_height, ci.pad_width, ci.stride_vertical, ci.stride_horizontal},
                                   {diff, w}, bottom.size);
  }
  Arr ConvBwdFilter(Arr diff, Arr bottom, Arr w, const ConvInfo& ci) {
    return Run<ConvBackwardFilterOp>(ConvBackwardFilterClosure{ci.pad_height, ci.pad_width, ci.stride_vertical, ci.stride_horizontal},
                                     {diff, bottom}, w.size);
  }
  Arr ConvBwdBias(Arr diff) { return Run<ConvBackwardBiasOp>(ConvBackwardBiasClosure{}, {diff}, Scale{diff.size[2]}); }
  static int Pooled(int x, int pad, int win, int stride) {   // narray/convolution.cpp:107-114
    int p = (x + 2 * pad - win + stride - 1) / stride + 1;
    if ((p - 1) * stride >= x + pad) --p;
    return p;
  }
  Arr Pool(Arr x, const PoolingInfo& pi) {
    return Run<PoolingForwardOp>(PoolingForwardClosure{pi.algorithm, pi.height, pi.width, pi.stride_vertical, pi.stride_horizontal,
                                                      pi.pad_height, pi.pad_width}, {x},
                                 Scale{Pooled(x.size[0], pi.pad_width, pi.width, pi.stride_horizontal),
                                       Pooled(x.size[1], pi.pad_height, pi.height, pi.stride_vertical), x.size[2], x.size[3]});
  }
  Arr PoolBwd(Arr diff, Arr top, Arr bottom, const PoolingInfo& pi) {
    return Run<PoolingBackwardOp>(PoolingBackwardClosure{pi.algorithm, pi.height, pi.width, pi.stride_vertical, pi.stride_horizontal,
                                                        pi.pad_height, pi.pad_width}, {diff, top, bottom}, bottom.size);
  }
  // w -= (alpha / mb) * grad   (mnist_common.h:193-204): ArithmeticConst mult, then Arithmetic sub; the old w is released
  void Update(Arr& w, Arr& grad, float scale) {
    Arr s = MulConst(grad, scale);
    Arr nw = Arith(ArithmeticType::kSub, w, s);
    Free(s); Free(w); Free(grad);
    w = nw;
  }
  std::vector<float> Get(Arr a) {
    std::vector<float> h(static_cast<size_t>(a.size.Prod()));
    dev.CopyToHost(a.id, h.data(), h.size());
    return h;
  }
  StreamDevice& dev;
  uint64_t next_id = 1;
  uint64_t ops = 0;
};

std::vector<float> Gaussian(size_t n, float sd, std::mt19937& rng) {
  std::normal_distribution<float> d(0.f, sd);
  std::vector<float> v(n);
  for (float& x : v) x = d(rng);
  return v;
}

void Dump(const std::string& dir, const std::string& name, const std::vector<float>& v) {
  if (dir.empty()) return;
  std::ofstream f(dir + "/" + name + ".dat", std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(float)));
}

struct Net {
  virtual ~Net() {}
  virtual void Init(G& g, std::mt19937& rng, const std::string& dump) = 0;
  virtual Arr Step(G& g, Arr data, Arr label, int mb, float alpha) = 0;   // returns the softmax output (caller frees)
  virtual void DumpParams(G& g, const std::string& dump, const std::string& tag) = 0;
  virtual Scale DataSize(int mb) const = 0;
};

struct Lenet : Net {
  Arr w[3], b[3];
  ConvInfo ci[2];
  PoolingInfo pi[2];
  void Init(G& g, std::mt19937& rng, const std::string& dump) override {
    ci[0] = ConvInfo(0, 0, 1, 1); ci[1] = ConvInfo(2, 2, 1, 1);
    pi[0] = PoolingInfo(PoolingInfo::Algorithm::kMax, 2, 2, 2, 2); pi[1] = PoolingInfo(PoolingInfo::Algorithm::kMax, 3, 3, 3, 3);
    const Scale ws[3] = {Scale{5, 5, 1, 16}, Scale{5, 5, 16, 32}, Scale{10, 512}}, bs[3] = {Scale{16}, Scale{32}, Scale{10, 1}};
    for (int i = 0; i < 3; ++i) {
      std::vector<float> hw = Gaussian(ws[i].Prod(), 0.1f, rng), hb = Gaussian(bs[i].Prod(), 0.1f, rng);
      Dump(dump, "w" + std::to_string(i) + "_init", hw); Dump(dump, "b" + std::to_string(i) + "_init", hb);
      w[i] = g.Load(hw, ws[i]); b[i] = g.Load(hb, bs[i]);
    }
  }
  Scale DataSize(int mb) const override { return Scale{28, 28, 1, mb}; }
  Arr Step(G& g, Arr data, Arr label, int mb, float alpha) override {
    Arr gw[3], gb[3];
    const Scale ws[3] = {w[0].size, w[1].size, w[2].size}, bs[3] = {b[0].size, b[1].size, b[2].size};
    for (int i = 0; i < 3; ++i) { gw[i] = g.Zeros(ws[i]); gb[i] = g.Zeros(bs[i]); }     // ResetGrad
    // FF (mnist_common.h:153-168)
    Arr a1 = g.Conv(data, w[0], b[0], ci[0]);
    Arr a2 = g.ActFwd(a1);
    Arr a3 = g.Pool(a2, pi[0]);
    Arr a4 = g.Conv(a3, w[1], b[1], ci[1]);
    Arr a5 = g.ActFwd(a4);
    Arr a6 = g.Pool(a5, pi[1]);
    Arr r6 = g.Reshape(a6, Scale{a6.size.Prod() / mb, mb});
    Arr m7 = g.MatMult(w[2], r6);
    Arr a7 = g.AddBias(m7, b[2]);
    Arr r7 = g.Reshape(a7, Scale{10, 1, 1, mb});
    Arr a8 = g.Softmax(r7);
    // BP (mnist_common.h:169-192)
    Arr s8 = g.Arith(ArithmeticType::kSub, a8, label);
    Arr s7 = g.Reshape(s8, Scale{10, mb});
    Arr wt = g.Trans(w[2]);
    Arr m6 = g.MatMult(wt, s7);
    Arr s6 = g.Reshape(m6, a6.size);
    Arr s5 = g.PoolBwd(s6, a6, a5, pi[1]);
    Arr s4 = g.ActBwd(s5, a5, a4);
    Arr s3 = g.ConvBwdData(s4, a3, w[1], ci[1]);
    Arr s2 = g.PoolBwd(s3, a3, a2, pi[0]);
    Arr s1 = g.ActBwd(s2, a2, a1);
    auto acc = [&](Arr& grad, Arr add) { Arr n = g.Arith(ArithmeticType::kAdd, grad, add); g.Free(grad); g.Free(add); grad = n; };
    acc(gw[0], g.ConvBwdFilter(s1, data, w[0], ci[0]));
    acc(gb[0], g.ConvBwdBias(s1));
    acc(gw[1], g.ConvBwdFilter(s4, a3, w[1], ci[1]));
    acc(gb[1], g.ConvBwdBias(s4));
    Arr r6t = g.Trans(r6);
    acc(gw[2], g.MatMult(s7, r6t));
    acc(gb[2], g.SumDim1(s7));
    for (Arr* t : {&a1, &a2, &a3, &a4, &a5, &a6, &r6, &m7, &a7, &r7, &s8, &s7, &wt, &m6, &s6, &s5, &s4, &s3, &s2, &s1, &r6t}) g.Free(*t);
    // Update (mnist_common.h:193-204)
    const float sc = alpha / mb;
    for (int i = 0; i < 3; ++i) { g.Update(w[i], gw[i], sc); g.Update(b[i], gb[i], sc); }
    return a8;
  }
  void DumpParams(G& g, const std::string& dump, const std::string& tag) override {
    for (int i = 0; i < 3; ++i) { Dump(dump, "w" + std::to_string(i) + "_" + tag, g.Get(w[i])); Dump(dump, "b" + std::to_string(i) + "_" + tag, g.Get(b[i])); }
  }
};

struct Mlp : Net {
  Arr w[2], b[2];
  void Init(G& g, std::mt19937& rng, const std::string& dump) override {
    const Scale ws[2] = {Scale{256, 784}, Scale{10, 256}}, bs[2] = {Scale{256, 1}, Scale{10, 1}};
    for (int i = 0; i < 2; ++i) {
      std::vector<float> hw = Gaussian(ws[i].Prod(), 0.1f, rng), hb = Gaussian(bs[i].Prod(), 0.1f, rng);
      Dump(dump, "w" + std::to_string(i) + "_init", hw); Dump(dump, "b" + std::to_string(i) + "_init", hb);
      w[i] = g.Load(hw, ws[i]); b[i] = g.Load(hb, bs[i]);
    }
  }
  Scale DataSize(int mb) const override { return Scale{784, mb}; }
  Arr Step(G& g, Arr data, Arr label, int mb, float alpha) override {
    Arr gw[2], gb[2];
    for (int i = 0; i < 2; ++i) { gw[i] = g.Zeros(w[i].size); gb[i] = g.Zeros(b[i].size); }
    // FF (mnist_common.h:238-247)
    Arr m1 = g.MatMult(w[0], data);
    Arr z1 = g.AddBias(m1, b[0]);
    Arr a1 = g.Relu(z1);
    Arr m2 = g.MatMult(w[1], a1);
    Arr a2 = g.AddBias(m2, b[1]);
    Arr r2 = g.Reshape(a2, Scale{10, 1, 1, mb});
    Arr sm = g.Softmax(r2);
    Arr a3 = g.Reshape(sm, Scale{10, mb});
    // BP (mnist_common.h:248-264); ReluBackward(diff, top, bottom) with top == bottom == acts[1]
    Arr l2 = g.Reshape(label, Scale{10, mb});
    Arr s2 = g.Arith(ArithmeticType::kSub, a3, l2);
    Arr wt = g.Trans(w[1]);
    Arr ms = g.MatMult(wt, s2);
    Arr s1 = g.ReluBack(ms, a1, a1);
    auto acc = [&](Arr& grad, Arr add) { Arr n = g.Arith(ArithmeticType::kAdd, grad, add); g.Free(grad); g.Free(add); grad = n; };
    Arr dt = g.Trans(data);
    acc(gw[0], g.MatMult(s1, dt));
    acc(gb[0], g.SumDim1(s1));
    Arr a1t = g.Trans(a1);
    acc(gw[1], g.MatMult(s2, a1t));
    acc(gb[1], g.SumDim1(s2));
    for (Arr* t : {&m1, &z1, &a1, &m2, &a2, &r2, &a3, &l2, &s2, &wt, &ms, &s1, &dt, &a1t}) g.Free(*t);
    const float sc = alpha / mb;
    for (int i = 0; i < 2; ++i) { g.Update(w[i], gw[i], sc); g.Update(b[i], gb[i], sc); }
    return sm;
  }
  void DumpParams(G& g, const std::string& dump, const std::string& tag) override {
    for (int i = 0; i < 2; ++i) { Dump(dump, "w" + std::to_string(i) + "_" + tag, g.Get(w[i])); Dump(dump, "b" + std::to_string(i) + "_" + tag, g.Get(b[i])); }
  }
};

// -sum(ln(p) o label) / mb through the op surface, read back with a blocking copy (the apps' PrintAccuracy-style read)
float Loss(G& g, Arr prob, Arr label, int mb) {
  Arr p2 = g.Reshape(prob, Scale{10, mb});
  Arr l2 = g.Reshape(label, Scale{10, mb});
  Arr ln = g.Ln(p2);
  Arr pr = g.Arith(ArithmeticType::kMult, ln, l2);
  Arr s0 = g.SumDim0(pr);
  Arr s1 = g.SumDim1(s0);
  const float v = g.Get(s1)[0];
  for (Arr* t : {&p2, &l2, &ln, &pr, &s0, &s1}) g.Free(*t);
  return -v / mb;
}

}  // namespace

int main(int argc, char** argv) {
  std::string net_name = "lenet", completion = "enqueue", dump;
  int mb = 256, steps = 50, warmup = 5, seed = 1, gpu = 0;
  float alpha = 0.01f;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i], v = argv[i + 1];
    if (k == "--net") net_name = v; else if (k == "--mb") mb = std::atoi(v.c_str()); else if (k == "--steps") steps = std::atoi(v.c_str());
    else if (k == "--warmup") warmup = std::atoi(v.c_str()); else if (k == "--completion") completion = v;
    else if (k == "--alpha") alpha = static_cast<float>(std::atof(v.c_str())); else if (k == "--seed") seed = std::atoi(v.c_str());
    else if (k == "--dump-dir") dump = v; else if (k == "--gpu") gpu = std::atoi(v.c_str());
    else { std::fprintf(stderr, "unknown option %s\n", k.c_str()); return 2; }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { std::fprintf(stderr, "mnist_apps: no CUDA device (there is no CPU path)\n"); return 3; }
  const Completion mode = completion == "blocking" ? Completion::kBlocking : completion == "event" ? Completion::kEvent : Completion::kEnqueue;
  try {
    Listener listener;
    StreamDevice dev(1, &listener, gpu, mode);
    G g(dev);
    std::unique_ptr<Net> net;
    if (net_name == "mlp") net.reset(new Mlp()); else net.reset(new Lenet());
    std::mt19937 rng(static_cast<unsigned>(seed));
    net->Init(g, rng, dump);
    // synthetic MNIST-shaped batch: pixels U[0,1), one-hot labels
    const Scale ds = net->DataSize(mb);
    std::vector<float> hx(static_cast<size_t>(ds.Prod())), hl(static_cast<size_t>(10) * mb, 0.f);
    std::uniform_real_distribution<float> u(0.f, 1.f);
    for (float& x : hx) x = u(rng);
    for (int i = 0; i < mb; ++i) hl[static_cast<size_t>(i) * 10 + rng() % 10] = 1.f;
    Dump(dump, "data", hx); Dump(dump, "label", hl);
    Arr data = g.Load(hx, ds), label = g.Load(hl, Scale{10, 1, 1, mb});
    float loss_first = 0.f, loss_last = 0.f;
    for (int it = 0; it < warmup; ++it) {
      Arr p = net->Step(g, data, label, mb, alpha);
      if (it == 0) loss_first = Loss(g, p, label, mb);
      g.Free(p);
    }
    dev.WaitForAll();
    const uint64_t ops0 = g.ops;
    const StreamDevice::Stats st0 = dev.stats();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // the timed steps are enqueued behind e0 on every stream's dependency chain: record on the legacy default stream after
    // a full device sync, and measure to a second full sync -- host-observed time of K steps with the queue kept full
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    const auto t0 = std::chrono::steady_clock::now();
    Arr last;
    for (int it = 0; it < steps; ++it) {
      if (last.valid()) g.Free(last);
      last = net->Step(g, data, label, mb, alpha);
    }
    dev.WaitForAll();
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    float ev_ms = 0.f;
    cudaEventElapsedTime(&ev_ms, e0, e1);
    loss_last = Loss(g, last, label, mb);
    g.Free(last);
    const StreamDevice::Stats st1 = dev.stats();
    net->DumpParams(g, dump, "final");
    dev.WaitForAll();
    std::printf("{\"app\": \"%s\", \"completion\": \"%s\", \"mb\": %d, \"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.6f, "
                "\"images_per_s\": %.3f, \"event_ms_per_step\": %.6f, \"ops_per_step\": %.1f, \"loss_first\": %.7f, \"loss_last\": %.7f, "
                "\"cuda_mallocs_in_timed_region\": %llu, \"pool_hits_in_timed_region\": %llu, \"cross_stream_waits_in_timed_region\": %llu, "
                "\"pool_bytes_reserved\": %llu, \"listener_completions\": %llu}\n",
                net_name.c_str(), completion.c_str(), mb, steps, warmup, wall_ms / steps, mb * steps / (wall_ms * 1e-3), ev_ms / steps,
                static_cast<double>(g.ops - ops0) / steps, loss_first, loss_last,
                static_cast<unsigned long long>(st1.cuda_mallocs - st0.cuda_mallocs), static_cast<unsigned long long>(st1.pool_hits - st0.pool_hits),
                static_cast<unsigned long long>(st1.cross_stream_waits - st0.cross_stream_waits),
                static_cast<unsigned long long>(st1.bytes_reserved), static_cast<unsigned long long>(listener.done.load()));
  } catch (const std::exception& ex) {
    std::fprintf(stderr, "mnist_apps: %s\n", ex.what());
    return 1;
  }
  return 0;
}
