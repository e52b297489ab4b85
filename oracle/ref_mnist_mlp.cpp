// ref_mnist_mlp.cpp -- BASELINE.json configs[0] ("apps/mnist_mlp 784-256-10 MLP, batch 256, on the CPU (basic) device") run
// THROUGH THE REFERENCE'S OWN STACK: NArray -> DagScheduler -> CpuDevice (4 worker threads) -> basic:: ops, all compiled from
// the sources where they lie under /root/reference by oracle/Makefile (target `refstack`).  TEST INFRASTRUCTURE / CPU baseline
// only; no reference source is copied here -- this file is a client of the reference's public API (minerva.h), with the op
// sequence of the reference's apps/mnist_common.h:224-288 (MnistMlpAlgo) on synthetic MNIST-shaped data.
//
// One deviation, forced by the reference itself: Elewise::ReluBackward has NO CPU implementation
// (op/impl/bundle.h:32 -> "no implementation for ReluBackwardClosure"), so the stock app cannot finish one training step on the
// CPU device.  The backward ReLU is supplied as a user ComputeFn through NArray::ComputeOne -- the plug-in surface the
// reference's own tests use for custom ops (tests/unittest_perf.cpp:30-51) -- with the formula the reference's CUDA path
// implements (dx = x > 0 ? dy : 0).
//
//   ref_mnist_mlp [mb=256] [steps=20] [warmup=3]   ->  one JSON line: images/s of the reference CPU stack
#include <minerva.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>

using namespace minerva;

class ReluBackwardFn : public ComputeFn {
 public:
  void Execute(const DataList& inputs, const DataList& outputs, const Context&) {   // inputs (diff, top, bottom)
    const float* dy = inputs[0].data_;
    const float* x = inputs[2].data_;
    float* dx = outputs[0].data_;
    const int n = outputs[0].size_.Prod();
    for (int i = 0; i < n; ++i) dx[i] = x[i] > 0 ? dy[i] : 0;
  }
  std::string Name() const { return "relu backward (user fn)"; }
};

static std::shared_ptr<float> Buf(size_t n) { return std::shared_ptr<float>(new float[n], [](float* p) { delete[] p; }); }

int main(int argc, char** argv) {
  const int mb = argc > 1 ? std::atoi(argv[1]) : 256, steps = argc > 2 ? std::atoi(argv[2]) : 20, warmup = argc > 3 ? std::atoi(argv[3]) : 3;
  const float alpha = 0.01f;
  int ac = 1;
  char** av = argv;
  MinervaSystem::Initialize(&ac, &av);
  MinervaSystem& ms = MinervaSystem::Instance();
  ms.SetDevice(ms.CreateCpuDevice());
  // apps/mnist_common.h:233-236
  NArray w0 = NArray::Randn({256, 784}, 0.0, 0.1), b0 = NArray::Randn({256, 1}, 0.0, 0.1);
  NArray w1 = NArray::Randn({10, 256}, 0.0, 0.1), b1 = NArray::Randn({10, 1}, 0.0, 0.1);
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> u(0.f, 1.f);
  double loss_proxy = 0;
  auto step = [&]() {
    auto data = Buf(static_cast<size_t>(784) * mb), label = Buf(static_cast<size_t>(10) * mb);
    for (size_t i = 0; i < static_cast<size_t>(784) * mb; ++i) data.get()[i] = u(rng);
    for (size_t i = 0; i < static_cast<size_t>(10) * mb; ++i) label.get()[i] = 0.f;
    for (int i = 0; i < mb; ++i) label.get()[static_cast<size_t>(i) * 10 + rng() % 10] = 1.f;
    // FF (mnist_common.h:238-247)
    NArray a0 = NArray::MakeNArray({784, mb}, data);
    NArray a1 = Elewise::ReluForward(w0 * a0 + b0);
    NArray a2 = w1 * a1 + b1;
    NArray a3 = Convolution::SoftmaxForward(a2.Reshape({10, 1, 1, mb}), SoftmaxAlgorithm::kInstance).Reshape({10, mb});
    // BP (mnist_common.h:248-264)
    NArray lab = NArray::MakeNArray({10, mb}, label);
    NArray s2 = a3 - lab;
    NArray pre = w1.Trans() * s2;
    NArray s1 = NArray::ComputeOne({pre, a1, a1}, a1.Size(), new ReluBackwardFn());
    NArray gw0 = s1 * a0.Trans(), gb0 = s1.Sum(1), gw1 = s2 * a1.Trans(), gb1 = s2.Sum(1);
    // Update (mnist_common.h:265-275)
    w0 -= alpha / mb * gw0; b0 -= alpha / mb * gb0; w1 -= alpha / mb * gw1; b1 -= alpha / mb * gb1;
    return a3;
  };
  for (int i = 0; i < warmup; ++i) step();
  ms.WaitForAll();
  const auto t0 = std::chrono::steady_clock::now();
  NArray last;
  for (int i = 0; i < steps; ++i) last = step();
  ms.WaitForAll();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::shared_ptr<float> p = last.Get();
  for (int i = 0; i < 10; ++i) loss_proxy += p.get()[i];       // column 0 of the softmax sums to 1
  std::printf("{\"app\": \"reference mnist_mlp (NArray -> DagScheduler -> CpuDevice -> basic::)\", \"mb\": %d, \"steps\": %d, \"warmup\": %d, "
              "\"ms_per_step\": %.4f, \"images_per_s\": %.2f, \"softmax_column_sum\": %.6f, \"cpu_worker_threads\": 4}\n",
              mb, steps, warmup, 1e3 * sec / steps, mb * steps / sec, loss_proxy);
  return 0;
}
