// Convolution / pooling descriptors and algorithm enums passed through the closures
// (reference: minerva/narray/convolution_info.h:5-57).
#pragma once

namespace minerva {

enum class SoftmaxAlgorithm { kInstance, kChannel };
enum class ActivationAlgorithm { kSigmoid, kRelu, kTanh };

struct ConvInfo {
  ConvInfo(int ph = 0, int pw = 0, int sv = 1, int sh = 1)
      : pad_height(ph), pad_width(pw), stride_vertical(sv), stride_horizontal(sh) {}
  int pad_height, pad_width, stride_vertical, stride_horizontal;
};

struct PoolingInfo {
  enum class Algorithm { kMax, kAverage };
  PoolingInfo(Algorithm alg = Algorithm::kMax, int h = 0, int w = 0, int sv = 1, int sh = 1, int ph = 0, int pw = 0)
      : algorithm(alg), height(h), width(w), stride_vertical(sv), stride_horizontal(sh), pad_height(ph), pad_width(pw) {}
  Algorithm algorithm;
  int height, width, stride_vertical, stride_horizontal, pad_height, pad_width;
};

}  // namespace minerva
