// Scale -- shape vector, FASTEST dimension first (column-major): index (i0,i1,..) of {d0,d1,..} lives at
// i0 + d0*(i1 + d1*(...)).  Interface subset of the reference's minerva/common/scale.h:14-128 that the
// op layer uses (construction, [], NumDims, Prod, ==, iteration, Contains); written for this repo.
#pragma once
#include <cstddef>
#include <initializer_list>
#include <ostream>
#include <utility>
#include <vector>

namespace minerva {

class Scale {
 public:
  Scale() = default;
  Scale(std::initializer_list<int> dims) : v_(dims) {}
  explicit Scale(std::vector<int> dims) : v_(std::move(dims)) {}
  template <typename It> Scale(It first, It last) : v_(first, last) {}

  static Scale Origin(std::size_t nd) { return Scale(std::vector<int>(nd, 0)); }
  static Scale Constant(std::size_t nd, int val) { return Scale(std::vector<int>(nd, val)); }

  int operator[](std::size_t i) const { return v_[i]; }
  int& operator[](std::size_t i) { return v_[i]; }
  int get(int i) const { return v_[static_cast<std::size_t>(i)]; }
  std::size_t NumDims() const { return v_.size(); }
  int Prod() const {   // int, like the reference: one NArray holds < 2^31 elements
    if (v_.empty()) return 0;
    int p = 1;
    for (int d : v_) p *= d;
    return p;
  }
  bool Contains(int a) const {
    for (int d : v_) if (d == a) return true;
    return false;
  }
  bool operator==(const Scale& o) const { return v_ == o.v_; }
  bool operator!=(const Scale& o) const { return v_ != o.v_; }
  std::vector<int>::const_iterator begin() const { return v_.begin(); }
  std::vector<int>::const_iterator end() const { return v_.end(); }
  const std::vector<int>& ToVector() const { return v_; }

 private:
  std::vector<int> v_;
};

inline std::ostream& operator<<(std::ostream& os, const Scale& s) {
  os << "[";
  for (std::size_t i = 0; i < s.NumDims(); ++i) os << (i ? " " : "") << s[i];
  return os << "]";
}

}  // namespace minerva
