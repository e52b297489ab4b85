// Minimal stand-in for <dmlc/logging.h> so the reference's CPU sources compile out of tree.
// dmlc-core is fetched unpinned from git by the reference build (third_party/CMakeLists.txt:29-46)
// and is not vendored; only these macros are used on the op path.  Test infrastructure only.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

namespace dmlc_shim {
struct Fatal {
  std::ostringstream os;
  Fatal(const char* file, int line) { os << file << ":" << line << ": "; }
  [[noreturn]] ~Fatal() noexcept(false) { throw std::runtime_error(os.str()); }
};
struct Sink {
  template <typename T> Sink& operator<<(const T&) { return *this; }
};
struct Voidify { void operator&(std::ostream&) {} };
}  // namespace dmlc_shim

#define CHECK(x) if (!(x)) ::dmlc_shim::Fatal(__FILE__, __LINE__).os << "Check failed: " #x << ' '
#define CHECK_BINARY_(a, b, op) if (!((a) op (b))) ::dmlc_shim::Fatal(__FILE__, __LINE__).os << "Check failed: " #a " " #op " " #b << ' '
#define CHECK_EQ(a, b) CHECK_BINARY_(a, b, ==)
#define CHECK_NE(a, b) CHECK_BINARY_(a, b, !=)
#define CHECK_LT(a, b) CHECK_BINARY_(a, b, <)
#define CHECK_LE(a, b) CHECK_BINARY_(a, b, <=)
#define CHECK_GT(a, b) CHECK_BINARY_(a, b, >)
#define CHECK_GE(a, b) CHECK_BINARY_(a, b, >=)
#define CHECK_NOTNULL(x) (x)
#define LOG_INFO ::dmlc_shim::Sink()
#define LOG_WARNING ::dmlc_shim::Sink()
#define LOG_ERROR ::dmlc_shim::Sink()
#define LOG_FATAL ::dmlc_shim::Fatal(__FILE__, __LINE__).os
#define LOG(sev) LOG_##sev
#define DLOG(sev) ::dmlc_shim::Sink()
#define DCHECK(x) CHECK(x)
