/* gflags shim -- the reference uses three DEFINE_bool flags and ParseCommandLineFlags (minerva/system/minerva_system.cpp:14-15,74,
 * minerva/device/device.cpp:26); the vendored gflags 2.1.1 does not configure under this image's cmake.
 * TEST INFRASTRUCTURE: lets oracle/Makefile build the reference's CPU stack from where it lies under /root/reference. */
#ifndef MNV_GFLAGS_SHIM_H_
#define MNV_GFLAGS_SHIM_H_
#include <string>
#define DEFINE_bool(name, val, txt) bool FLAGS_##name = (val)
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DEFINE_int32(name, val, txt) int FLAGS_##name = (val)
#define DEFINE_double(name, val, txt) double FLAGS_##name = (val)
#define DEFINE_string(name, val, txt) std::string FLAGS_##name = (val)
namespace gflags {
inline unsigned ParseCommandLineFlags(int*, char***, bool) { return 1; }
inline void SetUsageMessage(const std::string&) {}
}  // namespace gflags
namespace google = gflags;
#endif
