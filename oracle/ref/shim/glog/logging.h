// Stand-in for <glog/logging.h> — TEST INFRASTRUCTURE ONLY (oracle/build_ref.sh).
// glog is absent in this image; the reference's sources use its CHECK* macros for contract
// violations (abort with a message).  This header gives them the same behaviour: print + abort().
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <iostream>
#include <sstream>

namespace google {
inline void InitGoogleLogging(const char*) {}
inline void InstallFailureSignalHandler() {}
}  // namespace google

namespace glog_shim {
struct Fatal {
  std::ostringstream s;
  Fatal(const char* file, int line, const char* what) {
    s << "[" << file << ":" << line << "] Check failed: " << what << " ";
  }
  template <class T>
  Fatal& operator<<(const T& v) {
    s << v;
    return *this;
  }
  [[noreturn]] ~Fatal() {
    std::fprintf(stderr, "%s\n", s.str().c_str());
    std::abort();
  }
};
struct Null {
  template <class T>
  Null& operator<<(const T&) {
    return *this;
  }
};
struct Voidify {
  void operator&(const Fatal&) {}
  void operator&(const Null&) {}
};
template <class T>
T* NotNull(const char* file, int line, const char* what, T* p) {
  if (p == nullptr) Fatal(file, line, what);
  return p;
}
}  // namespace glog_shim

#define CHECK(c) \
  (c) ? (void)0 : glog_shim::Voidify() & glog_shim::Fatal(__FILE__, __LINE__, #c)
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_NOTNULL(p) glog_shim::NotNull(__FILE__, __LINE__, #p " must be non-NULL", (p))
#define DCHECK(c) CHECK(c)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define LOG(severity) glog_shim::Null()
#define VLOG(n) glog_shim::Null()
#define CHECK_NEAR(a, b, eps) CHECK(((a) > (b) ? (a) - (b) : (b) - (a)) <= (eps))
