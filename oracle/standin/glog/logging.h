// TEST INFRASTRUCTURE ONLY -- stand-in for <glog/logging.h>: LOG(...) swallows
// its stream, CHECK*(...) aborts with a message on failure like glog does.
#ifndef RSM_STANDIN_GLOG_LOGGING_H
#define RSM_STANDIN_GLOG_LOGGING_H

#include <cstdlib>
#include <iostream>
#include <sstream>

namespace rsm_standin {
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  std::ostringstream os;
  template <typename T> FatalStream& operator<<(const T& v) { os << v; return *this; }
  FatalStream& operator<<(std::ostream& (*f)(std::ostream&)) { os << f; return *this; }
  ~FatalStream() { std::cerr << "CHECK failed: " << os.str() << std::endl; std::abort(); }
};
struct Voidify { void operator&(NullStream&) {} void operator&(FatalStream&) {} };
}  // namespace rsm_standin

#define LOG(severity) ::rsm_standin::NullStream()
#define DLOG(severity) ::rsm_standin::NullStream()
#define VLOG(level) ::rsm_standin::NullStream()
#define LOG_IF(severity, cond) ::rsm_standin::NullStream()
#define CHECK(cond) (cond) ? (void)0 : ::rsm_standin::Voidify() & ::rsm_standin::FatalStream() << #cond << " "
#define CHECK_OP_(a, b, op) CHECK((a) op (b))
#define CHECK_EQ(a, b) CHECK_OP_(a, b, ==)
#define CHECK_NE(a, b) CHECK_OP_(a, b, !=)
#define CHECK_LT(a, b) CHECK_OP_(a, b, <)
#define CHECK_LE(a, b) CHECK_OP_(a, b, <=)
#define CHECK_GT(a, b) CHECK_OP_(a, b, >)
#define CHECK_GE(a, b) CHECK_OP_(a, b, >=)

namespace google {
inline void InitGoogleLogging(const char*) {}
}

#endif
