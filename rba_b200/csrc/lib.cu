// Library-level state: error string, launch counter, version.
#include "common.cuh"

namespace rba {
std::atomic<int64_t> g_launches{0};
std::string& last_error_ref() {
  thread_local std::string e;
  return e;
}
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}
}  // namespace rba

extern "C" const char* rba_last_error(void) { return rba::last_error_ref().c_str(); }
extern "C" int rba_version(void) { return 1; }
extern "C" int64_t rba_launch_count(void) { return rba::g_launches.load(); }
