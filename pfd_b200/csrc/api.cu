// Library-level entry points of the C ABI: version, thread-local error text, launch counter.
#include <stdarg.h>

#include "../../include/pfd_b200.h"
#include "common.h"

namespace pfd {
thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};

int set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return 1;
}
}  // namespace pfd

extern "C" PFD_API int pfd_version(void) { return PFD_ABI_VERSION; }
extern "C" PFD_API const char* pfd_last_error(void) { return pfd::g_last_error.c_str(); }
extern "C" PFD_API int64_t pfd_launch_count(void) { return pfd::g_launches.load(); }
