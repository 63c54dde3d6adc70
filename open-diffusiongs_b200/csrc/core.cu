// core.cu -- error reporting and version of libdgs_b200.so
#include <cstring>

#include "dgs_internal.h"

namespace dgs {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
}  // namespace dgs

extern "C" {
int dgs_version(void) { return DGS_VERSION; }
const char* dgs_last_error(void) { return dgs::get_error(); }
}
