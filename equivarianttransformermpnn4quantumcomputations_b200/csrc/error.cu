// Thread-local error string of the C ABI (include/eqv2_b200.h).
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void eqv2_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* eqv2_last_error(void) { return g_err; }
extern "C" int eqv2_abi_version(void) { return 1; }
