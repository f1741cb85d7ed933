// capi.cu -- library-level entry points and error reporting for libmvedit_b200.
#include "common.cuh"
#include "../../include/mvedit_b200.h"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void mve_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {
int mve_version(void) { return 1; }
const char* mve_last_error(void) { return g_err; }
}
