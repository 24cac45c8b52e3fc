// TEST INFRASTRUCTURE: the two error-reporting symbols of api.cu for the host-emulated build (tests/native/cuda_host_emu.h).
#include <stdarg.h>
#include <stdio.h>

static thread_local char g_err[512];

extern "C" void ym_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* ym_last_error() { return g_err; }
