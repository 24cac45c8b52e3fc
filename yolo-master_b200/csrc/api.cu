// C-ABI plumbing: thread-local error string, version, device probe.
#include <stdarg.h>
#include <stdlib.h>

#include "ym_common.cuh"

static thread_local char g_err[512] = "";

extern "C" void ym_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* ym_last_error(void) { return g_err; }

extern "C" int ym_version(void) { return 100; }

// Programmatic dependent launch of the forward-path kernels (ym_common.cuh: pdl_prologue / launch_pdl).  Default on; the
// environment variable YM_PDL=0 or ym_set_pdl(0) turns the launch attribute off (A/B measurements; results are identical).
static int g_pdl = -1;
extern "C" int ym_pdl_enabled(void) {
    if (g_pdl < 0) {
        const char* e = getenv("YM_PDL");
        g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return g_pdl;
}
extern "C" int ym_set_pdl(int on) {
    const int old = ym_pdl_enabled();
    g_pdl = on ? 1 : 0;
    return old;
}

// Returns 0 and fills sm_major/sm_minor/sm_count/l2_bytes for the current device.
extern "C" int ym_device_info(int* sm_major, int* sm_minor, int* sm_count, long long* l2_bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { ym_set_error("ym_device_info: %s", cudaGetErrorString(e)); return YM_ERR_CUDA; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { ym_set_error("ym_device_info: %s", cudaGetErrorString(e)); return YM_ERR_CUDA; }
    if (sm_major) *sm_major = prop.major;
    if (sm_minor) *sm_minor = prop.minor;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = (long long)prop.l2CacheSize;
    return YM_OK;
}

// Launch priority of the forward-path kernels relative to the long exp-bound attention kernels (which always launch at priority 0, the
// lowest): with several graph instances in flight, the CTA slots an attention kernel frees then go first to the short kernels of the
// other instances instead of to its own remaining CTAs.  0 = off (every kernel at the default priority); a negative value is passed to
// cudaLaunchAttributePriority (numerically lower = scheduled first).  Returns the previous value.
static int g_kernel_priority = 0;
extern "C" int ym_kernel_priority(void) { return g_kernel_priority; }
extern "C" int ym_set_kernel_priority(int prio) {
    const int old = g_kernel_priority;
    if (prio <= 8 && prio >= -8) g_kernel_priority = prio;   // > 0: the ATTENTION kernels launch at -prio and everything else at the default
    return old;
}
