// Library-level glue of libirc_sm100.so: error reporting, device gate.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "irc_common.cuh"
#include "../../include/irc_b200.h"

static thread_local char g_err[512] = "";

int irc_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int irc_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return irc_set_error(IRC_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
    return IRC_OK;
}

int irc_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        // IRC_SM_LIMIT: persistent grids use at most this many SMs (leaves room for NCCL's CTAs next to the 1-CTA-per-SM GEMMs when
        // gradient all-reduces overlap the backward pass)
        const char* e = getenv("IRC_SM_LIMIT");
        if (e && atoi(e) > 0 && atoi(e) < sms) sms = atoi(e);
    }
    return sms;
}

// IRC_PDL: bit 0 = memory-bound kernels, bit 1 = GEMMs launched with the programmatic-dependent-launch attribute.
// Default 0 (plain stream serialisation): measured on B200 inside the captured step, 15.10 ms/step off vs 15.26 / 15.14 /
// 15.65 ms with the attribute on the memory-bound kernels / the GEMMs / both (early-parked CTAs cost more than the launch
// gaps they hide).
int irc_pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("IRC_PDL"); v = e ? atoi(e) : 0; }
    return v;
}

extern "C" int irc_version(void) { return 100; }

extern "C" const char* irc_last_error(void) { return g_err; }

extern "C" int irc_arch_check(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return irc_set_error(IRC_ERR_ARCH, "no CUDA device");
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) return irc_set_error(IRC_ERR_ARCH, "libirc_sm100 needs an sm_100 device (found sm_%d%d); there is no fallback path", major, minor);
    return IRC_OK;
}
