// Error reporting + device queries shared by every entry point of libssd_b200.so.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

namespace ssd {

static thread_local char g_err[512] = "";

char* last_error_buffer() { return g_err; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return (int)e;
}

static int g_pdl_mode = -1;         // ssd_set_pdl: -1 the default policy below, 0 off, 1 on

bool pdl_enabled() {
    if (g_pdl_mode >= 0) return g_pdl_mode == 1;
    // Default ON (measured -4.6 % on the MobileNetV2 step, -6 % with the channel-grouped block kernels whose weight
    // prologue then runs under the predecessor's tail).  SSD_B200_PDL=0 switches it off, =1 forces it on; without the
    // variable it is switched off when a CUDA injection library is attached (Nsight Compute, compute-sanitizer): their
    // kernel replay serialises launches, where a programmatic chain gains nothing and `ncu --set full` has hung on it.
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SSD_B200_PDL");
        if (e && (e[0] == '0' || e[0] == '1')) v = e[0] == '1' ? 1 : 0;
        else v = (getenv("NV_NSIGHT_INJECTION_PORT_BASE") || getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_SANITIZER_INJECTION")) ? 0 : 1;
    }
    return v == 1;
}

static unsigned long long* g_debug_trace = nullptr;
unsigned long long* debug_trace_buffer() { return g_debug_trace; }
void set_debug_trace_buffer(unsigned long long* p) { g_debug_trace = p; }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace ssd

namespace ssd { void set_debug_trace_buffer(unsigned long long* p); }
extern "C" int ssd_debug_trace(void* d_buf) {
    ssd::set_debug_trace_buffer(static_cast<unsigned long long*>(d_buf));
    return SSD_OK;
}

extern "C" int ssd_abi_version(void) { return SSD_B200_ABI_VERSION; }

extern "C" const char* ssd_last_error(void) { return ssd::last_error_buffer(); }

extern "C" int ssd_device_info(char* h_name, int name_len, int* h_sm_count, int* h_cc) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return ssd::cuda_fail(e, "ssd_device_info: cudaGetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return ssd::cuda_fail(e, "ssd_device_info: cudaGetDeviceProperties");
    if (h_name && name_len > 0) {
        strncpy(h_name, prop.name, (size_t)name_len - 1);
        h_name[name_len - 1] = '\0';
    }
    if (h_sm_count) *h_sm_count = prop.multiProcessorCount;
    if (h_cc) *h_cc = prop.major * 10 + prop.minor;
    return SSD_OK;
}

extern "C" int ssd_set_pdl(int mode) {
    ssd::g_pdl_mode = mode < 0 ? -1 : (mode ? 1 : 0);
    return SSD_OK;
}
