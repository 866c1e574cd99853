// Library plumbing of libb200em: error string, launch counter, device info, zero-fill.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace b200em {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    // per-device cache (predict_with_halo drives several devices from several threads)
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev].store(v, std::memory_order_relaxed);
    return v;
}

// Caller-provided, all-zero scratch of the current device (b200em_set_workspace): the first WS_COUNTER_BYTES are tile counters, the
// rest fp32 partial sums.  Kernels that use it leave it all-zero again (the last CTA of a tile clears what it read).
namespace {
struct Workspace { void* p; int64_t bytes; };
Workspace g_ws[64];
}
bool get_workspace(float** acc, int64_t* acc_floats, unsigned** counters, int* ncounters) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    const Workspace w = g_ws[dev];
    if (!w.p || w.bytes <= WS_COUNTER_BYTES) return false;
    *counters = reinterpret_cast<unsigned*>(w.p);
    *ncounters = WS_COUNTER_BYTES / 4;
    *acc = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(w.p) + WS_COUNTER_BYTES);
    *acc_floats = (w.bytes - WS_COUNTER_BYTES) / 4;
    return true;
}

__global__ void memset_zero_kernel(uint4* __restrict__ p, int64_t n16, unsigned char* __restrict__ tail, int ntail) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0, 0, 0, 0);
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

}  // namespace b200em

using namespace b200em;

extern "C" {

int b200em_abi_version(void) { return B200EM_ABI_VERSION; }

const char* b200em_last_error(void) { return g_err; }

int b200em_device_info(int* sm_count_out, int* cc_major, int* cc_minor, int* umma_ok) {
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    int sms = 0, maj = 0, min = 0;
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B2_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    B2_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count_out) *sm_count_out = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    if (umma_ok) *umma_ok = (maj == 10) ? 1 : 0;
    return 0;
}

int b200em_set_workspace(void* workspace, int64_t bytes) {
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CHECK_ARG(dev >= 0 && dev < 64, "set_workspace: device index out of range");
    B2_CHECK_ARG((workspace == nullptr && bytes == 0) || (workspace && bytes > WS_COUNTER_BYTES && aligned16(workspace)),
                 "set_workspace: needs a 16-byte aligned buffer larger than %d bytes (or NULL, 0 to unregister)", WS_COUNTER_BYTES);
    g_ws[dev].p = workspace;
    g_ws[dev].bytes = bytes;
    return 0;
}

int64_t b200em_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void b200em_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

int b200em_memset_zero(void* p, int64_t bytes, void* stream) {
    B2_CHECK_ARG(p && bytes >= 0, "memset_zero: bad arguments");
    if (bytes == 0) return 0;
    B2_CHECK_ARG(aligned16(p), "memset_zero: pointer must be 16-byte aligned");
    int64_t n16 = bytes / 16;
    int ntail = (int)(bytes % 16);
    int64_t blocks = (n16 + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    memset_zero_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((uint4*)p, n16, (unsigned char*)p + n16 * 16, ntail);
    B2_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
