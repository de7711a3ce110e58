// extern "C" boundary of libpose_b200.so (see include/pose_b200.h for the contract): library-wide state.
// The library is four translation units -- api_core.cu (this: error text, launch counter, launch-configuration cache, source
// hash, host-side template), api_sbp.cu, api_spm.cu, api_oks.cu -- compiled separately and linked into one .so (build.py).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <unordered_map>

#include "host_common.h"

namespace pose_host {
namespace {
thread_local char g_err[512] = "";
thread_local int g_err_code = 0;
std::atomic<unsigned long long> g_launches{0};

struct CfgHash {
    size_t operator()(const CfgKey& k) const {
        return std::hash<const void*>()(k.fn) ^ (std::hash<size_t>()(k.smem) * 1000003u) ^ ((size_t)k.dev << 20) ^ (size_t)k.threads;
    }
};
std::mutex g_cfg_mutex;
std::unordered_map<CfgKey, int, CfgHash> g_resident;      // -> resident CTAs on the whole device
std::unordered_map<CfgKey, size_t, CfgHash> g_dyn_smem;   // (kernel, device, 0, 0) -> largest dynamic shared memory opted in to so far
int g_sms[64];                                            // SM count per device ordinal (0: not queried yet)
}  // namespace

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    g_err_code = code;
    return code;
}
int last_code() { return g_err_code ? g_err_code : POSE_EINVAL; }

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return POSE_OK;
}

int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

int sm_count() {
    const int dev = current_device();
    if (dev >= 0 && dev < 64 && g_sms[dev] > 0) return g_sms[dev];
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    if (dev >= 0 && dev < 64) g_sms[dev] = sms;
    return sms;
}

bool cfg_lookup(const CfgKey& key, int* resident) {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    auto it = g_resident.find(key);
    if (it == g_resident.end()) return false;
    *resident = it->second;
    return true;
}
void cfg_store(const CfgKey& key, int resident) {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    g_resident[key] = resident;
}
bool cfg_dyn_smem_needs_raise(const void* fn, int dev, size_t smem) {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    return smem > g_dyn_smem[CfgKey{fn, dev, 0, 0}];
}
void cfg_dyn_smem_set(const void* fn, int dev, size_t smem) {
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    size_t& have = g_dyn_smem[CfgKey{fn, dev, 0, 0}];
    if (smem > have) have = smem;
}
}  // namespace pose_host

using namespace pose_host;

extern "C" {

#ifndef POSE_B200_SOURCE_HASH_STR
#define POSE_B200_SOURCE_HASH_STR "unhashed-build"
#endif
// searched for in the file by build.py / _cabi.py (no dlopen needed): the hash of the sources this binary was compiled from
static const char g_source_hash[] = "POSE_B200_SOURCE_HASH=" POSE_B200_SOURCE_HASH_STR;

int pose_b200_version(void) { return 200; }
const char* pose_b200_source_hash(void) { return g_source_hash + sizeof("POSE_B200_SOURCE_HASH=") - 1; }
const char* pose_b200_last_error(void) { return pose_host::g_err; }
unsigned long long pose_b200_launch_count(void) { return pose_host::g_launches.load(std::memory_order_relaxed); }

int pose_gauss_template_host(double sigma, float* out_host, int capacity) {
    if (!(sigma > 0.0) || !out_host) return fail(POSE_EINVAL, "gauss template: sigma must be > 0");
    const double size = 6.0 * sigma + 3.0;
    const int n = (int)std::ceil(size);          // len(np.arange(0, size, 1.0))
    if (n * n > capacity) return fail(POSE_EINVAL, "gauss template: capacity %d < %d", capacity, n * n);
    const double c = 3.0 * sigma + 1.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const double dx = (double)j - c, dy = (double)i - c;
            out_host[i * n + j] = (float)std::exp(-(dx * dx + dy * dy) / (2.0 * (sigma * sigma)));
        }
    return n;
}


int pose_gauss_template_padded_host(double sigma, float* out_host, int capacity) {
    if (!(sigma > 0.0) || !out_host) return pose_host::fail(POSE_EINVAL, "gauss template: sigma must be > 0");
    const int n = (int)std::ceil(6.0 * sigma + 3.0);
    const int pw = n + 2 * pose::kTemplatePad;
    if ((n + 1) * pw > capacity) return pose_host::fail(POSE_EINVAL, "padded gauss template: capacity %d < %d", capacity, (n + 1) * pw);
    const double c = 3.0 * sigma + 1.0;
    for (int i = 0; i <= n; ++i)
        for (int j = 0; j < pw; ++j) {
            const int jj = j - pose::kTemplatePad;
            float v = 0.0f;
            if (i < n && jj >= 0 && jj < n) {
                const double dx = (double)jj - c, dy = (double)i - c;
                v = (float)std::exp(-(dx * dx + dy * dy) / (2.0 * (sigma * sigma)));
            }
            out_host[i * pw + j] = v;
        }
    return n;
}

}  // extern "C"
