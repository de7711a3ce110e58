// Host-side helpers shared by the translation units of libpose_b200.so (api_core.cu owns the state; the others include this).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/pose_b200.h"
#include "common.cuh"

namespace pose_host {

// thread-local error text + code behind pose_b200_last_error(); returns `code`
int fail(int code, const char* fmt, ...);
int last_code();
// counts the launch, turns a pending CUDA launch error into a return code
int check_launch(const char* what);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline pose::FastDiv make_div(int d) {
    pose::FastDiv f;
    f.d = (uint32_t)d;
    f.m = d > 1 ? (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d) : 0u;
    return f;
}

int current_device();
int sm_count();

// ---- launch-configuration cache.  The occupancy query, the device-attribute query and the opt-in to large dynamic shared
// memory are properties of (kernel, device, block size, shared memory): they are resolved once and kept, so a steady-state
// call is argument checks + the launches and nothing else (no driver queries, no getenv).
struct CfgKey {
    const void* fn; int dev, threads; size_t smem;
    bool operator==(const CfgKey& o) const { return fn == o.fn && dev == o.dev && threads == o.threads && smem == o.smem; }
};
bool cfg_lookup(const CfgKey& key, int* resident);                    // cached resident-CTA count of a configuration
void cfg_store(const CfgKey& key, int resident);
// the dynamic-shared-memory attribute is per kernel, not per launch, and is only ever raised: true when `smem` exceeds what
// was opted in to so far (the caller then sets the attribute and reports back with cfg_dyn_smem_set)
bool cfg_dyn_smem_needs_raise(const void* fn, int dev, size_t smem);
void cfg_dyn_smem_set(const void* fn, int dev, size_t smem);

// resident CTAs of `kern` on the current device; the first call for a configuration also opts the kernel in to `smem` bytes
// of dynamic shared memory (static + dynamic above 48 KB needs it) -- returns 0 and sets the error text when that fails
template <typename Kern>
int resident_ctas(Kern kern, int threads, size_t smem, const char* what) {
    const CfgKey key{reinterpret_cast<const void*>(kern), current_device(), threads, smem};
    int cached = 0;
    if (cfg_lookup(key, &cached)) return cached;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e == cudaSuccess && fa.sharedSizeBytes + smem > 48 * 1024 && cfg_dyn_smem_needs_raise(key.fn, key.dev, smem)) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) cfg_dyn_smem_set(key.fn, key.dev, smem);
    }
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess || per_sm < 1) {
        fail(e != cudaSuccess ? (int)e : POSE_EINVAL, "%s: kernel cannot be configured (%d threads, %zu B shared memory): %s", what, threads, smem,
             e != cudaSuccess ? cudaGetErrorString(e) : "no CTA fits on an SM");
        cudaGetLastError();
        return 0;                                         // not cached: the caller reports the failure every time
    }
    const int total = per_sm * sm_count();
    cfg_store(key, total);
    return total;
}

// persistent grid: resident CTAs (occupancy x SM count, cached), capped by the amount of work; 0 = configuration failed
template <typename Kern>
int persistent_grid(Kern kern, int threads, size_t smem, long long work_ctas, const char* what = "launch") {
    long long g = resident_ctas(kern, threads, smem, what);
    if (g <= 0) return 0;
    if (g > pose::kMaxPartialBlocks) g = pose::kMaxPartialBlocks;
    if (g > work_ctas) g = work_ctas;
    return (int)(g < 1 ? 1 : g);
}

// launch with programmatic stream serialisation: the grid may be scheduled while its predecessor drains; the kernel
// itself calls griddepcontrol.wait before touching the predecessor's results
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace pose_host
