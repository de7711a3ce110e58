// Shared device helpers for the heatmap hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define POSE_WARP 32
#define FULL_MASK 0xffffffffu

namespace pose {

constexpr int kTemplatePad = 4;                // zero columns either side of the padded Gaussian template (sbp_kernels.cuh)
constexpr int kMaxPartialBlocks = 148 * 16;    // upper bound on a persistent grid (sizes the loss-partials workspaces)

// ---------------------------------------------------------------- memory: 128-bit streaming accesses
// Every byte of logits / target / dlogits is touched exactly once per pass, so loads bypass L1
// allocation and stores are streaming: nothing here is worth keeping in L1.
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// second look at data the same warp has just streamed: keep it cacheable (L1/L2 hit expected)
__device__ __forceinline__ float4 ldg_cached(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float ldg_cached(const float* p) { return __ldg(p); }

__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(float* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// pull the line holding p into L2 ahead of use (no register, no scoreboard): turns a DRAM miss into an L2 hit for
// kernels whose per-unit compute phase is long enough to expose the load latency
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------- programmatic dependent launch (sm_90+)
// Primary grid: allow the next grid in the stream to be scheduled early.  Secondary grid: block until every
// prerequisite grid has completed and flushed its memory.  Both are no-ops for launches without the PDL attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- activations
// sigmoid(x) = 1 / (1 + 2^(-x*log2 e)) on the SFU: one MUFU.EX2 + one MUFU.RCP.
// Max error a few ulp (inside the 1e-5 parity tolerance); saturates cleanly: x -> +inf gives 1,
// x -> -inf gives +0 (ex2 -> +inf, rcp(+inf) = 0).  The .ftz forms drop the denormal fix-up code
// (6 extra instructions per element): results below 2^-126 (x < -87.3) flush to +0.  The same function is used by loss and decode so
// that "argmax of sigmoid" means the same thing in every kernel.  Monotonicity over all fp32
// inputs is verified exhaustively on the device (pose_sigmoid_monotone_check).
__device__ __forceinline__ float sigmoid_fast(float x) {
    float e, r;
    float t = x * -1.4426950408889634f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// Sigmoids of 4 logits with ONE reciprocal (5 SFU operations instead of 8): with d_i = 1 + 2^(-x_i log2 e),
//   r = 1 / (d0 d1 d2 d3),  s0 = r d1 (d2 d3),  s1 = r d0 (d2 d3),  s2 = r (d0 d1) d3,  s3 = r (d0 d1) d2.
// The exponent is clamped at 30 (x < -20.8 reads as x = -20.8, an absolute error below 1e-9) so that the product of four
// denominators stays finite.  A few ulp less accurate than sigmoid_fast (3 more roundings), far inside the 1e-5 tolerance of
// the loss; used by the read-only loss kernels, which are bound by the SFU pipe, never by the decoders.
__device__ __forceinline__ void sigmoid_fast4(const float (&x)[4], float (&s)[4]) {
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float e;
        const float t = fminf(x[j] * -1.4426950408889634f, 30.0f);
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
        d[j] = 1.0f + e;
    }
    const float p01 = d[0] * d[1], p23 = d[2] * d[3];
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
    const float a = r * p23, b = r * p01;
    s[0] = a * d[1]; s[1] = a * d[0]; s[2] = b * d[3]; s[3] = b * d[2];
}
// SHARE: one reciprocal per 128-bit vector (which kernels ask for it: POSE_FUSED_SHARE_* in sbp_kernels.cuh)
template <int V, bool SHARE>
__device__ __forceinline__ void sigmoid_vec(const float (&x)[V], float (&s)[V]) {
    if constexpr (V == 4 && SHARE) {
        sigmoid_fast4(x, s);
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) s[j] = sigmoid_fast(x[j]);
    }
}

// ---------------------------------------------------------------- the reference's sigmoid, bit for bit
// "argmax of sigmoid(x), first row-major index among equal values" (utils/sbp_utils.py:73-78) depends on WHICH fp32
// sigmoid is meant: fp32 sigmoid is many-to-one and two implementations that differ by one ulp merge different
// neighbours.  sigmoid_fast above is neither of torch's, so decode never decides a tie with it: candidates are found
// in logit space (sigmoid_window_lo) and ranked with one of the two functions below, selected by the caller.
//
// POSE_SIGMOID_ATEN_CPU: torch.sigmoid on a contiguous CPU tensor = 1 / (1 + Sleef_expf_u10(-x)) in the AVX2 / AVX512
// kernels of ATen (identical results in both; verified against torch 2.11 for all 2^32 inputs with the C restatement kept
// beside the tests, which this function follows operation for operation; every operation is IEEE so the device
// reproduces it exactly).  [Elements of a tail shorter than two vectors go through glibc expf in ATen; the reference's
// shapes (K*H*W a multiple of 32) have no tail.]
__device__ __forceinline__ float sleef_expf_u10(float d) {
    const int q = __float2int_rn(__fmul_rn(d, 1.442695040888963407359924681001892137426645954152985934135449406931f));
    const float qf = (float)q;
    float s = __fmaf_rn(qf, -0.693145751953125f, d);
    s = __fmaf_rn(qf, -1.428606765330187045e-06f, s);
    float u = 0.000198527617612853646278381f;
    u = __fmaf_rn(u, s, 0.00139304355252534151077271f);
    u = __fmaf_rn(u, s, 0.00833336077630519866943359f);
    u = __fmaf_rn(u, s, 0.0416664853692054748535156f);
    u = __fmaf_rn(u, s, 0.166666671633720397949219f);
    u = __fmaf_rn(u, s, 0.5f);
    u = __fadd_rn(1.0f, __fmaf_rn(__fmul_rn(s, s), u, s));
    const int qh = q >> 1;
    u = __fmul_rn(__fmul_rn(u, __int_as_float((qh + 0x7f) << 23)), __int_as_float((q - qh + 0x7f) << 23));
    if (d < -104.0f) u = 0.0f;
    if (100.0f < d) u = INFINITY;
    return u;
}
__device__ __forceinline__ float sigmoid_aten_cpu(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, sleef_expf_u10(-x)));
}
// POSE_SIGMOID_ATEN_CUDA: torch.sigmoid on a CUDA tensor = 1 / (1 + expf(-x)) compiled without fast-math
// (ATen/native/cuda/UnarySpecialOpsKernel.cu): libdevice expf, IEEE add and divide.  This file is compiled without
// --use_fast_math, so the same three operations are what nvcc emits here (a GPU test compares with torch bit for bit).
__device__ __forceinline__ float sigmoid_aten_cuda(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}
constexpr int kSigmoidAtenCpu = 0;
constexpr int kSigmoidAtenCuda = 1;
// runtime selection: evaluated a handful of times per map (never in the streaming loop), so a uniform branch is free
__device__ __forceinline__ float sigmoid_ref(float x, int ref) { return ref == kSigmoidAtenCuda ? sigmoid_aten_cuda(x) : sigmoid_aten_cpu(x); }

// Lower end of the candidate window of a map whose largest logit is m: every x < sigmoid_window_lo(m) has
// sigmoid_ref(x) < sigmoid_ref(m) for both references, so only elements >= lo can be the reference's argmax.
// Both references are sigma(x) (1 + eta), |eta| <= d (1 - sigma(x)) + 2^-22 with d = 2^-21 (exp within 4 ulp -- Sleef u10:
// 1, libdevice expf: 2 -- one rounding for 1 + e and one for the quotient).  With q = 1 - sigma(m), w = m - x <= 1/2:
// ln sigma(m) - ln sigma(x) >= w q and 1 - sigma(x) <= q e^w, so w q > d q (1 + e^w) + 2^-21 suffices:
// w = 2.7 d + 2^-21 / q (times a 1.125 safety factor; q from the fast sigmoid, relative error ~1e-6).  Where that exceeds 1/2
// (m > 13.8: the sigmoid is within a few ulp of 1) the window is everything above 13: sigma(13.8) - sigma(13) is 21 ulp.
__device__ __forceinline__ float sigmoid_window_lo(float m) {
    // 1/q = 1 + e^m exactly, so w = 1.125 d (2.7 + 1 + e^m): one MUFU.EX2 and an FMA (no reciprocal, no division -- this sits
    // on the dependent chain at the end of every map)
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(m * 1.4426950408889634f));
    const float w = 1.125f * 4.76837158e-7f * (3.7f + e);
    return (w < 0.5f) ? __fsub_rd(m, w) : 13.0f;
}

// ---------------------------------------------------------------- warp reductions
// (sm_100a: one CREDUX.MAX.F32 instead of five shuffle + FMNMX rounds; NaNs are ignored like fmaxf does)
__device__ __forceinline__ float warp_max(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
// largest 64-bit key of the warp, in every lane: two 32-bit CREDUX.MAX (high words, then the low words of the lanes that hold the
// largest high word)
__device__ __forceinline__ unsigned long long warp_max_key(unsigned long long k) {
    const unsigned hi = __reduce_max_sync(FULL_MASK, (unsigned)(k >> 32));
    const unsigned lo = __reduce_max_sync(FULL_MASK, (unsigned)(k >> 32) == hi ? (unsigned)k : 0u);
    return ((unsigned long long)hi << 32) | (unsigned long long)lo;
}
__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
// argmax with first-index tie-break: larger value wins, equal values -> smaller index wins.
__device__ __forceinline__ void warp_argmax_first(float& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(FULL_MASK, v, o);
        int oi = __shfl_xor_sync(FULL_MASK, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// ---------------------------------------------------------------- loss reduction (deterministic, fixed order, no float atomics)
// sum `n` (a, b) fp64 pairs, `stride` doubles apart, then loss = (w0*A + w1*B) * inv_norm.  Runs in one CTA.
__device__ __forceinline__ void reduce_pairs_cta(const double* __restrict__ pairs, int n, long long stride, double w0, double w1,
                                                 double inv_norm, float* __restrict__ loss_out, double* __restrict__ num_out) {
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { a += __ldcg(pairs + i * stride); b += __ldcg(pairs + i * stride + 1); }
    sa[threadIdx.x] = a; sb[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (num_out) { num_out[0] = sa[0]; num_out[1] = sb[0]; }
        if (loss_out) loss_out[0] = (float)((w0 * sa[0] + w1 * sb[0]) * inv_norm);
    }
}

// Two-level form for the per-map pairs of the fused kernel (69 632 of them at B = 4096): slice CTA r sums pairs
// [r*kReduceSlice, (r+1)*kReduceSlice) -- 8 independent 16-byte loads per thread, then a fixed tree -- into slices[r]; the
// slice CTA that finishes LAST (a ticket counter, zeroed by the fused kernel) returns true and goes on to add the R slice sums in
// index order.  Which CTA is last varies from run to run; the order of every addition does not.
constexpr int kReduceSlice = 2048;
__host__ __device__ inline int reduce_slices(long long n_pairs) { return n_pairs <= kReduceSlice ? 1 : (int)((n_pairs + kReduceSlice - 1) / kReduceSlice); }

__device__ __forceinline__ bool reduce_slice_and_elect(const double* __restrict__ pairs, long long n, double* __restrict__ slices,
                                                       unsigned int* __restrict__ ticket, int R, int r) {
    __shared__ double ta[256], tb[256];
    __shared__ int s_last;
    const double2* p2 = reinterpret_cast<const double2*>(pairs);
    const long long i0 = (long long)r * kReduceSlice + threadIdx.x;
    double2 v[kReduceSlice / 256];
#pragma unroll
    for (int k = 0; k < kReduceSlice / 256; ++k) {
        const long long i = i0 + 256 * k;
        v[k] = i < n ? __ldcg(p2 + i) : make_double2(0.0, 0.0);
    }
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < kReduceSlice / 256; ++k) { a += v[k].x; b += v[k].y; }
    ta[threadIdx.x] = a; tb[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { ta[threadIdx.x] += ta[threadIdx.x + s]; tb[threadIdx.x] += tb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        slices[2 * r] = ta[0];
        slices[2 * r + 1] = tb[0];
        int last = 1;
        if (R > 1) {
            __threadfence();                                   // slice sum visible device-wide before the ticket is taken
            last = atomicAdd(ticket, 1u) == (unsigned)(R - 1);
            if (last) __threadfence();                         // ... and the others' sums before they are read
        }
        s_last = last;
    }
    __syncthreads();
    return s_last != 0;
}

// ---------------------------------------------------------------- order-preserving float <-> uint key
__device__ __forceinline__ uint32_t float_key(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

// exact integer division by a runtime constant for small operands (n < 2^16 * d is plenty here):
// q = (n * m) >> 32 with m = ceil(2^32 / d); valid for n*d < 2^32 -- checked on the host.
struct FastDiv {
    uint32_t d, m;
};
__device__ __forceinline__ uint32_t fdiv(uint32_t n, FastDiv f) { return f.d == 1 ? n : __umulhi(n, f.m); }

}  // namespace pose

// ---------------------------------------------------------------- bulk async copies (TMA engine, 1-D) + mbarrier
// cp.async.bulk moves a contiguous 16-byte-multiple block between global and shared memory without occupying
// registers or LSU issue slots; completion of a load is signalled on an mbarrier (complete_tx), stores are tracked
// in per-thread bulk groups.  SASS: UBLKCP / SYNCS.
namespace pose {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "POSE_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra POSE_MBAR_DONE;\n"
        "bra POSE_MBAR_WAIT;\n"
        "POSE_MBAR_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// make this thread's generic-proxy shared-memory writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait_parity(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}


// ---------------------------------------------------------------- multi-GPU exchange: control block and flags
constexpr int kMaxPeers = 16;
constexpr int kExchangeSlots = 4;

struct ExchangeCtrl {
    unsigned long long step;       // newest step whose flag this rank has published (= completed epilogues made visible)
    unsigned int ticket;           // (unused, kept for layout stability)
    unsigned int error;            // set when a wait timed out (a peer never signalled)
    unsigned long long launched;   // in-band mode: fused kernels started so far (= index of the step being produced)
};

// what a producer kernel needs to publish "all my stores of step n have landed on every rank"
struct ExchangePub {
    int world, rank;               // world == 0: no exchange
    unsigned long long off_ctrl, off_flags;
    unsigned char* peer[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// In-band (deferred) mode, executed by ONE thread at the start of the fused kernel of step n+1: the epilogue grid of
// step n finished before this grid started (stream order), so all of its peer stores are performed -- publish flag n on
// every rank with a system-scope release store, then open step n+1.
__device__ __forceinline__ void exchange_open_step(const ExchangePub& X) {
    ExchangeCtrl* ctrl = reinterpret_cast<ExchangeCtrl*>(X.peer[X.rank] + X.off_ctrl);
    const unsigned long long n = ctrl->launched;
    if (n > 0 && ctrl->step < n) {
        for (int r = 0; r < X.world; ++r)
            st_release_sys(reinterpret_cast<unsigned long long*>(X.peer[r] + X.off_flags) + X.rank, n);
        ctrl->step = n;
    }
    ctrl->launched = n + 1;
}

}  // namespace pose
