// extern "C" boundary of libpose_b200.so (see include/pose_b200.h for the contract).
#include <atomic>
#include <cmath>
#include <mutex>
#include <unordered_map>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/pose_b200.h"
#include "sbp_kernels.cuh"
#include "sbp_tma_kernels.cuh"
#include "exchange_kernels.cuh"
#include "spm_kernels.cuh"
#include "oks_kernels.cuh"

namespace {

thread_local char g_err[512] = "";
thread_local int g_err_code = 0;
std::atomic<unsigned long long> g_launches{0};

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

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

pose::FastDiv make_div(int d) {
    pose::FastDiv f;
    f.d = (uint32_t)d;
    f.m = d > 1 ? (uint32_t)(((1ull << 32) + (uint64_t)d - 1) / (uint64_t)d) : 0u;
    return f;
}

// ---- launch-configuration cache.  The occupancy query, the device-attribute query and the opt-in to large dynamic shared
// memory are properties of (kernel, device, block size, shared memory): they are resolved once and kept, so a steady-state
// call is argument checks + the launches and nothing else (no driver queries, no getenv).
struct CfgKey {
    const void* fn; int dev, threads; size_t smem;
    bool operator==(const CfgKey& o) const { return fn == o.fn && dev == o.dev && threads == o.threads && smem == o.smem; }
};
struct CfgHash {
    size_t operator()(const CfgKey& k) const {
        return std::hash<const void*>()(k.fn) ^ (std::hash<size_t>()(k.smem) * 1000003u) ^ ((size_t)k.dev << 20) ^ (size_t)k.threads;
    }
};
std::mutex g_cfg_mutex;
std::unordered_map<CfgKey, int, CfgHash> g_resident;      // -> resident CTAs on the whole device
std::unordered_map<CfgKey, size_t, CfgHash> g_dyn_smem;   // (kernel, device, 0, 0) -> largest dynamic shared memory opted in to so far
int g_sms[64];                                            // SM count per device ordinal (0: not queried yet)

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

// resident CTAs of `kern` on the current device; the first call for a configuration also opts the kernel in to `smem` bytes
// of dynamic shared memory (static + dynamic above 48 KB needs it) -- returns 0 and sets the error text when that fails
template <typename Kern>
int resident_ctas(Kern kern, int threads, size_t smem, const char* what) {
    const CfgKey key{reinterpret_cast<const void*>(kern), current_device(), threads, smem};
    {
        std::lock_guard<std::mutex> lock(g_cfg_mutex);
        auto it = g_resident.find(key);
        if (it != g_resident.end()) return it->second;
    }
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e == cudaSuccess && fa.sharedSizeBytes + smem > 48 * 1024) {
        // the attribute is per kernel, not per launch: only ever raise it
        std::lock_guard<std::mutex> lock(g_cfg_mutex);
        size_t& have = g_dyn_smem[CfgKey{key.fn, key.dev, 0, 0}];
        if (smem > have) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) have = smem;
        }
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
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    g_resident[key] = total;
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
void launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

long long exchange_timeout_cycles() {
    // A peer that NEVER signals must not hang this GPU for good, but ranks legitimately drift by many seconds (checkpoint
    // writes, a dataloader stall, rank-0-only validation): the wait gives up after ~10 minutes of SM clock, poisons the
    // loss with NaN and sets ctrl->error, which PeerExchange.flush() / gathered_*() turn into an exception.
    // POSE_B200_EXCHANGE_TIMEOUT_CYCLES overrides (read once).
    static const long long cycles = [] {
        const char* e = getenv("POSE_B200_EXCHANGE_TIMEOUT_CYCLES");
        return e ? atoll(e) : 1200000000000ll;
    }();
    return cycles;
}

pose::ExchangeDev to_dev(const pose_exchange_t& x, double w0 = 0.0, double w1 = 0.0, double inv_norm = 0.0) {
    pose::ExchangeDev d;
    memset(&d, 0, sizeof(d));
    d.world = x.world; d.rank = x.rank; d.B = x.batch_local; d.K = x.num_keypoints; d.row_stride = x.row_stride;
    for (int r = 0; r < x.world && r < pose::kMaxPeers; ++r) d.peer[r] = reinterpret_cast<unsigned char*>(x.peer_base[r]);
    d.off_ctrl = x.off_ctrl; d.off_flags = x.off_flags;
    for (int p = 0; p < POSE_EXCHANGE_SLOTS; ++p) { d.off_rows[p] = x.off_rows[p]; d.off_nums[p] = x.off_nums[p]; d.off_ids[p] = x.off_ids[p]; }
    d.ids_local = x.ids_local;
    d.mc = reinterpret_cast<unsigned char*>(x.multicast_base);
    d.defer = x.defer; d.loss_prev = x.loss_prev; d.w0 = w0; d.w1 = w1; d.inv_norm_global = inv_norm;
    d.timeout_cycles = exchange_timeout_cycles();
    return d;
}

int check_map_shape(int N, int K, int H, int W) {
    if (N < 0 || K <= 0 || H <= 0 || W <= 0) return fail(POSE_EINVAL, "bad shape N=%d K=%d H=%d W=%d", N, K, H, W);
    if ((long long)H * W >= (1ll << 20) || W >= (1 << 11)) return fail(POSE_EINVAL, "map too large: H=%d W=%d", H, W);
    return POSE_OK;
}

}  // namespace

namespace {
template <int V, int TGT, bool GRAD, bool WTGT, bool DEC>
int launch_fused(const pose::SbpFusedParams& P0, size_t smem, cudaStream_t st, int* grid_out) {
    pose::SbpFusedParams P = P0;
    const long long ctas = (P.n_maps + pose::kSbpWarps - 1) / pose::kSbpWarps;
    const int grid = persistent_grid(pose::sbp_fused_kernel<V, TGT, GRAD, WTGT, DEC>, pose::kSbpThreads, smem, ctas, "sbp_fused");
    if (grid == 0) return last_code();
    pose::sbp_fused_kernel<V, TGT, GRAD, WTGT, DEC><<<grid, pose::kSbpThreads, smem, st>>>(P);
    *grid_out = grid;
    return check_launch("sbp_fused");
}
template <bool GRAD, bool DEC>
int launch_fused_tma(const pose::SbpFusedParams& P0, cudaStream_t st, int* grid_out) {
    pose::SbpFusedParams P = P0;
    const size_t smem = pose::sbp_tma_smem_bytes(P.lut_n);
    const long long ctas = (P.n_maps + pose::kTmaWarps - 1) / pose::kTmaWarps;
    const int grid = persistent_grid(pose::sbp_fused_tma_kernel<GRAD, DEC>, pose::kTmaThreads, smem, ctas, "sbp_fused(tma)");
    if (grid == 0) return last_code();
    pose::sbp_fused_tma_kernel<GRAD, DEC><<<grid, pose::kTmaThreads, smem, st>>>(P);
    *grid_out = grid;
    return check_launch("sbp_fused_tma");
}

template <int V, int TGT>
int dispatch_fused(const pose::SbpFusedParams& P, unsigned flags, size_t smem, cudaStream_t st, int* grid) {
    const bool g = flags & POSE_F_GRAD, t = (flags & POSE_F_TARGET_OUT) && TGT == pose::TGT_RENDER, d = flags & POSE_F_DECODE;
#define POSE_CASE(G, T, D) \
    if (g == G && t == T && d == D) return launch_fused<V, TGT, G, T, D>(P, smem, st, grid);
    POSE_CASE(false, false, false) POSE_CASE(true, false, false) POSE_CASE(false, false, true) POSE_CASE(true, false, true)
    if (TGT == pose::TGT_RENDER) {
        POSE_CASE(false, true, false) POSE_CASE(true, true, false) POSE_CASE(false, true, true) POSE_CASE(true, true, true)
    }
#undef POSE_CASE
    return fail(POSE_EINVAL, "sbp_fused: unsupported flag combination 0x%x", flags);
}
}  // namespace

extern "C" {

#ifndef POSE_B200_SOURCE_HASH_STR
#define POSE_B200_SOURCE_HASH_STR "unhashed-build"
#endif
// searched for in the file by build.py / _cabi.py (no dlopen needed): the hash of the sources this binary was compiled from
static const char g_source_hash[] = "POSE_B200_SOURCE_HASH=" POSE_B200_SOURCE_HASH_STR;

int pose_b200_version(void) { return 200; }
const char* pose_b200_source_hash(void) { return g_source_hash + sizeof("POSE_B200_SOURCE_HASH=") - 1; }
const char* pose_b200_last_error(void) { return g_err; }
unsigned long long pose_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

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

int pose_sbp_render(const void* kp, int kp_dtype, float* target, int N, int K, int H, int W, double sigma,
                    const float* lut, int lut_n, pose_stream_t stream) {
    if (int rc = check_map_shape(N, K, H, W)) return rc;
    if (N == 0) return POSE_OK;
    if (!kp || !target || !lut || lut_n <= 0 || lut_n > 64 || !(sigma > 0.0)) return fail(POSE_EINVAL, "sbp_render: bad argument");
    pose::SbpRenderParams P;
    P.kp = kp; P.kp_f64 = kp_dtype == POSE_KP_F64; P.target = target; P.lut = lut; P.lut_n = lut_n;
    P.three_sigma = 3 * sigma;
    P.n_maps = (long long)N * K; P.H = H; P.W = W; P.HW = H * W; P.divW = make_div(W);
    const size_t smem = (size_t)lut_n * lut_n * sizeof(float);
    const bool vec = (P.HW % 4 == 0) && W >= 4 && aligned16(target);
    const long long ctas = (P.n_maps + pose::kSbpWarps - 1) / pose::kSbpWarps;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) {
        const int grid = persistent_grid(pose::sbp_render_kernel<4>, pose::kSbpThreads, smem, ctas, "sbp_render");
        if (grid == 0) return last_code();
        pose::sbp_render_kernel<4><<<grid, pose::kSbpThreads, smem, st>>>(P);
    } else {
        const int grid = persistent_grid(pose::sbp_render_kernel<1>, pose::kSbpThreads, smem, ctas, "sbp_render");
        if (grid == 0) return last_code();
        pose::sbp_render_kernel<1><<<grid, pose::kSbpThreads, smem, st>>>(P);
    }
    return check_launch("sbp_render");
}

unsigned long long pose_sbp_fused_workspace_bytes(void) { return (unsigned long long)pose::kMaxPartialBlocks * 2 * sizeof(double); }


int pose_sbp_fused(const float* logits, const float* target_in, const void* kp, int kp_dtype, double sigma,
                   const float* lut, int lut_n, float* dlogits, float* target_out, float* loss_out,
                   double* loss_num_out, float* joints, float conf_threshold, float coord_scale, int N, int K,
                   int H, int W, float lambda_pos, float lambda_neg, double inv_norm, unsigned flags,
                   const double* bbox, float* packed_out, int input_h, int input_w,
                   const struct pose_exchange* exchange,
                   void* workspace, unsigned long long workspace_bytes, pose_stream_t stream) {
    if (int rc = check_map_shape(N, K, H, W)) return rc;
    if (!logits) return fail(POSE_EINVAL, "sbp_fused: logits is NULL");
    if ((target_in != nullptr) == (kp != nullptr)) return fail(POSE_EINVAL, "sbp_fused: pass exactly one of target_in / kp");
    if (kp && (!lut || lut_n <= 0 || lut_n > 64 || !(sigma > 0.0))) return fail(POSE_EINVAL, "sbp_fused: render mode needs sigma>0 and a template");
    if ((flags & POSE_F_GRAD) && !dlogits) return fail(POSE_EINVAL, "sbp_fused: POSE_F_GRAD without dlogits");
    if ((flags & POSE_F_TARGET_OUT) && (!target_out || !kp)) return fail(POSE_EINVAL, "sbp_fused: POSE_F_TARGET_OUT needs target_out and kp");
    if ((flags & POSE_F_DECODE) && !joints) return fail(POSE_EINVAL, "sbp_fused: POSE_F_DECODE without joints");
    if (!loss_out && !loss_num_out) return fail(POSE_EINVAL, "sbp_fused: no loss output");
    if (packed_out && !bbox) return fail(POSE_EINVAL, "sbp_fused: bbox and packed_out go together");
    if (bbox && !packed_out && !exchange) return fail(POSE_EINVAL, "sbp_fused: bbox and packed_out go together");
    if (exchange) {
        if (!bbox) return fail(POSE_EINVAL, "sbp_fused: the exchange needs bbox (back-projected rows are what is exchanged)");
        if (exchange->world < 1 || exchange->world > POSE_MAX_PEERS || exchange->rank < 0 || exchange->rank >= exchange->world ||
            exchange->batch_local != N || exchange->num_keypoints != K || !exchange->ids_local ||
            exchange->row_stride != (3 * K + 1 + 3) / 4 * 4 || exchange->row_stride > pose::kMaxRowStride ||
            exchange->defer < 0 || exchange->defer > 1 || (exchange->defer && !exchange->loss_prev))
            return fail(POSE_EINVAL, "sbp_fused: bad exchange descriptor");
    }
    if (bbox && (!(flags & POSE_F_DECODE) || input_h <= 0 || input_w <= 0)) return fail(POSE_EINVAL, "sbp_fused: back-projection needs POSE_F_DECODE and the input size");
    if (!workspace || workspace_bytes < pose_sbp_fused_workspace_bytes()) return fail(POSE_EWORKSPACE, "sbp_fused: workspace too small");
    if (!aligned16(workspace)) return fail(POSE_EALIGN, "sbp_fused: workspace must be 16-byte aligned");
    if (dlogits == logits || (target_out && target_out == logits)) return fail(POSE_EINVAL, "sbp_fused: outputs must not alias logits");
    cudaStream_t st = (cudaStream_t)stream;

    pose::SbpFusedParams P;
    memset(&P, 0, sizeof(P));
    P.logits = logits; P.target_in = target_in; P.kp = kp; P.kp_f64 = kp_dtype == POSE_KP_F64;
    P.lut = lut; P.lut_n = lut_n; P.three_sigma = 3 * sigma;
    P.dlogits = dlogits; P.target_out = target_out; P.joints = joints;
    P.partials = reinterpret_cast<double*>(workspace);
    P.thr = conf_threshold; P.scale = coord_scale;
    P.gpos = (float)(2.0 * (double)lambda_pos * inv_norm);
    P.gneg = (float)(2.0 * (double)lambda_neg * inv_norm);
    P.n_maps = (long long)N * K; P.H = H; P.W = W; P.HW = H * W; P.divW = make_div(W);
    P.sig_ref = (flags & POSE_F_SIGMOID_CUDA) ? POSE_SIGMOID_ATEN_CUDA : POSE_SIGMOID_ATEN_CPU;
    if (exchange && exchange->defer) {           // in-band mode: the fused kernel publishes the previous step and opens this one
        if (N == 0) return fail(POSE_EINVAL, "sbp_fused: the in-band exchange needs a non-empty shard");
        P.xpub.world = exchange->world; P.xpub.rank = exchange->rank;
        P.xpub.off_ctrl = exchange->off_ctrl; P.xpub.off_flags = exchange->off_flags;
        for (int r = 0; r < exchange->world; ++r) P.xpub.peer[r] = reinterpret_cast<unsigned char*>(exchange->peer_base[r]);
    }

    int grid = 0;
    if (N > 0) {
        // rendered target, read-only variants: a 128-bit vector must lie in one row (the padded-template lookup of render_loss_vec)
        const bool vec = (P.HW % 4 == 0) && W >= 4 && (!kp || (flags & POSE_F_GRAD) || W % 4 == 0) && aligned16(logits) && (!target_in || aligned16(target_in)) &&
                         (!(flags & POSE_F_GRAD) || aligned16(dlogits)) && (!(flags & POSE_F_TARGET_OUT) || aligned16(target_out));
        const size_t smem = kp ? pose::lut_padded_floats(lut_n) * sizeof(float) : 0;
        int rc;
        const bool tma = (flags & POSE_F_TMA) && kp && vec && !(flags & POSE_F_TARGET_OUT) && lut_n <= 31;
        if (tma) {
            const bool g = flags & POSE_F_GRAD, d = flags & POSE_F_DECODE;
            rc = g ? (d ? launch_fused_tma<true, true>(P, st, &grid) : launch_fused_tma<true, false>(P, st, &grid))
                   : (d ? launch_fused_tma<false, true>(P, st, &grid) : launch_fused_tma<false, false>(P, st, &grid));
        } else if (kp) rc = vec ? dispatch_fused<4, pose::TGT_RENDER>(P, flags, smem, st, &grid) : dispatch_fused<1, pose::TGT_RENDER>(P, flags, smem, st, &grid);
        else rc = vec ? dispatch_fused<4, pose::TGT_DENSE>(P, flags, smem, st, &grid) : dispatch_fused<1, pose::TGT_DENSE>(P, flags, smem, st, &grid);
        if (rc) return rc;
    }
    pose::SbpEpilogueParams E;
    E.partials = P.partials; E.nblocks = grid; E.w0 = (double)lambda_pos; E.w1 = (double)lambda_neg; E.inv_norm = inv_norm;
    E.loss_out = loss_out; E.num_out = loss_num_out;
    E.joints = joints; E.bbox = bbox; E.packed = packed_out; E.N = bbox ? N : 0; E.K = K; E.in_h = (double)input_h; E.in_w = (double)input_w;
    const unsigned bp_ctas = bbox ? (unsigned)(((long long)N * 32 + 255) / 256) : 0u;
    if (exchange) {
        launch_pdl(pose::sbp_epilogue_p2p_kernel, bp_ctas + 1u, 256u, st, E, to_dev(*exchange, (double)lambda_pos, (double)lambda_neg, inv_norm));
        return check_launch("sbp_epilogue_p2p");
    }
    launch_pdl(pose::sbp_epilogue_kernel, bp_ctas + 1u, 256u, st, E);
    return check_launch("sbp_epilogue");
}

unsigned long long pose_exchange_layout(pose_exchange_t* x) {
    if (!x || x->world < 1 || x->world > POSE_MAX_PEERS || x->batch_local < 0 || x->num_keypoints <= 0) return 0ull;
    auto up = [](unsigned long long v) { return (v + 255ull) / 256ull * 256ull; };
    x->row_stride = (3 * x->num_keypoints + 1 + 3) / 4 * 4;
    if (x->row_stride > pose::kMaxRowStride) return 0ull;
    const unsigned long long rows = up((unsigned long long)x->world * x->batch_local * (unsigned long long)x->row_stride * 4ull);
    const unsigned long long nums = up((unsigned long long)x->world * 16ull);
    const unsigned long long ids = up((unsigned long long)x->world * x->batch_local * 16ull);
    unsigned long long off = 0;
    x->off_ctrl = off; off += 256;
    x->off_flags = off; off += up((unsigned long long)POSE_MAX_PEERS * 8ull);
    for (int p = 0; p < POSE_EXCHANGE_SLOTS; ++p) {
        x->off_rows[p] = off; off += rows;
        x->off_nums[p] = off; off += nums;
        x->off_ids[p] = off; off += ids;
    }
    return off;
}

namespace {
int exchange_wait(const pose_exchange_t* x, int mode, double w0, double w1, double inv_norm, float* loss_out, pose_stream_t stream) {
    if (!x || !loss_out || x->world < 1 || x->world > POSE_MAX_PEERS || x->rank < 0 || x->rank >= x->world || x->defer < 0 || x->defer > 1)
        return fail(POSE_EINVAL, "exchange_finish: bad argument");
    if ((mode == 0) == (x->defer == 1)) return POSE_OK;      // in-band mode has no per-step finish; lock-step mode has nothing to flush
    launch_pdl(pose::exchange_wait_reduce_kernel, 1u, 256u, (cudaStream_t)stream, to_dev(*x), mode, w0, w1, inv_norm, loss_out,
               exchange_timeout_cycles());
    return check_launch("exchange_wait_reduce");
}
}  // namespace

int pose_exchange_finish(const pose_exchange_t* x, double w0, double w1, double inv_norm, float* loss_out, pose_stream_t stream) {
    return exchange_wait(x, 0, w0, w1, inv_norm, loss_out, stream);
}

int pose_exchange_flush(const pose_exchange_t* x, double w0, double w1, double inv_norm, float* loss_out, pose_stream_t stream) {
    return exchange_wait(x, 1, w0, w1, inv_norm, loss_out, stream);
}

int pose_loss_reduce(const double* pairs, int n, long long stride, double w0, double w1, double inv_norm, float* loss_out,
                     double* num_out, pose_stream_t stream) {
    if (!pairs || n <= 0 || stride < 2 || (!loss_out && !num_out)) return fail(POSE_EINVAL, "loss_reduce: bad argument");
    launch_pdl(pose::loss_reduce_kernel, 1u, 256u, (cudaStream_t)stream, pairs, n, stride, w0, w1, inv_norm, loss_out, num_out);
    return check_launch("loss_reduce");
}

int pose_scale_grad(float* dlogits, const float* grad_output, unsigned long long n, pose_stream_t stream) {
    if (!dlogits || !grad_output) return fail(POSE_EINVAL, "scale_grad: NULL pointer");
    if (n == 0) return POSE_OK;
    unsigned long long blocks = (n / 4 + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    pose::scale_grad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dlogits, grad_output, n);
    return check_launch("scale_grad");
}

int pose_sbp_decode(const float* x, float* joints, int N, int K, int H, int W, float conf_threshold, int apply_sigmoid,
                    float coord_scale, int refine, int sigmoid_ref, pose_stream_t stream) {
    if (int rc = check_map_shape(N, K, H, W)) return rc;
    if (sigmoid_ref != POSE_SIGMOID_ATEN_CPU && sigmoid_ref != POSE_SIGMOID_ATEN_CUDA) return fail(POSE_EINVAL, "sbp_decode: bad sigmoid_ref %d", sigmoid_ref);
    if (N == 0) return POSE_OK;
    if (!x || !joints) return fail(POSE_EINVAL, "sbp_decode: NULL pointer");
    pose::SbpDecodeParams P;
    P.x = x; P.joints = joints; P.thr = conf_threshold; P.scale = coord_scale;
    P.n_maps = (long long)N * K; P.H = H; P.W = W; P.HW = H * W; P.divW = make_div(W); P.refine = refine; P.sig_ref = sigmoid_ref;
    const bool vec = (P.HW % 4 == 0) && aligned16(x);
    const long long ctas = (P.n_maps + pose::kSbpWarps - 1) / pose::kSbpWarps;
    cudaStream_t st = (cudaStream_t)stream;
#define POSE_DEC(V, S)                                                                                        \
    {                                                                                                         \
        const int grid = persistent_grid(pose::sbp_decode_kernel<V, S>, pose::kSbpThreads, 0, ctas, "sbp_decode"); \
        if (grid == 0) return last_code();                                                                    \
        pose::sbp_decode_kernel<V, S><<<grid, pose::kSbpThreads, 0, st>>>(P);                                 \
    }
    const bool sig = apply_sigmoid != 0;
    if (vec) { if (sig) POSE_DEC(4, true) else POSE_DEC(4, false) }
    else { if (sig) POSE_DEC(1, true) else POSE_DEC(1, false) }
#undef POSE_DEC
    return check_launch("sbp_decode");
}

int pose_sbp_decode_flip(const float* x, const float* x_flip, const int* flip_perm, float* joints, int N, int K, int H, int W,
                         float conf_threshold, int apply_sigmoid, float coord_scale, int refine, pose_stream_t stream) {
    if (int rc = check_map_shape(N, K, H, W)) return rc;
    if (N == 0) return POSE_OK;
    if (!x || !x_flip || !flip_perm || !joints) return fail(POSE_EINVAL, "sbp_decode_flip: NULL pointer");
    pose::SbpDecodeFlipParams P;
    P.x = x; P.xf = x_flip; P.perm = flip_perm; P.joints = joints; P.thr = conf_threshold; P.scale = coord_scale;
    P.n_maps = (long long)N * K; P.K = K; P.H = H; P.W = W; P.HW = H * W; P.divW = make_div(W); P.divK = make_div(K); P.refine = refine;
    const bool vec = (W % 4 == 0) && aligned16(x) && aligned16(x_flip);
    const long long ctas = (P.n_maps + pose::kSbpWarps - 1) / pose::kSbpWarps;
    cudaStream_t st = (cudaStream_t)stream;
#define POSE_DECF(V, S)                                                                                \
    {                                                                                                  \
        const int grid = persistent_grid(pose::sbp_decode_flip_kernel<V, S>, pose::kSbpThreads, 0, ctas, "sbp_decode_flip"); \
        if (grid == 0) return last_code();                                                             \
        pose::sbp_decode_flip_kernel<V, S><<<grid, pose::kSbpThreads, 0, st>>>(P);                     \
    }
    const bool sig = apply_sigmoid != 0;
    if (vec) { if (sig) POSE_DECF(4, true) else POSE_DECF(4, false) }
    else { if (sig) POSE_DECF(1, true) else POSE_DECF(1, false) }
#undef POSE_DECF
    return check_launch("sbp_decode_flip");
}

int pose_sbp_backproject(const float* joints, const double* bbox, float* packed_out, int N, int K,
                         int input_h, int input_w, pose_stream_t stream) {
    if (N < 0 || K <= 0 || input_h <= 0 || input_w <= 0) return fail(POSE_EINVAL, "sbp_backproject: bad shape");
    if (N == 0) return POSE_OK;
    if (!joints || !bbox || !packed_out) return fail(POSE_EINVAL, "sbp_backproject: NULL pointer");
    const long long threads = (long long)N * 32;
    pose::sbp_backproject_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(joints, bbox, packed_out, N, K,
                                                                                                   (double)input_h, (double)input_w);
    return check_launch("sbp_backproject");
}

int pose_spm_render(const long long* centers, const long long* joints, const int* counts, float* target, int N, int Pmax,
                    int K, int R, double sigma, const float* lut, int lut_n, pose_stream_t stream) {
    if (N < 0 || Pmax < 0 || K <= 0 || R <= 0 || R % 4 != 0 || R > 2048) return fail(POSE_EINVAL, "spm_render: bad shape (R must be a multiple of 4)");
    if (!counts || !target || !lut || lut_n <= 0 || lut_n > 64 || !(sigma > 0.0) || (Pmax > 0 && (!centers || !joints)))
        return fail(POSE_EINVAL, "spm_render: bad argument");
    if (!aligned16(target)) return fail(POSE_EALIGN, "spm_render: target must be 16-byte aligned");
    if (N == 0) return POSE_OK;
    pose::SpmRenderParams P;
    P.centers = centers; P.joints = joints; P.counts = counts; P.target = target; P.lut = lut; P.lut_n = lut_n;
    P.three_sigma = 3 * sigma; P.half = (int)((6 * sigma + 2) / 2);
    P.z = std::sqrt((double)((long long)R * R + (long long)R * R));
    P.N = N; P.Pmax = Pmax; P.K = K; P.R = R;
    const int quads = R * R / 4;
    static const bool two_pass = getenv("POSE_B200_SPM_RENDER_TWO_PASS") != nullptr;      // diagnostics: force the fill + patch pair (read once)
    if (Pmax <= pose::kSpmFusedMaxPersons && !two_pass) {
        // single pass: the render-only form of the fused kernel (linear write stream, covered pixels filled in by the same pass)
        pose::SpmFusedParams F;
        memset(&F, 0, sizeof(F));
        F.target_out = target; F.centers = centers; F.joints = joints; F.counts = counts; F.lut = lut; F.lut_n = lut_n;
        F.three_sigma = P.three_sigma; F.half = P.half; F.z = P.z;
        F.N = N; F.Pmax = Pmax; F.K = K; F.R = R; F.quads = quads; F.div_qpr = make_div(R / 4);
        F.wpr = (R / 4 + 31) / 32;
        F.div_n = R <= 1024 ? 2 * R + 1 : 0;
        const size_t fsmem = pose::spm_fused_smem_bytes(F.div_n, R, K, F.wpr, lut_n);
        if (fsmem <= 200 * 1024) {
            const int fchunk = pose::kSpmStreamThreads * pose::spm_fused_u(false, true, false);
            const long long funits = (long long)N * (1 + 2 * K) * ((quads + fchunk - 1) / fchunk);
#define POSE_SPMR(RG, MP)                                                                                                      \
    {                                                                                                                          \
        const int fgrid = persistent_grid(pose::spm_fused_kernel<false, false, true, RG, MP>, pose::kSpmFusedThreads, fsmem, funits, "spm_render"); \
        if (fgrid == 0) return last_code();                                                                                    \
        pose::spm_fused_kernel<false, false, true, RG, MP><<<fgrid, pose::kSpmFusedThreads, fsmem, (cudaStream_t)stream>>>(F); \
    }
            if (pose::spm_fused_use_map(R)) { if (R % 128 == 0) POSE_SPMR(true, true) else POSE_SPMR(false, true) }
            else { if (R % 128 == 0) POSE_SPMR(true, false) else POSE_SPMR(false, false) }
#undef POSE_SPMR
            return check_launch("spm_render(single pass)");
        }
    }
    const size_t smem = (size_t)lut_n * lut_n * sizeof(float);
    const long long units = (long long)N * (1 + 2 * K) * ((quads + pose::kSpmRenderChunk - 1) / pose::kSpmRenderChunk);
    const int grid = persistent_grid(pose::spm_fill_kernel, pose::kSpmThreads, smem, units, "spm_fill");
    if (grid == 0) return last_code();
    pose::spm_fill_kernel<<<grid, pose::kSpmThreads, smem, (cudaStream_t)stream>>>(P);
    if (int rc = check_launch("spm_fill")) return rc;
    if (Pmax > 0) {
        if ((long long)N * Pmax > 0x7fffffffll) return fail(POSE_EINVAL, "spm_render: N*Pmax too large");
        // quotient table in dynamic shared memory ((2R+1) doubles): needs an explicit launch config for PDL + smem
        const int div_n = R <= 1024 ? 2 * R + 1 : 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((long long)N * Pmax));
        cfg.blockDim = dim3((unsigned)pose::kSpmThreads);
        cfg.dynamicSmemBytes = (size_t)div_n * sizeof(double);
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, pose::spm_patch_kernel, P, div_n);
    }
    return check_launch("spm_render");
}

unsigned long long pose_spm_loss_workspace_bytes(void) { return (unsigned long long)pose::kMaxPartialBlocks * 2 * sizeof(double); }

int pose_spm_loss(const float* logits, const float* target, float* dlogits, float* loss_out, double* loss_num_out, int N,
                  int K, int R, float lambda_root, float lambda_disp, double inv_norm, int write_grad, void* workspace,
                  unsigned long long workspace_bytes, pose_stream_t stream) {
    if (N < 0 || K <= 0 || R <= 0 || R % 4 != 0) return fail(POSE_EINVAL, "spm_loss: bad shape (R must be a multiple of 4)");
    if (!logits || !target || (write_grad && !dlogits) || (!loss_out && !loss_num_out)) return fail(POSE_EINVAL, "spm_loss: NULL pointer");
    if (!aligned16(logits) || !aligned16(target) || (write_grad && !aligned16(dlogits))) return fail(POSE_EALIGN, "spm_loss: tensors must be 16-byte aligned");
    if (!workspace || workspace_bytes < pose_spm_loss_workspace_bytes() || !aligned16(workspace)) return fail(POSE_EWORKSPACE, "spm_loss: workspace too small / unaligned");
    if (write_grad && dlogits == logits) return fail(POSE_EINVAL, "spm_loss: dlogits must not alias logits");
    cudaStream_t st = (cudaStream_t)stream;
    pose::SpmLossParams P;
    P.logits = logits; P.target = target; P.dlogits = dlogits; P.partials = reinterpret_cast<double*>(workspace);
    P.quads = R * R / 4; P.C = 1 + 2 * K; P.planes = (long long)N * P.C;
    P.groot = (float)(2.0 * (double)lambda_root * inv_norm);
    P.gdisp = (float)((double)lambda_disp * inv_norm);
    int grid = 0;
    if (N > 0) {
        const long long ctas = P.planes * ((P.quads + pose::kSpmLossChunk - 1) / pose::kSpmLossChunk);
        if (write_grad) {
            grid = persistent_grid(pose::spm_loss_kernel<true>, pose::kSpmThreads, 0, ctas, "spm_loss");
            if (grid == 0) return last_code();
            pose::spm_loss_kernel<true><<<grid, pose::kSpmThreads, 0, st>>>(P);
        } else {
            grid = persistent_grid(pose::spm_loss_kernel<false>, pose::kSpmThreads, 0, ctas, "spm_loss");
            if (grid == 0) return last_code();
            pose::spm_loss_kernel<false><<<grid, pose::kSpmThreads, 0, st>>>(P);
        }
        if (int rc = check_launch("spm_loss")) return rc;
    }
    launch_pdl(pose::loss_reduce_kernel, 1u, 256u, st, (const double*)P.partials, grid, 2ll, (double)lambda_root, (double)lambda_disp,
               inv_norm, loss_out, loss_num_out);
    return check_launch("loss_reduce");
}

unsigned long long pose_spm_fused_workspace_bytes(void) { return (unsigned long long)pose::kMaxPartialBlocks * 2 * sizeof(double); }

int pose_spm_fused(const float* logits, const long long* centers, const long long* joints, const int* counts, float* dlogits,
                   float* target_out, float* loss_out, double* loss_num_out, int N, int Pmax, int K, int R, double sigma,
                   const float* lut, int lut_n, float lambda_root, float lambda_disp, double inv_norm, unsigned flags,
                   void* workspace, unsigned long long workspace_bytes, pose_stream_t stream) {
    if (N < 0 || Pmax < 0 || K <= 0 || R <= 0 || R % 4 != 0 || R > 2048) return fail(POSE_EINVAL, "spm_fused: bad shape (R must be a multiple of 4, <= 2048)");
    if (Pmax > pose::kSpmFusedMaxPersons) return fail(POSE_EINVAL, "spm_fused: Pmax=%d > %d persons per image (use pose_spm_render + pose_spm_loss)", Pmax, pose::kSpmFusedMaxPersons);
    if (flags & ~(POSE_F_GRAD | POSE_F_TARGET_OUT)) return fail(POSE_EINVAL, "spm_fused: unsupported flags 0x%x", flags);
    const bool grad = flags & POSE_F_GRAD, wtgt = flags & POSE_F_TARGET_OUT;
    if (!loss_out && !loss_num_out) return fail(POSE_EINVAL, "spm_fused: no loss output");
    if (!workspace || workspace_bytes < pose_spm_fused_workspace_bytes() || !aligned16(workspace)) return fail(POSE_EWORKSPACE, "spm_fused: workspace too small / unaligned");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = 0;
    if (N > 0) {
        if (!logits || !counts || !lut || lut_n <= 0 || lut_n > 64 || !(sigma > 0.0) || (Pmax > 0 && (!centers || !joints)) ||
            (grad && !dlogits) || (wtgt && !target_out))
            return fail(POSE_EINVAL, "spm_fused: bad argument");
        if (!aligned16(logits) || (grad && !aligned16(dlogits)) || (wtgt && !aligned16(target_out))) return fail(POSE_EALIGN, "spm_fused: tensors must be 16-byte aligned");
        if ((grad && dlogits == logits) || (wtgt && target_out == logits)) return fail(POSE_EINVAL, "spm_fused: outputs must not alias logits");
        pose::SpmFusedParams P;
        memset(&P, 0, sizeof(P));
        P.logits = logits; P.dlogits = dlogits; P.target_out = target_out;
        P.centers = centers; P.joints = joints; P.counts = counts; P.lut = lut; P.lut_n = lut_n;
        P.three_sigma = 3 * sigma; P.half = (int)((6 * sigma + 2) / 2);
        P.z = std::sqrt((double)((long long)R * R + (long long)R * R));
        P.partials = reinterpret_cast<double*>(workspace);
        P.N = N; P.Pmax = Pmax; P.K = K; P.R = R; P.quads = R * R / 4; P.div_qpr = make_div(R / 4);
        P.wpr = (R / 4 + 31) / 32;
        P.div_n = R <= 1024 ? 2 * R + 1 : 0;
        P.groot = (float)(2.0 * (double)lambda_root * inv_norm);
        P.gdisp = (float)((double)lambda_disp * inv_norm);
        const size_t smem = pose::spm_fused_smem_bytes(P.div_n, R, K, P.wpr, lut_n);
        if (smem > 200 * 1024) return fail(POSE_EINVAL, "spm_fused: R=%d K=%d needs %zu bytes of shared memory (use pose_spm_render + pose_spm_loss)", R, K, smem);
        const int uchunk = pose::kSpmStreamThreads * pose::spm_fused_u(grad, wtgt);
        const long long units = (long long)N * (1 + 2 * K) * ((P.quads + uchunk - 1) / uchunk);
#define POSE_SPMF3(G, T, RG, MP)                                                                                              \
    {                                                                                                                          \
        grid = persistent_grid(pose::spm_fused_kernel<true, G, T, RG, MP>, pose::kSpmFusedThreads, smem, units, "spm_fused");  \
        if (grid == 0) return last_code();                                                                                     \
        pose::spm_fused_kernel<true, G, T, RG, MP><<<grid, pose::kSpmFusedThreads, smem, st>>>(P);                             \
    }
#define POSE_SPMF(G, T)                                                                                                        \
    {                                                                                                                          \
        if (pose::spm_fused_use_map(R)) { if (R % 128 == 0) POSE_SPMF3(G, T, true, true) else POSE_SPMF3(G, T, false, true) }  \
        else { if (R % 128 == 0) POSE_SPMF3(G, T, true, false) else POSE_SPMF3(G, T, false, false) }                           \
    }
        if (grad && wtgt) POSE_SPMF(true, true) else if (grad) POSE_SPMF(true, false) else if (wtgt) POSE_SPMF(false, true) else POSE_SPMF(false, false)
#undef POSE_SPMF3
#undef POSE_SPMF
        if (int rc = check_launch("spm_fused")) return rc;
    }
    launch_pdl(pose::loss_reduce_kernel, 1u, 256u, st, (const double*)workspace, grid, 2ll, (double)lambda_root, (double)lambda_disp,
               inv_norm, loss_out, loss_num_out);
    return check_launch("loss_reduce");
}

int pose_spm_decode(const float* x, float* roots, float* kps, int* counts, int* counts_total, int N, int Pmax, int K, int R,
                    float conf_threshold, double dist_threshold, int apply_act, int sigmoid_ref, float input_size, pose_stream_t stream) {
    if (N < 0 || Pmax <= 0 || K <= 0 || R <= 0) return fail(POSE_EINVAL, "spm_decode: bad shape");
    if (!x || !roots || !kps || !counts) return fail(POSE_EINVAL, "spm_decode: NULL pointer");
    if (!(dist_threshold >= 0.0) || dist_threshold > 1024.0) return fail(POSE_EINVAL, "spm_decode: bad dist_threshold");
    if (sigmoid_ref != POSE_SIGMOID_ATEN_CPU && sigmoid_ref != POSE_SIGMOID_ATEN_CUDA) return fail(POSE_EINVAL, "spm_decode: bad sigmoid_ref %d", sigmoid_ref);
    if (R > 8192) return fail(POSE_EINVAL, "spm_decode: R=%d too large", R);
    // suppressed-pixel bitmap of the dense-map fallback: R*R bits
    const size_t smem = ((size_t)R * R + 31) / 32 * sizeof(unsigned int);
    if (smem > 160 * 1024) return fail(POSE_EINVAL, "spm_decode: R=%d: the suppression bitmap does not fit in shared memory", R);
    if (N == 0) return POSE_OK;
    // (static + dynamic shared memory above 48 KB needs the opt-in: resolved once per (device, size) by the configuration cache)
    if (resident_ctas(pose::spm_decode_kernel, pose::kSpmDecThreads, smem, "spm_decode") == 0) return last_code();
    pose::SpmDecodeParams P;
    P.x = x; P.roots = roots; P.kps = kps; P.counts = counts; P.counts_total = counts_total;
    P.N = N; P.Pmax = Pmax; P.K = K; P.R = R; P.C = 1 + 2 * K;
    P.thr = conf_threshold; P.dist_thr = dist_threshold; P.apply_act = apply_act; P.sig_ref = sigmoid_ref;
    {   // logits that cannot pass `sigmoid(x) > thr` under either reference sigmoid (relative error < 2^-20): x <= logit(thr (1 - 4e-6))
        const double t = (double)conf_threshold;
        if (!(t > 0.0)) P.x_lo = -INFINITY;                     // thr <= 0 (or NaN): every logit is evaluated exactly
        else if (t >= 1.0) P.x_lo = INFINITY;                   // a sigmoid never exceeds 1
        else {
            const double tl = t * (1.0 - 4e-6);
            P.x_lo = std::nextafterf((float)std::log(tl / (1.0 - tl)), -INFINITY);
        }
    }
    P.zf = (float)std::sqrt((double)((long long)R * R + (long long)R * R));
    P.input_size = input_size;
    {   // nms_spm keeps candidates with sqrt(dx^2+dy^2) > dist_thr (fp64 sqrt of an integer): the same predicate as an integer bound
        long long s = (long long)std::floor(dist_threshold * dist_threshold);
        while (s > 0 && std::sqrt((double)(s - 1)) > dist_threshold) --s;
        while (!(std::sqrt((double)s) > dist_threshold)) ++s;
        P.s_min = s;
    }
    pose::spm_decode_kernel<<<N, pose::kSpmDecThreads, smem, (cudaStream_t)stream>>>(P);
    return check_launch("spm_decode");
}

int pose_spm_rescale(const float* kps, const int* counts, const long long* image_w, const long long* image_h, float* out,
                     int N, int Pmax, int K, float input_size, pose_stream_t stream) {
    if (N < 0 || Pmax <= 0 || K <= 0 || !(input_size > 0.0f)) return fail(POSE_EINVAL, "spm_rescale: bad shape");
    if (N == 0) return POSE_OK;
    if (!kps || !counts || !image_w || !image_h || !out) return fail(POSE_EINVAL, "spm_rescale: NULL pointer");
    const long long total = (long long)N * Pmax * K;
    long long blocks = (total + 255) / 256, cap = (long long)sm_count() * 8;
    pose::spm_rescale_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(kps, counts, image_w, image_h, out, N, Pmax, K,
                                                                                                     input_size);
    return check_launch("spm_rescale");
}

int pose_spm_gather(const float* roots, const float* disp, float* kps, int n_roots, int K, int R, double dist_threshold,
                    pose_stream_t stream) {
    if (n_roots < 0 || K <= 0 || R <= 0 || !disp || (n_roots > 0 && (!roots || !kps))) return fail(POSE_EINVAL, "spm_gather: bad argument");
    if (n_roots == 0) return POSE_OK;
    const float zf = (float)std::sqrt((double)((long long)R * R + (long long)R * R));
    const int total = n_roots * K;
    pose::spm_gather_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(roots, disp, kps, n_roots, K, R, zf, dist_threshold);
    return check_launch("spm_gather");
}

int pose_sigmoid_ref_eval(const float* x, float* y, unsigned long long n, int sigmoid_ref, pose_stream_t stream) {
    if (sigmoid_ref != POSE_SIGMOID_ATEN_CPU && sigmoid_ref != POSE_SIGMOID_ATEN_CUDA) return fail(POSE_EINVAL, "sigmoid_ref_eval: bad sigmoid_ref %d", sigmoid_ref);
    if (n == 0) return POSE_OK;
    if (!x || !y) return fail(POSE_EINVAL, "sigmoid_ref_eval: NULL pointer");
    pose::sigmoid_ref_eval_kernel<<<sm_count() * 8, 256, 0, (cudaStream_t)stream>>>(x, y, n, sigmoid_ref);
    return check_launch("sigmoid_ref_eval");
}

int pose_sigmoid_window_check(unsigned long long* violations_out, int sigmoid_ref, pose_stream_t stream) {
    if (!violations_out) return fail(POSE_EINVAL, "sigmoid_window_check: NULL pointer");
    if (sigmoid_ref != POSE_SIGMOID_ATEN_CPU && sigmoid_ref != POSE_SIGMOID_ATEN_CUDA) return fail(POSE_EINVAL, "sigmoid_window_check: bad sigmoid_ref %d", sigmoid_ref);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(violations_out, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return fail((int)e, "sigmoid_window_check: %s", cudaGetErrorString(e));
    pose::sigmoid_window_check_kernel<<<sm_count() * 8, 256, 0, st>>>(violations_out, sigmoid_ref);
    return check_launch("sigmoid_window_check");
}

int pose_oks_matrix(const double* det_kp, const double* gt_kp, const double* gt_bbox, const double* gt_area,
                    const int* det_off, const int* gt_off, const long long* pair_off, const double* sigmas,
                    double* oks_out, double* det_area_out, int Q, int D, int G, long long n_pairs, int K, pose_stream_t stream) {
    if (Q < 0 || D < 0 || G < 0 || n_pairs < 0 || K <= 0 || K > pose::kOksMaxK)
        return fail(POSE_EINVAL, "oks_matrix: bad size (Q=%d D=%d G=%d pairs=%lld K=%d, K <= %d)", Q, D, G, n_pairs, K, pose::kOksMaxK);
    if (!det_off || !gt_off || !pair_off || !sigmas || (D > 0 && (!det_kp || !det_area_out)) || (G > 0 && (!gt_kp || !gt_bbox || !gt_area)) ||
        (n_pairs > 0 && !oks_out))
        return fail(POSE_EINVAL, "oks_matrix: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (D > 0) {
        pose::oks_det_area_kernel<<<(D + 127) / 128, 128, 0, st>>>(det_kp, det_area_out, D, K);
        if (int rc = check_launch("oks_det_area")) return rc;
    }
    if (n_pairs > 0) {
        long long blocks = (n_pairs + 127) / 128, cap = (long long)sm_count() * 16;
        pose::oks_matrix_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, st>>>(det_kp, gt_kp, gt_bbox, gt_area, det_off, gt_off,
                                                                                     pair_off, sigmas, oks_out, Q, K);
        if (int rc = check_launch("oks_matrix")) return rc;
    }
    return POSE_OK;
}

unsigned long long pose_oks_match_workspace_bytes(int A, int T, int G) {
    return (unsigned long long)(A > 0 ? A : 0) * (unsigned long long)(T > 0 ? T : 0) * (unsigned long long)(G > 0 ? G : 0);
}

int pose_oks_match(const double* oks, const long long* pair_off, const int* det_off, const int* gt_off, const double* det_area,
                   const double* gt_area, const unsigned char* gt_flags, const double* area_rng, const double* iou_thrs,
                   int Q, int A, int T, int D, int G, int* dt_match, unsigned char* dt_ignore, unsigned char* gt_ignore_out,
                   void* workspace, unsigned long long workspace_bytes, pose_stream_t stream) {
    if (Q < 0 || A <= 0 || T <= 0 || D < 0 || G < 0) return fail(POSE_EINVAL, "oks_match: bad size");
    if (!pair_off || !det_off || !gt_off || !area_rng || !iou_thrs || (D > 0 && (!det_area || !dt_match || !dt_ignore)) ||
        (G > 0 && (!gt_area || !gt_flags || !gt_ignore_out)))
        return fail(POSE_EINVAL, "oks_match: NULL pointer");
    if (workspace_bytes < pose_oks_match_workspace_bytes(A, T, G) || (G > 0 && !workspace))
        return fail(POSE_EWORKSPACE, "oks_match: workspace of %llu bytes needed", pose_oks_match_workspace_bytes(A, T, G));
    if (Q == 0) return POSE_OK;
    long long total = (long long)Q * A * T, blocks = (total + 127) / 128, cap = (long long)sm_count() * 16;
    pose::oks_match_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 128, 0, (cudaStream_t)stream>>>(
        oks, pair_off, det_off, gt_off, det_area, gt_area, gt_flags, area_rng, iou_thrs, Q, A, T, D, G, dt_match, dt_ignore,
        gt_ignore_out, (unsigned char*)workspace);
    return check_launch("oks_match");
}

unsigned long long pose_ap_accumulate_workspace_bytes(int A, int T, int D) {
    return 2ull * sizeof(double) * (unsigned long long)(A > 0 ? A : 0) * (unsigned long long)(T > 0 ? T : 0) * (unsigned long long)(D > 0 ? D : 0);
}

int pose_ap_accumulate(const long long* order, const int* dt_match, const unsigned char* dt_ignore, const unsigned char* gt_ignore,
                       const int* cat_det_off, const int* cat_gt_off, const double* rec_thrs, int C, int A, int T, int R, int D, int G,
                       double* precision, double* recall, void* workspace, unsigned long long workspace_bytes, pose_stream_t stream) {
    if (C <= 0 || A <= 0 || T <= 0 || R <= 0 || D < 0 || G < 0) return fail(POSE_EINVAL, "ap_accumulate: bad size");
    if (!cat_det_off || !cat_gt_off || !rec_thrs || !precision || !recall || (D > 0 && (!order || !dt_match || !dt_ignore)) || (G > 0 && !gt_ignore))
        return fail(POSE_EINVAL, "ap_accumulate: NULL pointer");
    if (workspace_bytes < pose_ap_accumulate_workspace_bytes(A, T, D) || (D > 0 && !workspace))
        return fail(POSE_EWORKSPACE, "ap_accumulate: workspace of %llu bytes needed", pose_ap_accumulate_workspace_bytes(A, T, D));
    if (reinterpret_cast<uintptr_t>(workspace) & 7u) return fail(POSE_EALIGN, "ap_accumulate: workspace must be 8-byte aligned");
    double* pr = (double*)workspace;
    double* rc = pr + (size_t)A * T * D;
    pose::ap_accumulate_kernel<<<C * A * T, pose::kApThreads, 0, (cudaStream_t)stream>>>(order, dt_match, dt_ignore, gt_ignore, cat_det_off,
                                                                                      cat_gt_off, rec_thrs, C, A, T, R, D, G, precision,
                                                                                      recall, pr, rc);
    return check_launch("ap_accumulate");
}

}  // extern "C"
