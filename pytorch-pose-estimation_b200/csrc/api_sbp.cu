// extern "C" boundary of libpose_b200.so: Simple-Baselines entry points, the multi-GPU exchange, loss reduce, diagnostics.
#include "host_common.h"
#include "sbp_kernels.cuh"
#include "exchange_kernels.cuh"

using namespace pose_host;

namespace {
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

#ifndef POSE_TMA_FUSED_PDL
#define POSE_TMA_FUSED_PDL 1       // fused bulk-staged kernel launched with programmatic stream serialisation (A/B: tools/tune_fused.py)
#endif
namespace {
template <int V, int TGT, bool GRAD, bool WTGT, bool DEC>
int launch_fused(const pose::SbpFusedParams& P0, size_t smem, cudaStream_t st) {
    pose::SbpFusedParams P = P0;
    // one CTA per heat map (first use of a configuration also opts in to the shared memory it needs)
    if (resident_ctas(pose::sbp_fused_kernel<V, TGT, GRAD, WTGT, DEC>, pose::kSbpThreads, smem, "sbp_fused") == 0) return last_code();
    pose::sbp_fused_kernel<V, TGT, GRAD, WTGT, DEC><<<(unsigned)P.n_maps, pose::kSbpThreads, smem, st>>>(P);
    return check_launch("sbp_fused");
}

template <bool GRAD, bool DEC>
int launch_fused_tma(const pose::SbpFusedParams& P0, cudaStream_t st) {
    pose::SbpFusedParams P = P0;
    constexpr int cls = pose::tma_cls(GRAD), mpc = pose::tma_mpc(cls), threads = pose::tma_threads(cls);
    const size_t smem = pose::sbp_tma_smem_bytes(P.HW, pose::tma_cls(GRAD));
    if (resident_ctas(pose::sbp_fused_tma_kernel<GRAD, DEC>, threads, smem, "sbp_fused(tma)") == 0) return last_code();
    const long long ctas = (P.n_maps + mpc - 1) / mpc;
#if POSE_TMA_FUSED_PDL
    launch_pdl(pose::sbp_fused_tma_kernel<GRAD, DEC>, (unsigned)ctas, (unsigned)threads, smem, st, P);
#else
    pose::sbp_fused_tma_kernel<GRAD, DEC><<<(unsigned)ctas, threads, smem, st>>>(P);
#endif
    return check_launch("sbp_fused_tma");
}

template <int V, int TGT>
int dispatch_fused(const pose::SbpFusedParams& P, unsigned flags, size_t smem, cudaStream_t st) {
    const bool g = flags & POSE_F_GRAD, t = (flags & POSE_F_TARGET_OUT) && TGT == pose::TGT_RENDER, d = flags & POSE_F_DECODE;
#define POSE_CASE(G, T, D) \
    if (g == G && t == T && d == D) return launch_fused<V, TGT, G, T, D>(P, smem, st);
    POSE_CASE(false, false, false) POSE_CASE(true, false, false) POSE_CASE(false, false, true) POSE_CASE(true, false, true)
    if (TGT == pose::TGT_RENDER) {
        POSE_CASE(false, true, false) POSE_CASE(true, true, false) POSE_CASE(false, true, true) POSE_CASE(true, true, true)
    }
#undef POSE_CASE
    return fail(POSE_EINVAL, "sbp_fused: unsupported flag combination 0x%x", flags);
}
}  // namespace

extern "C" {

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
    cudaStream_t st = (cudaStream_t)stream;
    // one CTA per heat map
    if (vec) {
        if (resident_ctas(pose::sbp_render_kernel<4>, pose::kSbpThreads, smem, "sbp_render") == 0) return last_code();
        pose::sbp_render_kernel<4><<<(unsigned)P.n_maps, pose::kSbpThreads, smem, st>>>(P);
    } else {
        if (resident_ctas(pose::sbp_render_kernel<1>, pose::kSbpThreads, smem, "sbp_render") == 0) return last_code();
        pose::sbp_render_kernel<1><<<(unsigned)P.n_maps, pose::kSbpThreads, smem, st>>>(P);
    }
    return check_launch("sbp_render");
}

// workspace: [N*K][2] fp64 loss pairs of the maps | [R][2] fp64 slice sums | the reduction's ticket counter
unsigned long long pose_sbp_fused_workspace_bytes(int N, int K) {
    const long long maps = (long long)(N > 0 ? N : 0) * (long long)(K > 0 ? K : 0);
    return (unsigned long long)(maps + pose::reduce_slices(maps)) * 2ull * sizeof(double) + 16ull;
}


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
    if (!workspace || workspace_bytes < pose_sbp_fused_workspace_bytes(N, K)) return fail(POSE_EWORKSPACE, "sbp_fused: workspace too small (%llu bytes needed)", pose_sbp_fused_workspace_bytes(N, K));
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

    const long long n_maps = (long long)N * K;
    if (n_maps > 0x7fffffffll) return fail(POSE_EINVAL, "sbp_fused: N*K=%lld maps exceed one grid", n_maps);
    const int R = pose::reduce_slices(n_maps);
    double* slices = reinterpret_cast<double*>(workspace) + 2 * n_maps;
    P.ticket = reinterpret_cast<unsigned int*>(slices + 2 * R);
    if (N > 0) {
        // rendered target, read-only variants: a 128-bit vector must lie in one row (the padded-template lookup of render_loss_vec)
        const bool vec = (P.HW % 4 == 0) && W >= 4 && (!kp || (flags & POSE_F_GRAD) || W % 4 == 0) && aligned16(logits) && (!target_in || aligned16(target_in)) &&
                         (!(flags & POSE_F_GRAD) || aligned16(dlogits)) && (!(flags & POSE_F_TARGET_OUT) || aligned16(target_out));
        const size_t smem = 0;
        int rc;
        // bulk-async staging (the default of the Python layer): render mode, aligned maps that fit in shared memory
        const bool tma = (flags & POSE_F_TMA) && kp && vec && !(flags & POSE_F_TARGET_OUT) && pose::sbp_tma_smem_bytes(P.HW, pose::tma_cls((flags & POSE_F_GRAD) != 0)) <= 112 * 1024;
        if (tma) {
            const bool g = flags & POSE_F_GRAD, d = flags & POSE_F_DECODE;
            rc = g ? (d ? launch_fused_tma<true, true>(P, st) : launch_fused_tma<true, false>(P, st))
                   : (d ? launch_fused_tma<false, true>(P, st) : launch_fused_tma<false, false>(P, st));
        } else if (kp) rc = vec ? dispatch_fused<4, pose::TGT_RENDER>(P, flags, smem, st) : dispatch_fused<1, pose::TGT_RENDER>(P, flags, smem, st);
        else rc = vec ? dispatch_fused<4, pose::TGT_DENSE>(P, flags, smem, st) : dispatch_fused<1, pose::TGT_DENSE>(P, flags, smem, st);
        if (rc) return rc;
    }
    pose::SbpEpilogueParams E;
    E.partials = P.partials; E.n_pairs = n_maps; E.slices = slices; E.ticket = P.ticket; E.R = R;
    E.w0 = (double)lambda_pos; E.w1 = (double)lambda_neg; E.inv_norm = inv_norm;
    E.loss_out = loss_out; E.num_out = loss_num_out;
    E.joints = joints; E.bbox = bbox; E.packed = packed_out; E.N = bbox ? N : 0; E.K = K; E.in_h = (double)input_h; E.in_w = (double)input_w;
    const unsigned bp_ctas = bbox ? (unsigned)(((long long)N * 32 + 255) / 256) : 0u;
    E.bp_ctas = (int)bp_ctas;
    if (exchange) {
        launch_pdl(pose::sbp_epilogue_p2p_kernel, bp_ctas + (unsigned)R, 256u, 0, st, E, to_dev(*exchange, (double)lambda_pos, (double)lambda_neg, inv_norm));
        return check_launch("sbp_epilogue_p2p");
    }
    launch_pdl(pose::sbp_epilogue_kernel, bp_ctas + (unsigned)R, 256u, 0, st, E);
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
    launch_pdl(pose::exchange_wait_reduce_kernel, 1u, 256u, 0, (cudaStream_t)stream, to_dev(*x), mode, w0, w1, inv_norm, loss_out,
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
    if (!pairs || n < 0 || stride < 2 || (!loss_out && !num_out)) return fail(POSE_EINVAL, "loss_reduce: bad argument");
    launch_pdl(pose::loss_reduce_kernel, 1u, 256u, 0, (cudaStream_t)stream, pairs, n, stride, w0, w1, inv_norm, loss_out, num_out);
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
    P.n_maps = (long long)N * K; P.H = H; P.W = W; P.HW = H * W; P.divW = make_div(W); P.refine = refine & POSE_DEC_REFINE; P.sig_ref = sigmoid_ref;
    const bool vec = (P.HW % 4 == 0) && aligned16(x);
    if (P.n_maps > 0x7fffffffll) return fail(POSE_EINVAL, "sbp_decode: N*K=%lld maps exceed one grid", P.n_maps);
    cudaStream_t st = (cudaStream_t)stream;
    const bool sig = apply_sigmoid != 0;
    const size_t tsmem = pose::sbp_tma_smem_bytes(P.HW, pose::kTmaDec);
    constexpr int tthreads = pose::tma_threads(pose::kTmaDec), tmpc = pose::tma_mpc(pose::kTmaDec);
    if (vec && !(refine & POSE_DEC_NO_TMA) && tsmem <= 112 * 1024) {
        // bulk-async staging: several maps per CTA
        const long long ctas = (P.n_maps + tmpc - 1) / tmpc;
        if (sig) {
            if (resident_ctas(pose::sbp_decode_tma_kernel<true>, tthreads, tsmem, "sbp_decode(tma)") == 0) return last_code();
            pose::sbp_decode_tma_kernel<true><<<(unsigned)ctas, tthreads, tsmem, st>>>(P);
        } else {
            if (resident_ctas(pose::sbp_decode_tma_kernel<false>, tthreads, tsmem, "sbp_decode(tma)") == 0) return last_code();
            pose::sbp_decode_tma_kernel<false><<<(unsigned)ctas, tthreads, tsmem, st>>>(P);
        }
        return check_launch("sbp_decode_tma");
    }
#define POSE_DEC(V, S) pose::sbp_decode_kernel<V, S><<<(unsigned)P.n_maps, pose::kSbpThreads, 0, st>>>(P);      /* one CTA per heat map */
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

}  // extern "C"
