// extern "C" boundary of libpose_b200.so: SPM entry points (render, loss, fused render+loss, decode, gather, rescale).
#include "host_common.h"
#include "spm_kernels.cuh"

using namespace pose_host;

namespace {
// workspace of the fused / single-pass render path:
//   [units][2] fp64 loss pairs | [R][2] slice sums | ticket (16 B) | [2R+1] fp64 quotient table | per-image geometry records
struct SpmWs {
    double* partials; double* slices; unsigned int* ticket; double* div_tab; unsigned char* geom;
    long long units; int slices_n; unsigned long long bytes;
};
SpmWs spm_ws_layout(void* base, int N, int K, int R) {
    SpmWs w;
    w.units = pose::spm_units(N, K, R);
    w.slices_n = pose::reduce_slices(w.units);
    unsigned long long off = 0;
    const uintptr_t b = reinterpret_cast<uintptr_t>(base);          // (base may be NULL: size query)
    w.partials = reinterpret_cast<double*>(b + off); off += (unsigned long long)w.units * 16;
    w.slices = reinterpret_cast<double*>(b + off); off += (unsigned long long)w.slices_n * 16;
    w.ticket = reinterpret_cast<unsigned int*>(b + off); off += 16;
    w.div_tab = reinterpret_cast<double*>(b + off); off += (unsigned long long)(2 * R + 1) * 8;
    off = (off + 255) / 256 * 256;
    w.geom = reinterpret_cast<unsigned char*>(b + off); off += (unsigned long long)(N > 0 ? N : 0) * pose::spm_geom_layout(R).stride;
    w.bytes = off;
    return w;
}

void fill_fused_params(pose::SpmFusedParams& P, const SpmWs& w, int N, int Pmax, int K, int R, double sigma, const float* lut, int lut_n) {
    P.lut = lut; P.lut_n = lut_n;
    P.three_sigma = 3 * sigma; P.half = (int)((6 * sigma + 2) / 2);
    P.z = std::sqrt((double)((long long)R * R + (long long)R * R));
    P.partials = w.partials; P.ticket = w.ticket; P.geom = w.geom; P.div_tab = w.div_tab; P.gl = pose::spm_geom_layout(R);
    P.N = N; P.Pmax = Pmax; P.K = K; P.R = R; P.quads = R * R / 4; P.div_qpr = make_div(R / 4);
    P.wpr = (R / 4 + 31) / 32;
    P.div_n = R <= 1024 ? 2 * R + 1 : 0;
}

// geometry pre-pass + the unit kernel (programmatic dependent launch: the unit CTAs start their bulk loads while the pre-pass drains)
template <bool LOSS, bool GRAD, bool WTGT>
int launch_spm_units(const pose::SpmFusedParams& P, const SpmWs& w, cudaStream_t st, const char* what) {
    const size_t gsmem = pose::spm_geom_smem_bytes(P.R, P.lut_n);
    if (resident_ctas(pose::spm_geometry_kernel, pose::kSpmGeomThreads, gsmem, what) == 0) return last_code();
    pose::spm_geometry_kernel<<<P.N, pose::kSpmGeomThreads, gsmem, st>>>(P);
    if (int rc = check_launch("spm_geometry")) return rc;
    const size_t smem = pose::spm_unit_smem_bytes(LOSS);
    if (resident_ctas(pose::spm_unit_kernel<LOSS, GRAD, WTGT>, pose::kSpmUnitThreads, smem, what) == 0) return last_code();
    if (P.N > 65535 || 1 + 2 * P.K > 65535) return fail(POSE_EINVAL, "%s: N=%d / K=%d exceed the grid (65535 images per call)", what, P.N, P.K);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)pose::spm_units_per_plane(P.R, LOSS), (unsigned)(1 + 2 * P.K), (unsigned)P.N);
        cfg.blockDim = dim3((unsigned)pose::kSpmUnitThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, pose::spm_unit_kernel<LOSS, GRAD, WTGT>, P);
    }
    return check_launch(what);
}
}  // namespace

extern "C" {

int pose_spm_render(const long long* centers, const long long* joints, const int* counts, float* target, int N, int Pmax,
                    int K, int R, double sigma, const float* lut, int lut_n, void* workspace, unsigned long long workspace_bytes,
                    pose_stream_t stream) {
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
    if (Pmax <= pose::kSpmFusedMaxPersons && !two_pass && workspace && workspace_bytes >= pose_spm_fused_workspace_bytes(N, K, R) &&
        aligned16(workspace)) {
        // single pass: the render-only form of the fused path (geometry pre-pass + one CTA per 16 KB unit, a pure write stream)
        const SpmWs w = spm_ws_layout(workspace, N, K, R);
        pose::SpmFusedParams F;
        memset(&F, 0, sizeof(F));
        F.target_out = target; F.centers = centers; F.joints = joints; F.counts = counts;
        fill_fused_params(F, w, N, Pmax, K, R, sigma, lut, lut_n);
        return launch_spm_units<false, false, true>(F, w, (cudaStream_t)stream, "spm_render(single pass)");
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
        launch_pdl(pose::spm_patch_kernel, (unsigned)((long long)N * Pmax), (unsigned)pose::kSpmThreads, (size_t)div_n * sizeof(double),
                   (cudaStream_t)stream, P, div_n);
    }
    return check_launch("spm_render");
}

// workspace of pose_spm_loss: one fp64 pair per 16 KB unit, the slice sums and the ticket of the two-level reduction, the root-mask bits
namespace {
struct SpmLossWs {
    double* partials; double* slices; unsigned int* ticket; unsigned int* mask;
    long long units; int slices_n, mask_words; unsigned long long bytes;
};
SpmLossWs spm_loss_ws_layout(void* base, int N, int K, int R) {
    SpmLossWs w;
    const long long n = N > 0 ? N : 0;
    w.units = n * (1 + 2 * K) * pose::spm_loss_units_per_plane(R);
    w.slices_n = pose::reduce_slices(w.units);
    w.mask_words = pose::spm_mask_words(R);
    unsigned long long off = 0;
    const uintptr_t b = reinterpret_cast<uintptr_t>(base);          // (base may be NULL: size query)
    w.partials = reinterpret_cast<double*>(b + off); off += (unsigned long long)w.units * 16;
    w.slices = reinterpret_cast<double*>(b + off); off += (unsigned long long)w.slices_n * 16;
    w.ticket = reinterpret_cast<unsigned int*>(b + off); off += 16;
    w.mask = reinterpret_cast<unsigned int*>(b + off); off += (unsigned long long)n * w.mask_words * 4;
    w.bytes = off > 16 ? off : 16;
    return w;
}
}  // namespace

unsigned long long pose_spm_loss_workspace_bytes(int N, int K, int R) {
    if (N < 0 || K <= 0 || R <= 0) return 0ull;
    return spm_loss_ws_layout(nullptr, N, K, R).bytes;
}

int pose_spm_loss(const float* logits, const float* target, float* dlogits, float* loss_out, double* loss_num_out, int N,
                  int K, int R, float lambda_root, float lambda_disp, double inv_norm, int write_grad, void* workspace,
                  unsigned long long workspace_bytes, pose_stream_t stream) {
    if (N < 0 || K <= 0 || R <= 0 || R % 4 != 0) return fail(POSE_EINVAL, "spm_loss: bad shape (R must be a multiple of 4)");
    if (!loss_out && !loss_num_out) return fail(POSE_EINVAL, "spm_loss: no loss output");
    if (!workspace || workspace_bytes < pose_spm_loss_workspace_bytes(N, K, R) || !aligned16(workspace))
        return fail(POSE_EWORKSPACE, "spm_loss: workspace too small / unaligned (%llu bytes needed)", pose_spm_loss_workspace_bytes(N, K, R));
    if (N == 0)    // an empty batch (its tensors may have no storage at all) reduces zero pairs: loss 0
        return pose_loss_reduce((const double*)workspace, 0, 2ll, (double)lambda_root, (double)lambda_disp, inv_norm, loss_out, loss_num_out, stream);
    if (!logits || !target || (write_grad && !dlogits)) return fail(POSE_EINVAL, "spm_loss: NULL pointer");
    if (!aligned16(logits) || !aligned16(target) || (write_grad && !aligned16(dlogits))) return fail(POSE_EALIGN, "spm_loss: tensors must be 16-byte aligned");
    if (write_grad && dlogits == logits) return fail(POSE_EINVAL, "spm_loss: dlogits must not alias logits");
    if (N > 65535 || 1 + 2 * K > 65535) return fail(POSE_EINVAL, "spm_loss: N=%d / K=%d exceed the grid (65535 images per call)", N, K);
    cudaStream_t st = (cudaStream_t)stream;
    const SpmLossWs w = spm_loss_ws_layout(workspace, N, K, R);
    pose::SpmLossParams P;
    P.logits = logits; P.target = target; P.dlogits = dlogits; P.partials = w.partials; P.mask = w.mask; P.ticket = w.ticket;
    P.quads = R * R / 4; P.mask_words = w.mask_words; P.C = 1 + 2 * K;
    P.groot = (float)(2.0 * (double)lambda_root * inv_norm);
    P.gdisp = (float)((double)lambda_disp * inv_norm);
    // (1) root-mask bits of every image; (2) one CTA per 16 KB unit, launched with programmatic stream serialisation: its loads
    // are in flight while (1) drains; (3) the two-level fixed-order reduction of the per-unit pairs
    pose::spm_root_mask_kernel<<<dim3((unsigned)((P.quads + 255) / 256), (unsigned)N), 256, 0, st>>>(P);
    if (int rc = check_launch("spm_root_mask")) return rc;
    const unsigned upp = (unsigned)pose::spm_loss_units_per_plane(R);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(upp, (unsigned)P.C, (unsigned)N);
        cfg.blockDim = dim3((unsigned)pose::kSpmThreads);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (write_grad) cudaLaunchKernelEx(&cfg, pose::spm_loss_kernel<true>, P);
        else cudaLaunchKernelEx(&cfg, pose::spm_loss_kernel<false>, P);
    }
    if (int rc = check_launch("spm_loss")) return rc;
    launch_pdl(pose::spm_loss_reduce_kernel, (unsigned)w.slices_n, 256u, 0, st, (const double*)w.partials, w.units, w.slices, w.ticket, w.slices_n,
               (double)lambda_root, (double)lambda_disp, inv_norm, loss_out, loss_num_out);
    return check_launch("spm_loss_reduce");
}

unsigned long long pose_spm_fused_workspace_bytes(int N, int K, int R) {
    if (N < 0 || K <= 0 || R <= 0) return 0ull;
    return spm_ws_layout(nullptr, N, K, R).bytes;
}

int pose_spm_fused(const float* logits, const long long* centers, const long long* joints, const int* counts, float* dlogits,
                   float* target_out, float* loss_out, double* loss_num_out, int N, int Pmax, int K, int R, double sigma,
                   const float* lut, int lut_n, float lambda_root, float lambda_disp, double inv_norm, unsigned flags,
                   void* workspace, unsigned long long workspace_bytes, pose_stream_t stream) {
    if (N < 0 || Pmax < 0 || K <= 0 || R <= 0 || R % 4 != 0 || R > 2048) return fail(POSE_EINVAL, "spm_fused: bad shape (R must be a multiple of 4, <= 2048)");
    if (Pmax > pose::kSpmFusedMaxPersons) return fail(POSE_EINVAL, "spm_fused: Pmax=%d > %d persons per image (use pose_spm_render + pose_spm_loss)", Pmax, pose::kSpmFusedMaxPersons);
    if (flags & ~(POSE_F_GRAD | POSE_F_TARGET_OUT)) return fail(POSE_EINVAL, "spm_fused: unsupported flags 0x%x", flags);
    const bool grad = flags & POSE_F_GRAD, wtgt = flags & POSE_F_TARGET_OUT;
    if (!loss_out && !loss_num_out) return fail(POSE_EINVAL, "spm_fused: no loss output");
    if (!workspace || workspace_bytes < pose_spm_fused_workspace_bytes(N, K, R) || !aligned16(workspace))
        return fail(POSE_EWORKSPACE, "spm_fused: workspace too small / unaligned (%llu bytes needed)", pose_spm_fused_workspace_bytes(N, K, R));
    cudaStream_t st = (cudaStream_t)stream;
    const SpmWs w = spm_ws_layout(workspace, N, K, R);
    if (N > 0) {
        if (!logits || !counts || !lut || lut_n <= 0 || lut_n > 64 || !(sigma > 0.0) || (Pmax > 0 && (!centers || !joints)) ||
            (grad && !dlogits) || (wtgt && !target_out))
            return fail(POSE_EINVAL, "spm_fused: bad argument");
        if (!aligned16(logits) || (grad && !aligned16(dlogits)) || (wtgt && !aligned16(target_out))) return fail(POSE_EALIGN, "spm_fused: tensors must be 16-byte aligned");
        if ((grad && dlogits == logits) || (wtgt && target_out == logits)) return fail(POSE_EINVAL, "spm_fused: outputs must not alias logits");
        pose::SpmFusedParams P;
        memset(&P, 0, sizeof(P));
        P.logits = logits; P.dlogits = dlogits; P.target_out = target_out;
        P.centers = centers; P.joints = joints; P.counts = counts;
        fill_fused_params(P, w, N, Pmax, K, R, sigma, lut, lut_n);
        P.groot = (float)(2.0 * (double)lambda_root * inv_norm);
        P.gdisp = (float)((double)lambda_disp * inv_norm);
        int rc;
        if (grad && wtgt) rc = launch_spm_units<true, true, true>(P, w, st, "spm_fused");
        else if (grad) rc = launch_spm_units<true, true, false>(P, w, st, "spm_fused");
        else if (wtgt) rc = launch_spm_units<true, false, true>(P, w, st, "spm_fused");
        else rc = launch_spm_units<true, false, false>(P, w, st, "spm_fused");
        if (rc) return rc;
        launch_pdl(pose::spm_loss_reduce_kernel, (unsigned)w.slices_n, 256u, 0, st, (const double*)w.partials, w.units, w.slices, w.ticket, w.slices_n,
                   (double)lambda_root, (double)lambda_disp, inv_norm, loss_out, loss_num_out);
        return check_launch("spm_loss_reduce");
    }
    // an empty batch reduces zero pairs: loss 0
    return pose_loss_reduce((const double*)workspace, 0, 2ll, (double)lambda_root, (double)lambda_disp, inv_norm, loss_out, loss_num_out, stream);
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
    const int threads = N <= 2 * sm_count() ? 512 : (N <= 4 * sm_count() ? 256 : 128);
    {
        int ok;
        if (threads == 512) ok = resident_ctas(pose::spm_decode_kernel<512>, 512, smem, "spm_decode");
        else if (threads == 256) ok = resident_ctas(pose::spm_decode_kernel<256>, 256, smem, "spm_decode");
        else ok = resident_ctas(pose::spm_decode_kernel<128>, 128, smem, "spm_decode");
        if (ok == 0) return last_code();
    }
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
        // the joint test of get_spm_keypoints, sqrt_fp64(fp32 q) < dist_thr, as an fp32 bound (sqrt is monotone and correctly rounded)
        float ql = (float)(dist_threshold * dist_threshold);
        while (ql > 0.0f && !(std::sqrt((double)std::nextafterf(ql, -INFINITY)) < dist_threshold)) ql = std::nextafterf(ql, -INFINITY);
        while (std::sqrt((double)ql) < dist_threshold) ql = std::nextafterf(ql, INFINITY);
        P.q_lt = ql;
    }
    if (threads == 512) pose::spm_decode_kernel<512><<<N, 512, smem, (cudaStream_t)stream>>>(P);
    else if (threads == 256) pose::spm_decode_kernel<256><<<N, 256, smem, (cudaStream_t)stream>>>(P);
    else pose::spm_decode_kernel<128><<<N, 128, smem, (cudaStream_t)stream>>>(P);
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

int pose_spm_gather_chain(const float* roots, const float* disp, const int* parent, float* kps, int n_roots, int K, int R,
                          double dist_threshold, pose_stream_t stream) {
    if (n_roots < 0 || K <= 0 || R <= 0 || !disp || !parent || (n_roots > 0 && (!roots || !kps))) return fail(POSE_EINVAL, "spm_gather_chain: bad argument");
    if (n_roots == 0) return POSE_OK;
    const float zf = (float)std::sqrt((double)((long long)R * R + (long long)R * R));
    const int total = n_roots * K;
    pose::spm_gather_chain_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(roots, disp, parent, kps, n_roots, K, R, zf, dist_threshold);
    return check_launch("spm_gather_chain");
}

}  // extern "C"
