// extern "C" boundary of libpose_b200.so: SPM entry points (render, loss, fused render+loss, decode, gather, rescale).
#include "host_common.h"
#include "spm_kernels.cuh"

using namespace pose_host;

extern "C" {

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
            const int fchunk = pose::kSpmThreads * pose::spm_fused_u(false, true, false);
            const long long funits = (long long)N * (1 + 2 * K) * ((quads + fchunk - 1) / fchunk);
#define POSE_SPMR(RG, MP)                                                                                                      \
    {                                                                                                                          \
        const int fgrid = persistent_grid(pose::spm_fused_kernel<false, false, true, RG, MP>, pose::kSpmThreads, fsmem, funits, "spm_render"); \
        if (fgrid == 0) return last_code();                                                                                    \
        pose::spm_fused_kernel<false, false, true, RG, MP><<<fgrid, pose::kSpmThreads, fsmem, (cudaStream_t)stream>>>(F); \
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
        launch_pdl(pose::spm_patch_kernel, (unsigned)((long long)N * Pmax), (unsigned)pose::kSpmThreads, (size_t)div_n * sizeof(double),
                   (cudaStream_t)stream, P, div_n);
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
    // (an empty batch reduces zero pairs: loss 0)
    return pose_loss_reduce(P.partials, grid, 2ll, (double)lambda_root, (double)lambda_disp, inv_norm, loss_out, loss_num_out, stream);
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
        const int uchunk = pose::kSpmThreads * pose::spm_fused_u(grad, wtgt);
        const long long units = (long long)N * (1 + 2 * K) * ((P.quads + uchunk - 1) / uchunk);
#define POSE_SPMF3(G, T, RG, MP)                                                                                              \
    {                                                                                                                          \
        grid = persistent_grid(pose::spm_fused_kernel<true, G, T, RG, MP>, pose::kSpmThreads, smem, units, "spm_fused");  \
        if (grid == 0) return last_code();                                                                                     \
        pose::spm_fused_kernel<true, G, T, RG, MP><<<grid, pose::kSpmThreads, smem, st>>>(P);                             \
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
    return pose_loss_reduce((const double*)workspace, grid, 2ll, (double)lambda_root, (double)lambda_disp, inv_norm, loss_out, loss_num_out, stream);
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

}  // extern "C"
