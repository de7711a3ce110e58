// extern "C" boundary of libpose_b200.so: head fusion (1x1 convolution on tcgen05 + the heat-map loss / gradient / decode epilogue).
#include <cuda.h>

#include <mutex>

#define POSE_GLOBAL static __global__      /* sbp_kernels.cuh is also compiled into api_sbp.cu */

#include "host_common.h"
#include "head_kernels.cuh"

using namespace pose_host;

namespace {

// cuTensorMapEncodeTiled comes from the driver (libcuda); the library links only the static runtime, so the entry point is
// resolved through the runtime once (no link-time dependency: the .so still loads on a machine without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
    });
    return fn;
}

template <bool RT>
int launch_head(const CUtensorMap& tmX, const CUtensorMap& tmW, pose::SbpHeadParams P, int tuning_raw, int tuning_lo, cudaStream_t st) {
    auto kern = pose::sbp_head_fused_kernel<RT>;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return fail((int)e, "sbp_head_fused: %s", cudaGetErrorString(e));
    const long long w_bytes = (long long)P.n_kc * pose::kHeadWChunkBytes;
    const long long avail = (long long)pose::kHeadSmemCap - (long long)fa.sharedSizeBytes - 1024 - w_bytes;
    int stages = (int)(avail / pose::kHeadStageBytes);
    if (stages > pose::kHeadMaxStages + 2) stages = pose::kHeadMaxStages + 2;
    const bool residual = !(P.flags & pose::kHeadNoResidual);
    int lo = (residual && !RT) ? 2 : 0, raw = stages - lo;      // RT: the residuals live in tensor memory, every stage holds features
    if (tuning_raw > 0) raw = tuning_raw;
    if (tuning_lo > 0 && residual && !RT) lo = tuning_lo;
    if (raw > pose::kHeadMaxStages) raw = pose::kHeadMaxStages;
    if (raw < 2 || (residual && !RT && lo < 1) || raw + lo > stages)
        return fail(POSE_EINVAL, "sbp_head_fused: C=%d does not leave room for the tile ring (%lld B of shared memory for %d+%d stages)", P.C, avail, raw, lo);
    P.raw_stages = raw; P.lo_stages = lo > 0 ? lo : 1;
    const size_t smem = (size_t)(1024 + w_bytes + (long long)(raw + lo) * pose::kHeadStageBytes);
    if (resident_ctas(kern, pose::kHeadThreads, smem, "sbp_head_fused") == 0) return last_code();
    const int grid = P.N < sm_count() ? P.N : sm_count();
    kern<<<grid, pose::kHeadThreads, smem, st>>>(tmX, tmW, P);
    return check_launch("sbp_head_fused");
}

}  // namespace

extern "C" {

unsigned long long pose_sbp_head_workspace_bytes(int N, int K, int C) {
    const unsigned long long wcat = (unsigned long long)pose::kHeadN * (unsigned long long)(C > 0 ? C : 0) * sizeof(float);
    return (wcat + 255ull) / 256ull * 256ull + pose_sbp_fused_workspace_bytes(N, K);
}

int pose_sbp_head_fused(const float* features, const float* weight, const void* kp, int kp_dtype, double sigma,
                        const float* lut, int lut_n, float* dlogits, float* logits_out, float* loss_out, double* loss_num_out,
                        float* joints, float conf_threshold, float coord_scale, int N, int C, int K, int H, int W,
                        float lambda_pos, float lambda_neg, double inv_norm, unsigned flags,
                        const double* bbox, float* packed_out, int input_h, int input_w, int tuning,
                        void* workspace, unsigned long long workspace_bytes, pose_stream_t stream) {
    if (N < 0 || K <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(POSE_EINVAL, "sbp_head_fused: bad shape N=%d C=%d K=%d H=%d W=%d", N, C, K, H, W);
    if (K > pose::kHeadMaxK) return fail(POSE_EINVAL, "sbp_head_fused: K=%d joints (at most %d)", K, pose::kHeadMaxK);
    if (C % pose::kHeadKC) return fail(POSE_EINVAL, "sbp_head_fused: C=%d must be a multiple of %d", C, pose::kHeadKC);
    const long long HW = (long long)H * W;
    if (HW % pose::kHeadM || HW >= (1ll << 20) || W >= (1 << 11)) return fail(POSE_EINVAL, "sbp_head_fused: H*W=%lld must be a multiple of %d (and < 2^20)", HW, pose::kHeadM);
    if (!features || !weight) return fail(POSE_EINVAL, "sbp_head_fused: NULL input");
    // kp == NULL: decode-only call (inference: head -> DecodeSBP, no target): the loss outputs then hold the loss against an all-zero target
    if (!kp && ((flags & POSE_F_GRAD) || !(flags & POSE_F_DECODE))) return fail(POSE_EINVAL, "sbp_head_fused: without keypoints only POSE_F_DECODE (no POSE_F_GRAD) makes sense");
    if (kp && (!lut || lut_n <= 0 || lut_n > pose::kHeadMaxLut || !(sigma > 0.0))) return fail(POSE_EINVAL, "sbp_head_fused: template side %d (at most %d), sigma must be > 0", lut_n, pose::kHeadMaxLut);
    if (!kp) { lut_n = 0; sigma = 1.0; }
    if ((flags & POSE_F_GRAD) && !dlogits) return fail(POSE_EINVAL, "sbp_head_fused: POSE_F_GRAD without dlogits");
    if ((flags & POSE_F_DECODE) && !joints) return fail(POSE_EINVAL, "sbp_head_fused: POSE_F_DECODE without joints");
    if ((flags & POSE_F_HEAD_LOGITS_OUT) && !logits_out) return fail(POSE_EINVAL, "sbp_head_fused: POSE_F_HEAD_LOGITS_OUT without logits_out");
    if (!loss_out && !loss_num_out) return fail(POSE_EINVAL, "sbp_head_fused: no loss output");
    if ((packed_out != nullptr) != (bbox != nullptr)) return fail(POSE_EINVAL, "sbp_head_fused: bbox and packed_out go together");
    if (bbox && (!(flags & POSE_F_DECODE) || input_h <= 0 || input_w <= 0)) return fail(POSE_EINVAL, "sbp_head_fused: back-projection needs POSE_F_DECODE and the input size");
    if (!workspace || workspace_bytes < pose_sbp_head_workspace_bytes(N, K, C)) return fail(POSE_EWORKSPACE, "sbp_head_fused: workspace too small (%llu bytes needed)", pose_sbp_head_workspace_bytes(N, K, C));
    if (!aligned16(features) || !aligned16(workspace) || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return fail(POSE_EALIGN, "sbp_head_fused: features must be 16-byte, workspace 256-byte aligned");
    if ((long long)N * K > 0x7fffffffll) return fail(POSE_EINVAL, "sbp_head_fused: too many maps");
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return fail(POSE_EINVAL, "sbp_head_fused: cuTensorMapEncodeTiled is not available from this driver");
    cudaStream_t st = (cudaStream_t)stream;

    float* wcat = reinterpret_cast<float*>(workspace);
    const unsigned long long wcat_bytes = ((unsigned long long)pose::kHeadN * C * sizeof(float) + 255ull) / 256ull * 256ull;
    double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + wcat_bytes);
    const long long n_maps = (long long)N * K;
    const int R = pose::reduce_slices(n_maps);
    double* slices = partials + 2 * n_maps;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(slices + 2 * R);

    if (N > 0) {
        pose::head_split_weights_kernel<<<(pose::kHeadN * C + 255) / 256, 256, 0, st>>>(weight, wcat, K, C);
        if (int rc = check_launch("head_split_weights")) return rc;

        CUtensorMap tmX, tmW;
        {   // features [N][C][HW] fp32 viewed as {32 pixels, C, HW/32 pixel blocks, N}: a box {32, 32, 4, 1} lands in shared memory as
            // [4 pixel blocks][32 channels][32 pixels] with the 128-byte / 32-byte-atom swizzle = the canonical MN-major SW128_32B UMMA operand (the one layout tf32 accepts transposed)
            const cuuint64_t dims[4] = {32, (cuuint64_t)C, (cuuint64_t)(HW / 32), (cuuint64_t)N};
            const cuuint64_t strides[3] = {(cuuint64_t)HW * 4, 128, (cuuint64_t)C * (cuuint64_t)HW * 4};
            const cuuint32_t box[4] = {32, (cuuint32_t)pose::kHeadKC, (cuuint32_t)(pose::kHeadM / 32), 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(features), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(POSE_EINVAL, "sbp_head_fused: cuTensorMapEncodeTiled(features) failed (%d)", (int)r);
        }
        {   // Wcat [48][C]: box {32 channels, 48 rows} = one K-major SW128 chunk
            const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)pose::kHeadN};
            const cuuint64_t strides[1] = {(cuuint64_t)C * 4};
            const cuuint32_t box[2] = {(cuuint32_t)pose::kHeadKC, (cuuint32_t)pose::kHeadN};
            const cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, wcat, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(POSE_EINVAL, "sbp_head_fused: cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
        }

        pose::SbpHeadParams P;
        memset(&P, 0, sizeof(P));
        P.N = N; P.K = K; P.C = C; P.H = H; P.W = W; P.HW = (int)HW;
        P.tiles_per_img = (int)(HW / pose::kHeadM); P.n_kc = C / pose::kHeadKC;
        P.divW = make_div(W);
        P.kp = kp; P.kp_f64 = kp_dtype == POSE_KP_F64; P.lut = lut; P.lut_n = lut_n; P.three_sigma = 3 * sigma;
        P.dlogits = dlogits; P.logits_out = logits_out; P.joints = joints; P.partials = partials; P.ticket = ticket;
        P.thr = conf_threshold; P.scale = coord_scale;
        P.gpos = (float)(2.0 * (double)lambda_pos * inv_norm);
        P.gneg = (float)(2.0 * (double)lambda_neg * inv_norm);
        P.sig_ref = (flags & POSE_F_SIGMOID_CUDA) ? POSE_SIGMOID_ATEN_CUDA : POSE_SIGMOID_ATEN_CPU;
        P.flags = ((flags & POSE_F_GRAD) ? pose::kHeadGrad : 0u) | ((flags & POSE_F_DECODE) ? pose::kHeadDecode : 0u) |
                  ((flags & POSE_F_HEAD_LOGITS_OUT) ? pose::kHeadLogitsOut : 0u) | ((flags & POSE_F_HEAD_NO_RESIDUAL) ? pose::kHeadNoResidual : 0u);
        // tuning (0 = defaults): bits 0-7 feature stages, 8-15 residual stages (shared-memory variant), bit 24 = residuals through
        // shared memory instead of tensor memory
        const int t_raw = tuning & 0xff, t_lo = (tuning >> 8) & 0xff;
        const int rc = (tuning & (1 << 24)) ? launch_head<false>(tmX, tmW, P, t_raw, t_lo, st) : launch_head<true>(tmX, tmW, P, t_raw, t_lo, st);
        if (rc) return rc;
    }
    pose::SbpEpilogueParams E;
    E.partials = partials; E.n_pairs = n_maps; E.slices = slices; E.ticket = ticket; E.R = R;
    E.w0 = (double)lambda_pos; E.w1 = (double)lambda_neg; E.inv_norm = inv_norm;
    E.loss_out = loss_out; E.num_out = loss_num_out;
    E.joints = joints; E.bbox = bbox; E.packed = packed_out; E.N = bbox ? N : 0; E.K = K; E.in_h = (double)input_h; E.in_w = (double)input_w;
    const unsigned bp_ctas = bbox ? (unsigned)(((long long)N * 32 + 255) / 256) : 0u;
    E.bp_ctas = (int)bp_ctas;
    launch_pdl(pose::sbp_epilogue_kernel, bp_ctas + (unsigned)R, 256u, 0, st, E);
    return check_launch("sbp_epilogue");
}

}  // extern "C"
