// extern "C" boundary of libpose_b200.so: OKS / AP evaluation entry points.
#include "host_common.h"
#include "oks_kernels.cuh"

using namespace pose_host;

extern "C" {

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
