// OKS / AP kernels: the COCO keypoint evaluation behind SBPmAPCOCO.result (utils/sbp_utils.py:166-189),
// SPMmAPCOCO.result (utils/spm_utils.py:325-351) and SBPmAPPIS.result.  The reference delegates it to pycocotools
// (COCOeval "keypoints": computeOks -> evaluateImg -> accumulate); these kernels restate that published algorithm on
// the device in fp64 so that `result()` needs no third-party package and no per-image Python loops.
//
// Data model.  A "group" is one (category, image) pair, q = cat * n_images + img.  Detections are stored sorted by
// (group, -score) (stable) and cut to max_det per group; ground truths in annotation order.  det_off / gt_off [Q+1]
// are prefix offsets, pair_off [Q+1] (int64) the offsets of each group's D_q x G_q OKS block (row = detection).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pose {

constexpr int kOksMaxK = 32;        // keypoints per instance (COCO: 17)
constexpr int kApThreads = 256;

__device__ __forceinline__ int upper_group(const long long* __restrict__ off, int n, long long v) {
    // largest q in [0, n) with off[q] <= v   (off is non-decreasing, off[0] = 0, v < off[n])
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// numpy's pairwise summation for n <= 128 contiguous doubles (what np.sum does on the <= 17 exp(-e) terms):
// n < 8 sequential; otherwise 8 running sums over the multiple-of-8 prefix, a fixed tree, then the tail in order.
__device__ __forceinline__ double numpy_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// One thread per detection: area of the box spanned by ALL its (x, y) entries, zeros of missing joints included
// (pycocotools coco.py loadRes, keypoints branch).
__global__ void oks_det_area_kernel(const double* __restrict__ det_kp, double* __restrict__ det_area, int D, int K) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const double* p = det_kp + (size_t)d * K * 3;
    double x0 = p[0], x1 = p[0], y0 = p[1], y1 = p[1];
    for (int k = 1; k < K; ++k) {
        double x = p[3 * k], y = p[3 * k + 1];
        x0 = fmin(x0, x); x1 = fmax(x1, x);
        y0 = fmin(y0, y); y1 = fmax(y1, y);
    }
    det_area[d] = (x1 - x0) * (y1 - y0);
}

// One thread per (detection, ground truth) pair of the same group: COCOeval.computeOks.
//   e_k = (dx^2 + dy^2) / (2 sigma_k)^2 / (area + eps) / 2 ; OKS = sum_k exp(-e_k) / n over the labelled joints, or,
//   for a GT without labelled joints, over all K with the distance measured to the doubled GT box.
__global__ void oks_matrix_kernel(const double* __restrict__ det_kp, const double* __restrict__ gt_kp,
                                  const double* __restrict__ gt_bbox, const double* __restrict__ gt_area,
                                  const int* __restrict__ det_off, const int* __restrict__ gt_off,
                                  const long long* __restrict__ pair_off, const double* __restrict__ sigmas,
                                  double* __restrict__ oks, int Q, int K) {
    const long long total = pair_off[Q];
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int q = upper_group(pair_off, Q, p);
        const int ng = gt_off[q + 1] - gt_off[q];
        const int local = (int)(p - pair_off[q]);
        const int d = det_off[q] + local / ng, g = gt_off[q] + local % ng;
        const double* dk = det_kp + (size_t)d * K * 3;
        const double* gk = gt_kp + (size_t)g * K * 3;
        int k1 = 0;
        for (int k = 0; k < K; ++k) k1 += gk[3 * k + 2] > 0. ? 1 : 0;
        const double bx = gt_bbox[4 * g], by = gt_bbox[4 * g + 1], bw = gt_bbox[4 * g + 2], bh = gt_bbox[4 * g + 3];
        const double x0 = bx - bw, x1 = bx + bw * 2, y0 = by - bh, y1 = by + bh * 2;
        const double area = gt_area[g] + 2.220446049250313e-16;     // np.spacing(1)
        double term[kOksMaxK];
        int n = 0;
        for (int k = 0; k < K; ++k) {
            const double xd = dk[3 * k], yd = dk[3 * k + 1];
            double dx, dy;
            if (k1 > 0) {
                if (!(gk[3 * k + 2] > 0.)) continue;
                dx = xd - gk[3 * k];
                dy = yd - gk[3 * k + 1];
            } else {
                dx = fmax(0., x0 - xd) + fmax(0., xd - x1);
                dy = fmax(0., y0 - yd) + fmax(0., yd - y1);
            }
            const double s2 = sigmas[k] * 2;
            // the reference evaluates (dx**2 + dy**2) / vars / area / 2 left to right: no FMA contraction, no reciprocal
            const double e = __ddiv_rn(__ddiv_rn(__ddiv_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(s2, s2)), area), 2.);
            term[n++] = exp(-e);
        }
        oks[p] = numpy_sum(term, n) / (double)n;
    }
}

// One thread per (group, area range, OKS threshold): COCOeval.evaluateImg's greedy matching.  Detections are visited
// in score order; each takes the best still-free ground truth with OKS >= max(threshold, best so far), visiting the
// counted ground truths first and the ignored ones (crowd / unlabelled / outside the area range) only when none of
// the counted ones matched; crowd ground truths may be matched repeatedly.
//   dt_match [A][T][D]  global GT index + 1, 0 = unmatched          dt_ignore [A][T][D]
//   gt_ignore_out [A][G]                                             gt_taken [A][T][G] scratch (no initialisation needed)
__global__ void oks_match_kernel(const double* __restrict__ oks, const long long* __restrict__ pair_off,
                                 const int* __restrict__ det_off, const int* __restrict__ gt_off,
                                 const double* __restrict__ det_area, const double* __restrict__ gt_area,
                                 const unsigned char* __restrict__ gt_flags, const double* __restrict__ area_rng,
                                 const double* __restrict__ iou_thrs, int Q, int A, int T, int D, int G,
                                 int* __restrict__ dt_match, unsigned char* __restrict__ dt_ignore,
                                 unsigned char* __restrict__ gt_ignore_out, unsigned char* __restrict__ gt_taken) {
    const long long total = (long long)Q * A * T;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(w % T), a = (int)((w / T) % A), q = (int)(w / ((long long)T * A));
        const int d0 = det_off[q], d1 = det_off[q + 1], g0 = gt_off[q], g1 = gt_off[q + 1], ng = g1 - g0;
        const double lo = area_rng[2 * a], hi = area_rng[2 * a + 1];
        const size_t at = (size_t)a * T + t;
        unsigned char* taken = gt_taken + at * G;
        for (int g = g0; g < g1; ++g) {
            taken[g] = 0;
            if (t == 0) gt_ignore_out[(size_t)a * G + g] = ((gt_flags[g] & 1) || gt_area[g] < lo || gt_area[g] > hi) ? 1 : 0;
        }
        const double* blk = oks + pair_off[q];
        for (int d = d0; d < d1; ++d) {
            double best = fmin(iou_thrs[t], 1 - 1e-10);
            int m = -1, m_ig = 0;
            for (int pass = 0; pass < 2 && !(pass == 1 && m >= 0); ++pass) {
                for (int g = g0; g < g1; ++g) {
                    const int ig = ((gt_flags[g] & 1) || gt_area[g] < lo || gt_area[g] > hi) ? 1 : 0;
                    if (ig != pass) continue;
                    if (taken[g] && !(gt_flags[g] & 2)) continue;
                    const double v = blk[(size_t)(d - d0) * ng + (g - g0)];
                    if (v < best) continue;
                    best = v;
                    m = g;
                    m_ig = ig;
                }
            }
            const size_t o = at * D + d;
            if (m >= 0) {
                taken[m] = 1;
                dt_match[o] = m + 1;
                dt_ignore[o] = (unsigned char)m_ig;
            } else {
                dt_match[o] = 0;
                dt_ignore[o] = (det_area[d] < lo || det_area[d] > hi) ? 1 : 0;
            }
        }
    }
}

// One CTA per (category, area range, threshold): COCOeval.accumulate.  `order` lists the category's detections by
// descending score (stable); the CTA scans them once for the running TP / FP counts, once backwards for the
// precision envelope (running maximum from the right), then samples the envelope at the R recall thresholds.
//   precision [T][R][C][A], recall [T][C][A]   (-1 where the category has no counted ground truth / nothing at all)
//   pr_ws, rc_ws [A*T][D] fp64 scratch
__global__ void __launch_bounds__(kApThreads)
ap_accumulate_kernel(const long long* __restrict__ order, const int* __restrict__ dt_match,
                     const unsigned char* __restrict__ dt_ignore, const unsigned char* __restrict__ gt_ignore,
                     const int* __restrict__ cat_det_off, const int* __restrict__ cat_gt_off,
                     const double* __restrict__ rec_thrs, int C, int A, int T, int R, int D, int G,
                     double* __restrict__ precision, double* __restrict__ recall,
                     double* __restrict__ pr_ws, double* __restrict__ rc_ws) {
    const int t = blockIdx.x % T, a = (blockIdx.x / T) % A, c = blockIdx.x / (T * A);
    const int tid = threadIdx.x;
    const int d0 = cat_det_off[c], nd = cat_det_off[c + 1] - d0, g0 = cat_gt_off[c], ngt = cat_gt_off[c + 1] - g0;
    __shared__ int s_a[kApThreads], s_b[kApThreads];
    __shared__ double s_m[kApThreads];
    __shared__ int s_carry[2];

    // counted ground truths of this category in this area range
    int cnt = 0;
    for (int g = tid; g < ngt; g += kApThreads) cnt += gt_ignore[(size_t)a * G + g0 + g] == 0 ? 1 : 0;
    s_a[tid] = cnt;
    __syncthreads();
    for (int s = kApThreads / 2; s > 0; s >>= 1) {
        if (tid < s) s_a[tid] += s_a[tid + s];
        __syncthreads();
    }
    const int npig = s_a[0];
    __syncthreads();

    auto put = [&](int r, double v) { precision[(((size_t)t * R + r) * C + c) * A + a] = v; };
    if ((nd == 0 && ngt == 0) || npig == 0) {
        for (int r = tid; r < R; r += kApThreads) put(r, -1.);
        if (tid == 0) recall[((size_t)t * C + c) * A + a] = -1.;
        return;
    }

    const size_t at = (size_t)a * T + t;
    double* pr = pr_ws + at * D + d0;
    double* rc = rc_ws + at * D + d0;
    if (tid == 0) s_carry[0] = s_carry[1] = 0;
    __syncthreads();
    for (int base = 0; base < nd; base += kApThreads) {
        const int i = base + tid;
        int tp = 0, fp = 0;
        if (i < nd) {
            const size_t o = at * D + (size_t)order[d0 + i];
            const int ig = dt_ignore[o], mt = dt_match[o] != 0;
            tp = (mt && !ig) ? 1 : 0;
            fp = (!mt && !ig) ? 1 : 0;
        }
        s_a[tid] = tp;
        s_b[tid] = fp;
        __syncthreads();
        for (int s = 1; s < kApThreads; s <<= 1) {          // Hillis-Steele inclusive scan of both counts
            int va = 0, vb = 0;
            if (tid >= s) { va = s_a[tid - s]; vb = s_b[tid - s]; }
            __syncthreads();
            s_a[tid] += va;
            s_b[tid] += vb;
            __syncthreads();
        }
        if (i < nd) {
            const double tps = (double)(s_carry[0] + s_a[tid]), fps = (double)(s_carry[1] + s_b[tid]);
            rc[i] = tps / (double)npig;
            pr[i] = tps / (fps + tps + 2.220446049250313e-16);
        }
        __syncthreads();
        if (tid == kApThreads - 1) { s_carry[0] += s_a[tid]; s_carry[1] += s_b[tid]; }
        __syncthreads();
    }

    // envelope: pr[i] = max(pr[i], pr[i+1], ...), chunk by chunk from the right
    double carry = -1.;
    for (int base = ((nd - 1) / kApThreads) * kApThreads; base >= 0; base -= kApThreads) {
        const int i = base + tid;
        s_m[tid] = i < nd ? pr[i] : -1.;
        __syncthreads();
        for (int s = 1; s < kApThreads; s <<= 1) {          // inclusive suffix maximum
            double v = -1.;
            if (tid + s < kApThreads) v = s_m[tid + s];
            __syncthreads();
            s_m[tid] = fmax(s_m[tid], v);
            __syncthreads();
        }
        if (i < nd) pr[i] = fmax(s_m[tid], carry);
        carry = fmax(carry, s_m[0]);
        __syncthreads();
    }
    __threadfence_block();
    __syncthreads();

    for (int r = tid; r < R; r += kApThreads) {              // np.searchsorted(rc, thr, side='left')
        const double thr = rec_thrs[r];
        int lo = 0, hi = nd;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (rc[mid] < thr) lo = mid + 1; else hi = mid;
        }
        put(r, lo < nd ? pr[lo] : 0.);
    }
    if (tid == 0) recall[((size_t)t * C + c) * A + a] = nd ? rc[nd - 1] : 0.;
}

}  // namespace pose
