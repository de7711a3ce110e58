// SPM (single-stage multi-person) hot path kernels: target render, loss fwd/bwd, root NMS + displacement decode.
#pragma once
#include "common.cuh"

namespace pose {

constexpr int kSpmThreads = 256;
constexpr int kSpmMaxPersonsSmem = 64;   // persons of one image staged in shared memory per pass

// ---------------------------------------------------------------- render
// SPMHeatmapGenerator (utils/spm_utils.py:29-47): root map = max over persons of the Gaussian patch, no clamp,
//   skip iff cx<=0 and cy<=0.
// SPMMaskGenerator (:57-71): box [cx-half, cx+half+1) x [cy-half, cy+half+1), half = int((6s+2)/2).
// SPMDisplacementGenerator (:84-95): per person p (in order), per joint j not (x<=0 and y<=0):
//   disp[2j]   = fp32( fp64(disp[2j])   + mask_p * (xj - col) / z )
//   disp[2j+1] = fp32( fp64(disp[2j+1]) + mask_p * (yj - row) / z ),   z = sqrt(2 R^2)
// One thread owns 4 consecutive pixels of the plane and walks all 1+2K channels; stores are coalesced
// float4 streams (the tensor is ~97% zeros, the kernel is a write stream).
struct SpmRenderParams {
    const long long* centers;   // [N][Pmax][2]
    const long long* joints;    // [N][Pmax][K][2]
    const int* counts;          // [N]
    float* target;              // [N][1+2K][R][R]
    const float* lut; int lut_n;
    double three_sigma; int half;
    double z;
    int N, Pmax, K, R;
};

__global__ void __launch_bounds__(kSpmThreads) spm_render_kernel(SpmRenderParams P) {
    extern __shared__ float lut_s[];
    __shared__ int s_cx[kSpmMaxPersonsSmem], s_cy[kSpmMaxPersonsSmem];
    const int quads = P.R * P.R / 4;
    const int ctas_per_img = (quads + kSpmThreads - 1) / kSpmThreads;
    const int img = blockIdx.x / ctas_per_img;
    const int q = (blockIdx.x - img * ctas_per_img) * kSpmThreads + threadIdx.x;
    for (int i = threadIdx.x; i < P.lut_n * P.lut_n; i += blockDim.x) lut_s[i] = P.lut[i];

    const int np = min(max(P.counts[img], 0), P.Pmax);
    const int row = (q * 4) / P.R, col0 = (q * 4) - row * P.R;
    const bool active = q < quads;
    const long long plane = (long long)P.R * P.R;
    float* out = P.target + (long long)img * (1 + 2 * P.K) * plane + (long long)q * 4;

    float root[4] = {0.f, 0.f, 0.f, 0.f};
    // displacement accumulators are kept per channel pair inside the channel loop; persons are staged
    // through shared memory in chunks so Pmax is unbounded
    bool any_cover = false;
    for (int p0 = 0; p0 < np; p0 += kSpmMaxPersonsSmem) {
        const int pc = min(kSpmMaxPersonsSmem, np - p0);
        __syncthreads();
        for (int i = threadIdx.x; i < pc; i += blockDim.x) {
            s_cx[i] = (int)P.centers[((long long)img * P.Pmax + p0 + i) * 2];
            s_cy[i] = (int)P.centers[((long long)img * P.Pmax + p0 + i) * 2 + 1];
        }
        __syncthreads();
        if (!active) continue;
        for (int p = 0; p < pc; ++p) {
            const int cx = s_cx[p], cy = s_cy[p];
            if (cx <= 0 && cy <= 0) continue;
            // root Gaussian
            const int ulx = (int)rint(((double)cx - P.three_sigma) - 1.0), uly = (int)rint(((double)cy - P.three_sigma) - 1.0);
            const int brx = (int)rint(((double)cx + P.three_sigma) + 2.0), bry = (int)rint(((double)cy + P.three_sigma) + 2.0);
            const int gy = row - uly;
            if (row >= max(0, uly) && row < min(bry, P.R) && gy < P.lut_n) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = col0 + j, gx = c - ulx;
                    if (c >= max(0, ulx) && c < min(brx, P.R) && gx < P.lut_n) root[j] = fmaxf(root[j], lut_s[gy * P.lut_n + gx]);
                }
            }
            if (row >= cy - P.half && row < cy + P.half + 1 && col0 + 3 >= cx - P.half && col0 < cx + P.half + 1) any_cover = true;
        }
    }
    if (!active) return;
    __stcs(reinterpret_cast<float4*>(out), make_float4(root[0], root[1], root[2], root[3]));

    if (!any_cover) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 1; c <= 2 * P.K; ++c) __stcs(reinterpret_cast<float4*>(out + c * plane), z4);
        return;
    }
    // covered pixels (a few percent of the plane): walk persons in order for every joint
    for (int j = 0; j < P.K; ++j) {
        float ax[4] = {0.f, 0.f, 0.f, 0.f}, ay[4] = {0.f, 0.f, 0.f, 0.f};
        for (int p = 0; p < np; ++p) {
            const long long* cp = P.centers + ((long long)img * P.Pmax + p) * 2;
            const int cx = (int)cp[0], cy = (int)cp[1];
            if (cx <= 0 && cy <= 0) continue;
            if (!(row >= max(0, cy - P.half) && row < min(P.R, cy + P.half + 1))) continue;
            const long long* jp = P.joints + (((long long)img * P.Pmax + p) * P.K + j) * 2;
            const long long jx = jp[0], jy = jp[1];
            if (jx <= 0 && jy <= 0) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = col0 + e;
                if (c >= max(0, cx - P.half) && c < min(P.R, cx + P.half + 1)) {
                    ax[e] = (float)((double)ax[e] + (double)(jx - (long long)c) / P.z);
                    ay[e] = (float)((double)ay[e] + (double)(jy - (long long)row) / P.z);
                }
            }
        }
        __stcs(reinterpret_cast<float4*>(out + (1 + 2 * j) * plane), make_float4(ax[0], ax[1], ax[2], ax[3]));
        __stcs(reinterpret_cast<float4*>(out + (2 + 2 * j) * plane), make_float4(ay[0], ay[1], ay[2], ay[3]));
    }
}

// ---------------------------------------------------------------- loss
// SPMLoss.forward (models/loss/spm_loss.py:32-83), closed form per pixel:
//   m = (t0 > 0);  root: (sig(p0) m - t0)^2;  disp: SmoothL1_{beta=1}(tanh(p) m - t)
//   dL/dp0 = lr 2 (s m - t0) m s (1-s) inv_norm;  dL/dp = ld clip(d,-1,1) m (1 - tanh^2) inv_norm
struct SpmLossParams {
    const float* logits; const float* target; float* dlogits;
    double* partials;            // [grid][2]  (S_root, S_disp)
    long long units;             // N * R*R/4
    int quads;                   // R*R/4
    int C;                       // 1 + 2K
    float groot, gdisp;          // 2*lambda_root*inv_norm, lambda_disp*inv_norm
};

template <bool GRAD>
__global__ void __launch_bounds__(kSpmThreads) spm_loss_kernel(SpmLossParams P) {
    __shared__ double red[kSpmThreads / 32][2];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long plane4 = P.quads;    // float4 per channel plane
    double droot = 0.0, ddisp = 0.0;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < P.units; u += stride) {
        const long long img = u / P.quads;
        const int q = (int)(u - img * P.quads);
        const float4* lp = reinterpret_cast<const float4*>(P.logits) + img * P.C * plane4 + q;
        const float4* tp = reinterpret_cast<const float4*>(P.target) + img * P.C * plane4 + q;
        float4* gp = GRAD ? reinterpret_cast<float4*>(P.dlogits) + img * P.C * plane4 + q : nullptr;
        const float4 p0 = ldg_stream(lp), t0 = ldg_stream(tp);
        const float pv[4] = {p0.x, p0.y, p0.z, p0.w}, tv[4] = {t0.x, t0.y, t0.z, t0.w};
        bool m[4];
        float g0[4];
        float aroot = 0.f, adisp = 0.f;
        bool anym = false;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            m[e] = tv[e] > 0.0f;
            anym |= m[e];
            const float s = sigmoid_fast(pv[e]);
            const float d = (m[e] ? s : s * 0.0f) - tv[e];
            aroot = fmaf(d, d, aroot);
            g0[e] = m[e] ? P.groot * d * ((1.0f - s) * s) : 0.0f;
        }
        if (GRAD) __stcs(gp, make_float4(g0[0], g0[1], g0[2], g0[3]));
        constexpr int CU = 4;
        for (int c = 1; c < P.C; c += CU) {
            float4 pc[CU], tc[CU];
#pragma unroll
            for (int k = 0; k < CU; ++k)
                if (c + k < P.C) { pc[k] = ldg_stream(lp + (c + k) * plane4); tc[k] = ldg_stream(tp + (c + k) * plane4); }
#pragma unroll
            for (int k = 0; k < CU; ++k) {
                if (c + k >= P.C) break;
                const float pe[4] = {pc[k].x, pc[k].y, pc[k].z, pc[k].w}, te[4] = {tc[k].x, tc[k].y, tc[k].z, tc[k].w};
                float ge[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // tanh only where the root mask is set; elsewhere tanh(p)*0 == 0 for every finite or infinite p
                    float th = 0.0f, pm = pe[e] != pe[e] ? pe[e] : 0.0f;    // NaN logits propagate as in the reference
                    if (m[e]) { th = tanhf(pe[e]); pm = th; }
                    const float d = pm - te[e];
                    const float ad = fabsf(d);
                    adisp += ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
                    ge[e] = m[e] ? P.gdisp * fminf(fmaxf(d, -1.0f), 1.0f) * (1.0f - th * th) : 0.0f;
                }
                if (GRAD) __stcs(gp + (c + k) * plane4, make_float4(ge[0], ge[1], ge[2], ge[3]));
            }
        }
        droot += (double)aroot;
        ddisp += (double)adisp;
    }
    droot = warp_sum(droot);
    ddisp = warp_sum(ddisp);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[wid][0] = droot; red[wid][1] = ddisp; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < kSpmThreads / 32; ++w) { a += red[w][0]; b += red[w][1]; }
        P.partials[2 * blockIdx.x] = a;
        P.partials[2 * blockIdx.x + 1] = b;
    }
}

// ---------------------------------------------------------------- decode
// nms_spm (utils/spm_utils.py:112-161): candidates conf > thr, repeatedly take the best remaining candidate
// (ties: lowest row-major index) and drop everything within dist_threshold of it (strict: d > thr survives).
// get_spm_keypoints (:175-200) + DecodeSPM.forward (:228-248) on the picked roots.
// One CTA per image; the activated root map lives in shared memory, suppressed / non-candidate pixels = -inf.
struct SpmDecodeParams {
    const float* x;              // [N][C][R][R]
    float* roots; float* kps; int* counts; int* counts_total;
    int N, Pmax, K, R, C;
    float thr; double dist_thr; int apply_act;
    float zf;                    // fp32(sqrt(2 R^2))
    float input_size;            // DecodeSPM.input_size
};

// One body joint of one root (get_spm_keypoints utils/spm_utils.py:187-197 + the rescale at :247-248):
//   kx = disp[2k][y,x] * fp32(z) + x   (fp32 multiply, then fp32 add: two roundings, no FMA)
//   d  = sqrt_fp64( fp32((x-kx)^2 + (y-ky)^2) );  d < dist_thr -> (0,0,0) else (kx, ky, conf) * num / den
__device__ __forceinline__ void spm_joint(const float* __restrict__ disp, long long plane, int pix, int rx, int ry, float conf,
                                          int k, int apply_act, float zf, double dist_thr, float num, float den, float* o) {
    float dx = __ldg(disp + (2 * k) * plane + pix);
    float dy = __ldg(disp + (2 * k + 1) * plane + pix);
    if (apply_act) { dx = tanhf(dx); dy = tanhf(dy); }
    const float fx = (float)rx, fy = (float)ry;
    const float kx = __fadd_rn(__fmul_rn(dx, zf), fx);
    const float ky = __fadd_rn(__fmul_rn(dy, zf), fy);
    const float ex = __fsub_rn(fx, kx), ey = __fsub_rn(fy, ky);
    const float q = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    if (sqrt((double)q) < dist_thr) {
        o[0] = 0.f; o[1] = 0.f; o[2] = 0.f;
    } else {
        o[0] = __fdiv_rn(__fmul_rn(kx, num), den);
        o[1] = __fdiv_rn(__fmul_rn(ky, num), den);
        o[2] = conf;
    }
}

// get_spm_keypoints drop-in: roots [n][3] (x, y, conf in map pixels) + activated displacements [2K][R][R] -> [n][K][3]
__global__ void __launch_bounds__(128) spm_gather_kernel(const float* __restrict__ roots, const float* __restrict__ disp,
                                                         float* __restrict__ kps, int n, int K, int R, float zf, double dist_thr) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * K) return;
    const int r = t / K, k = t - r * K;
    const float x = roots[3 * r], y = roots[3 * r + 1], c = roots[3 * r + 2];
    const int xi = (int)x, yi = (int)y;                      // .long() truncation in the reference
    if (xi < 0 || xi >= R || yi < 0 || yi >= R) { kps[3 * t] = kps[3 * t + 1] = kps[3 * t + 2] = 0.f; return; }
    // the reference adds the fp32 root coordinate itself (not the truncated one)
    float dx = __ldg(disp + (long long)(2 * k) * R * R + yi * R + xi);
    float dy = __ldg(disp + (long long)(2 * k + 1) * R * R + yi * R + xi);
    const float kx = __fadd_rn(__fmul_rn(dx, zf), x), ky = __fadd_rn(__fmul_rn(dy, zf), y);
    const float ex = __fsub_rn(x, kx), ey = __fsub_rn(y, ky);
    const float q = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    float* o = kps + 3 * t;
    if (sqrt((double)q) < dist_thr) { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
    else { o[0] = kx; o[1] = ky; o[2] = c; }
}

__global__ void __launch_bounds__(kSpmThreads) spm_decode_kernel(SpmDecodeParams P) {
    extern __shared__ float hmap[];                 // R*R
    __shared__ float s_v[kSpmThreads / 32];
    __shared__ int s_i[kSpmThreads / 32];
    __shared__ float s_best;
    __shared__ int s_besti;
    const int img = blockIdx.x;
    const int RR = P.R * P.R;
    const long long plane = RR;
    const float* base = P.x + (long long)img * P.C * plane;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < RR; i += blockDim.x) {
        const float v = ldg_stream(base + i);
        const float h = P.apply_act ? sigmoid_fast(v) : v;
        hmap[i] = h > P.thr ? h : -INFINITY;
    }
    __syncthreads();

    const int rad = (int)floor(P.dist_thr);
    const int side = 2 * rad + 1;
    int found = 0;
    float* roots = P.roots + (long long)img * P.Pmax * 3;
    float* kps = P.kps + (long long)img * P.Pmax * P.K * 3;
    while (true) {
        float best = -INFINITY;
        int besti = 0x7fffffff;
        for (int i = threadIdx.x; i < RR; i += blockDim.x) {
            const float v = hmap[i];
            if (v > best) { best = v; besti = i; }
        }
        warp_argmax_first(best, besti);
        if (lane == 0) { s_v[wid] = best; s_i[wid] = besti; }
        __syncthreads();
        if (wid == 0) {
            best = lane < kSpmThreads / 32 ? s_v[lane] : -INFINITY;
            besti = lane < kSpmThreads / 32 ? s_i[lane] : 0x7fffffff;
            warp_argmax_first(best, besti);
            if (lane == 0) { s_best = best; s_besti = besti; }
        }
        __syncthreads();
        best = s_best; besti = s_besti;
        if (!(best > -INFINITY)) break;
        const int ry = besti / P.R, rx = besti - ry * P.R;
        if (found < P.Pmax) {
            // root row + its K joints: one thread per joint
            if (threadIdx.x == 0) {
                roots[found * 3 + 0] = __fdiv_rn(__fmul_rn((float)rx, P.input_size), (float)P.R);
                roots[found * 3 + 1] = __fdiv_rn(__fmul_rn((float)ry, P.input_size), (float)P.R);
                roots[found * 3 + 2] = best;
            }
            for (int k = threadIdx.x; k < P.K; k += blockDim.x) {
                spm_joint(base + plane, plane, besti, rx, ry, best, k, P.apply_act, P.zf, P.dist_thr, P.input_size, (float)P.R,
                          kps + ((long long)found * P.K + k) * 3);
            }
        }
        ++found;
        // suppress the disc (survivors satisfy sqrt(dx^2+dy^2) > dist_thr)
        for (int t = threadIdx.x; t < side * side; t += blockDim.x) {
            const int oy = t / side - rad, ox = t - (t / side) * side - rad;
            const int y = ry + oy, x = rx + ox;
            if (y < 0 || y >= P.R || x < 0 || x >= P.R) continue;
            if (!(sqrt((double)(ox * ox + oy * oy)) > P.dist_thr)) hmap[y * P.R + x] = -INFINITY;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        P.counts[img] = min(found, P.Pmax);
        if (P.counts_total) P.counts_total[img] = found;
    }
}

}  // namespace pose
