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
// The tensor is ~97% zeros: the kernel is a write stream (see spm_render_kernel).
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

struct SpmPerson {
    int cx, cy;                 // centre
    int ulx, uly, brx, bry;     // Gaussian patch corners (fp64 half-to-even rounding done once per person and work unit)
};

constexpr int kSpmRenderU = 4;
constexpr int kSpmRenderChunk = kSpmThreads * kSpmRenderU;    // float4 per work unit (16 KB of one plane)

// Render = two launches.
//  (1) spm_fill_kernel: ONE linear write stream over the whole target, plane by plane (unit = 16 KB of a plane).
//      Displacement planes are written as zeros with no other work; root planes stage the image's persons (integer
//      patch geometry + a per-row bitmask of the persons touching each row) and take the max of the template patches.
//  (2) spm_patch_kernel: the ~3% of displacement pixels inside some person's box, one thread per (joint, pixel):
//      full lane efficiency where v1-v3 ran a 32-lane warp for 3 covered lanes (profiles/: that path, not the 587 MB of
//      stores, set the run time).
__global__ void __launch_bounds__(kSpmThreads) spm_fill_kernel(SpmRenderParams P) {
    extern __shared__ float lut_s[];
    __shared__ SpmPerson s_p[kSpmMaxPersonsSmem];
    __shared__ unsigned long long s_rootmask[72];                       // per row of the unit: persons whose patch touches it
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < P.lut_n * P.lut_n; i += blockDim.x) lut_s[i] = P.lut[i];
    const int C = 1 + 2 * P.K;
    const int quads = P.R * P.R / 4;
    const int qpr = P.R / 4;                                            // quads per row
    const int upp = (quads + kSpmRenderChunk - 1) / kSpmRenderChunk;
    const long long units = (long long)P.N * C * upp;
    float4* out4 = reinterpret_cast<float4*>(P.target);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const long long plane = unit / upp;
        const int chunk = (int)(unit - plane * upp);
        const int img = (int)(plane / C), c = (int)(plane - (long long)img * C);
        const int q_lo = chunk * kSpmRenderChunk, q_hi = min(quads, q_lo + kSpmRenderChunk);   // [q_lo, q_hi)
        float4* dst = out4 + plane * quads;
        if (c != 0) {                                                   // CTA-uniform: 34 of 35 planes
#pragma unroll
            for (int u = 0; u < kSpmRenderU; ++u) {
                const int q = q_lo + u * kSpmThreads + threadIdx.x;
                if (q < q_hi) __stcs(dst + q, z4);
            }
            continue;
        }
        // root plane: SPMHeatmapGenerator -- max with each person's template patch, clipped to the map (no clamp of the centre)
        const int np = min(max(P.counts[img], 0), P.Pmax);
        const int row_lo = q_lo / qpr, nrows = (q_hi - 1) / qpr - row_lo + 1;                    // <= 66 rows
        float a[kSpmRenderU][4];
#pragma unroll
        for (int u = 0; u < kSpmRenderU; ++u) a[u][0] = a[u][1] = a[u][2] = a[u][3] = 0.0f;
        for (int p0 = 0; p0 < np; p0 += kSpmMaxPersonsSmem) {
            const int pc = min(kSpmMaxPersonsSmem, np - p0);
            __syncthreads();
            for (int i = threadIdx.x; i < pc; i += blockDim.x) {
                const long long pi = (long long)img * P.Pmax + p0 + i;
                SpmPerson sp;
                sp.cx = (int)P.centers[pi * 2];
                sp.cy = (int)P.centers[pi * 2 + 1];
                sp.ulx = (int)rint(((double)sp.cx - P.three_sigma) - 1.0);
                sp.uly = (int)rint(((double)sp.cy - P.three_sigma) - 1.0);
                sp.brx = (int)rint(((double)sp.cx + P.three_sigma) + 2.0);
                sp.bry = (int)rint(((double)sp.cy + P.three_sigma) + 2.0);
                s_p[i] = sp;
            }
            __syncthreads();
            for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
                const int row = row_lo + r;
                unsigned long long m = 0ull;
                for (int p = 0; p < pc; ++p) {
                    const SpmPerson sp = s_p[p];
                    if (sp.cx <= 0 && sp.cy <= 0) continue;
                    if (row >= max(0, sp.uly) && row < min(sp.bry, P.R) && row - sp.uly < P.lut_n) m |= 1ull << p;
                }
                s_rootmask[r] = m;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < kSpmRenderU; ++u) {
                const int q = q_lo + u * kSpmThreads + threadIdx.x;
                if (q >= q_hi) continue;
                const int row = q / qpr, col0 = (q - row * qpr) * 4;
                unsigned long long m = s_rootmask[row - row_lo];
                while (m) {
                    const int p = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const SpmPerson sp = s_p[p];
                    const int gy = row - sp.uly;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int cc = col0 + e, gx = cc - sp.ulx;
                        if (cc >= max(0, sp.ulx) && cc < min(sp.brx, P.R) && gx < P.lut_n)
                            a[u][e] = fmaxf(a[u][e], lut_s[gy * P.lut_n + gx]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kSpmRenderU; ++u) {
            const int q = q_lo + u * kSpmThreads + threadIdx.x;
            if (q < q_hi) __stcs(dst + q, make_float4(a[u][0], a[u][1], a[u][2], a[u][3]));
        }
    }
}

// SPMMaskGenerator + SPMDisplacementGenerator on the covered pixels.  One CTA per (image, person p); its threads walk
// (joint j, pixel of p's box).  A pixel is handled by the LOWEST-index person whose box covers it (so every covered pixel
// is written exactly once) and that thread replays ALL covering persons in order:
//   disp[2j]   = fp32(fp64(disp[2j])   + (xj - col) / z),   disp[2j+1] = fp32(fp64(disp[2j+1]) + (yj - row) / z)
// skipping persons whose centre or whose joint j is (<=0, <=0) -- the reference's accumulation order and rounding.
constexpr int kSpmPatchPersons = 256;       // centres of one image kept in shared memory (more: read through L1)

__global__ void __launch_bounds__(kSpmThreads) spm_patch_kernel(SpmRenderParams P, int div_n) {
    // dynamic shared memory: quotient table (joint - coord) / z for every integer difference in [-R, R] -- exactly the fp64
    // values the reference computes, so a covered pixel costs a shared-memory load instead of a software fp64 division
    extern __shared__ __align__(16) double div_s[];
    __shared__ int s_cx[kSpmPatchPersons], s_cy[kSpmPatchPersons];
    const int img = blockIdx.x / P.Pmax, p = blockIdx.x - img * P.Pmax;
    const int np = min(max(P.counts[img], 0), P.Pmax);
    const long long* cen = P.centers + (long long)img * P.Pmax * 2;
    if (p >= np) return;                                                // CTA-uniform
    const int cx = (int)__ldg(cen + 2 * p), cy = (int)__ldg(cen + 2 * p + 1);
    if (cx <= 0 && cy <= 0) return;
    for (int i = threadIdx.x; i < div_n; i += blockDim.x) div_s[i] = (double)(i - P.R) / P.z;
    for (int i = threadIdx.x; i < min(np, kSpmPatchPersons); i += blockDim.x) {
        s_cx[i] = (int)__ldg(cen + 2 * i);
        s_cy[i] = (int)__ldg(cen + 2 * i + 1);
    }
    __syncthreads();
    pdl_wait();                                                         // the zero fill of the same planes comes first
    const int side = 2 * P.half + 1;
    const long long plane = (long long)P.R * P.R;
    float* out = P.target + (long long)img * (1 + 2 * P.K) * plane;
    constexpr int kCov = 8;                                             // covering persons kept in registers per pixel
    // one thread per pixel of this person's box; the set of persons covering a pixel is the same for every joint, so it
    // is found once and the joints are then walked with a handful of instructions each
    const int npix = side * side;
    const int groups = max(1, (int)blockDim.x / npix);                  // joints are split over `groups` threads per pixel
    for (int t = threadIdx.x; t < npix * groups; t += blockDim.x) {
        const int g = t / npix, pi = t - g * npix;
        const int dy = pi / side, dx = pi - dy * side;
        const int row = cy - P.half + dy, col = cx - P.half + dx;
        if (row < 0 || row >= P.R || col < 0 || col >= P.R) continue;
        int cov[kCov];
        int ncov = 0;
        bool mine = true;
        for (int q = 0; q < np; ++q) {
            int qx, qy;
            if (q < kSpmPatchPersons) { qx = s_cx[q]; qy = s_cy[q]; }
            else { qx = (int)__ldg(cen + 2 * q); qy = (int)__ldg(cen + 2 * q + 1); }
            if (qx <= 0 && qy <= 0) continue;
            if (row < qy - P.half || row > qy + P.half || col < qx - P.half || col > qx + P.half) continue;
            if (q < p) { mine = false; break; }                         // an earlier person's CTA owns this pixel
            if (ncov < kCov) cov[ncov] = q;
            ++ncov;
        }
        if (!mine) continue;
        const long long* jbase = P.joints + (long long)img * P.Pmax * P.K * 2;
        float* o = out + (long long)row * P.R + col;
        for (int j = g; j < P.K; j += groups) {
            float ax = 0.0f, ay = 0.0f;
            auto add = [&](int q) {
                const longlong2 jv = __ldg(reinterpret_cast<const longlong2*>(jbase + ((long long)q * P.K + j) * 2));
                if (jv.x <= 0 && jv.y <= 0) return;
                const long long ddx = jv.x - (long long)col, ddy = jv.y - (long long)row;
                const double qx_ = (div_n && ddx >= -P.R && ddx <= P.R) ? div_s[(int)ddx + P.R] : (double)ddx / P.z;
                const double qy_ = (div_n && ddy >= -P.R && ddy <= P.R) ? div_s[(int)ddy + P.R] : (double)ddy / P.z;
                ax = (float)((double)ax + qx_);
                ay = (float)((double)ay + qy_);
            };
            if (ncov <= kCov) {
#pragma unroll
                for (int c = 0; c < kCov; ++c)
                    if (c < ncov) add(cov[c]);
            } else {                                                    // crowded pixel: re-scan the persons in order
                for (int q = p; q < np; ++q) {
                    const int qx = (int)__ldg(cen + 2 * q), qy = (int)__ldg(cen + 2 * q + 1);
                    if (qx <= 0 && qy <= 0) continue;
                    if (row < qy - P.half || row > qy + P.half || col < qx - P.half || col > qx + P.half) continue;
                    add(q);
                }
            }
            o[(1 + 2 * j) * plane] = ax;
            o[(2 + 2 * j) * plane] = ay;
        }
    }
}

// ---------------------------------------------------------------- loss
// SPMLoss.forward (models/loss/spm_loss.py:32-83), closed form per pixel:
//   m = (t0 > 0);  root: (sig(p0) m - t0)^2;  disp: SmoothL1_{beta=1}(tanh(p) m - t)
//   dL/dp0 = lr 2 (s m - t0) m s (1-s) inv_norm;  dL/dp = ld clip(d,-1,1) m (1 - tanh^2) inv_norm
struct SpmLossParams {
    const float* logits; const float* target; float* dlogits;
    double* partials;            // [units][2]  (S_root, S_disp) of every 16 KB unit
    unsigned int* mask;          // [N][mask_words]: bit i of an image = (root target of pixel i > 0), written by spm_root_mask_kernel
    unsigned int* ticket;        // two-level loss reduction counter (zeroed by spm_root_mask_kernel)
    int quads;                   // R*R/4 float4 per plane
    int mask_words;              // 32-bit words per image (a multiple of 4)
    int C;                       // 1 + 2K
    float groot, gdisp;          // 2*lambda_root*inv_norm, lambda_disp*inv_norm
};

constexpr int kSpmLossU = 4;                                  // float4 per thread and tensor in flight
constexpr int kSpmLossChunk = kSpmThreads * kSpmLossU;        // float4 per work unit (1024 -> 16 KB per tensor)
__host__ __device__ inline int spm_loss_units_per_plane(int R) { return (R * R / 4 + kSpmLossChunk - 1) / kSpmLossChunk; }
__host__ __device__ inline int spm_mask_words(int R) { return ((R * R + 31) / 32 + 3) / 4 * 4; }

// Root mask of every image as bits (2 KB per image at R = 128): the displacement planes of the dense loss need m = (t0 > 0) of
// their pixel, and re-reading the image's root TARGET plane for it (r01: a third 128-bit load per quad, through L2) cost the
// streaming kernel 7 % (1 159 vs 1 079 us per 1024 images with the load taken out); here it is 4 bits per quad from a word that
// eight neighbouring threads share.  grid = (ceil(quads / 256), N), 256 threads, one quad per thread.
__global__ void __launch_bounds__(256) spm_root_mask_kernel(SpmLossParams P) {
    pdl_launch_dependents();
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *P.ticket = 0u;
    const int q = blockIdx.x * 256 + threadIdx.x;
    const long long img = blockIdx.y;
    unsigned nib = 0u;
    if (q < P.quads) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(P.target) + img * P.C * P.quads + q);
        nib = (v.x > 0.0f ? 1u : 0u) | (v.y > 0.0f ? 2u : 0u) | (v.z > 0.0f ? 4u : 0u) | (v.w > 0.0f ? 8u : 0u);
    }
    unsigned w = nib << ((threadIdx.x & 7) * 4);
    w |= __shfl_xor_sync(FULL_MASK, w, 1);
    w |= __shfl_xor_sync(FULL_MASK, w, 2);
    w |= __shfl_xor_sync(FULL_MASK, w, 4);
    if ((threadIdx.x & 7) == 0 && (q >> 3) < P.mask_words) P.mask[img * P.mask_words + (q >> 3)] = w;
}

// NOT persistent (round 2, later): one 256-thread CTA per 16 KB unit of one channel plane, grid = (units per plane, 1+2K, N) --
// no index divisions, CTAs handed out in memory order; every thread's loads (4 quads of logits + 4 of target) are issued before
// anything else.  One fp64 (S_root, S_disp) pair per unit, reduced in a fixed order by spm_loss_reduce_kernel.  Against the
// persistent r01 form (every CTA strode over the units, root mask from the target's root plane): 1 189 -> 1 017.6 us per 1024
// images, 282.9 -> 258.6 us per 256 (104-106 % of the measured copy peak).
template <bool GRAD>
__global__ void __launch_bounds__(kSpmThreads) spm_loss_kernel(SpmLossParams P) {
    __shared__ float red[kSpmThreads / 32];
    pdl_launch_dependents();
    const int chunk = blockIdx.x, c = blockIdx.y;
    const long long img = blockIdx.z;
    const long long plane = img * gridDim.y + c;
    const long long unit = plane * gridDim.x + chunk;
    const long long off = plane * P.quads;
    const float4* L4 = reinterpret_cast<const float4*>(P.logits);
    const float4* T4 = reinterpret_cast<const float4*>(P.target);
    float4* G4 = reinterpret_cast<float4*>(P.dlogits);
    float4 pv[kSpmLossU], tv[kSpmLossU];
    int q[kSpmLossU];
#pragma unroll
    for (int u = 0; u < kSpmLossU; ++u) {
        q[u] = chunk * kSpmLossChunk + u * kSpmThreads + threadIdx.x;
        if (q[u] < P.quads) {
            pv[u] = ldg_stream(L4 + off + q[u]);
            tv[u] = ldg_stream(T4 + off + q[u]);
        }
    }
    unsigned mw[kSpmLossU];
    if (c != 0) {
        pdl_wait();                                             // the mask bits of spm_root_mask_kernel
        const unsigned int* mimg = P.mask + img * P.mask_words;
#pragma unroll
        for (int u = 0; u < kSpmLossU; ++u) mw[u] = q[u] < P.quads ? __ldg(mimg + (q[u] >> 3)) : 0u;
    }
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < kSpmLossU; ++u) {
        if (q[u] >= P.quads) break;
        const float pe[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w}, te[4] = {tv[u].x, tv[u].y, tv[u].z, tv[u].w};
        float ge[4];
        if (c == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool m = te[e] > 0.0f;
                const float s = sigmoid_fast(pe[e]);
                const float d = (m ? s : s * 0.0f) - te[e];
                acc = fmaf(d, d, acc);
                ge[e] = m ? P.groot * d * ((1.0f - s) * s) : 0.0f;
            }
        } else {
            const unsigned nib = (mw[u] >> ((q[u] & 7) * 4)) & 15u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // tanh only where the root mask is set; elsewhere tanh(p)*0 == 0 for every finite or infinite p
                const bool m = (nib >> e) & 1u;
                float th = 0.0f, pm = pe[e] != pe[e] ? pe[e] : 0.0f;    // NaN logits propagate as in the reference
                if (m) { th = tanhf(pe[e]); pm = th; }
                const float d = pm - te[e];
                const float ad = fabsf(d);
                acc += ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
                ge[e] = m ? P.gdisp * fminf(fmaxf(d, -1.0f), 1.0f) * (1.0f - th * th) : 0.0f;
            }
        }
        if (GRAD) __stcs(G4 + off + q[u], make_float4(ge[0], ge[1], ge[2], ge[3]));
    }
    acc = warp_sum(acc);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < kSpmThreads / 32; ++w) a += (double)red[w];
        reinterpret_cast<double2*>(P.partials)[unit] = c == 0 ? make_double2(a, 0.0) : make_double2(0.0, a);
    }
}

// ---------------------------------------------------------------- fused render + loss (+ grad)
// SPMHeatmapGenerator + SPMMaskGenerator + SPMDisplacementGenerator (utils/spm_utils.py:16-95) evaluated on the fly and fed
// straight into SPMLoss.forward (models/loss/spm_loss.py:23-105): the [N,1+2K,R,R] target is never written to or read from
// HBM, so a training step moves 2 tensor passes (read logits, write dlogits) instead of 1 (render) + 3 (dense loss).
//
// The SPM target is ~97 % zeros and the zero part costs nothing but bandwidth: target 0, mask 0 => the loss term is 0 unless the
// logit is NaN (sigmoid(p)*0 and tanh(p)*0 are 0 for every other p) and dlogits = 0.  Two launches:
//   (1) spm_geometry_kernel, one CTA per image: everything about an image that is the same for all of its 1+2K planes -- the
//       persons' boxes and Gaussian-patch windows, a per-row bitmask of the persons touching the row, a per-row bitmask of the
//       COVERED float4 quads (quads that intersect some person's box or patch) and, for R <= 256, one byte per pixel (bit 7:
//       root mask t0 > 0; bits 0-6: 0 = in no person's box, p+1 = in the box of person p only, 127 = in several) -- goes to a
//       per-image record in the caller's workspace (20 KB at R = 128; ~5 us for 1024 images);
//   (2) spm_unit_kernel, NOT persistent: one 128-thread CTA per 16 KB unit of one channel plane, handed out in memory order
//       by the hardware scheduler.  One thread issues a cp.async.bulk (TMA engine, mbarrier completion) for the unit's logits
//       the moment the CTA starts; while they fly the CTA fetches the unit's coverage bits and the persons;
//       phase A: every uncovered quad -- NaN check from shared memory, one 128-bit store of zeros;
//       phase B: the covered quads of the unit are listed (deterministic positions from prefix pop-counts, no atomics) and
//       their PIXELS are dealt out one per thread: logit from shared memory, geometry byte, joint, quotient table, tanhf only
//       under the root mask, one 32-bit store.  Pixels in several boxes replay the covering persons in index order.
// Why not persistent (r01: every CTA owned a contiguous range of units and staged an image's geometry in shared memory once
// per image): measured on B200, the same streaming work runs 10-15 % faster when the grid covers the data and CTAs retire
// (tools/stream_patterns.cu), and with the maps in flight held in shared memory by the bulk copies (9 CTAs x 16 KB per SM)
// rather than in registers.  r01: 222.9 us per 256 images (loss + grad), 116 us read-only, 125 us render-only.
struct SpmFusedPerson;
struct SpmGeomLayout {
    unsigned long long stride;      // bytes per image
    unsigned off_persons, off_rowmask, off_covq, off_map;
    int use_map;
};
constexpr int kSpmFusedMaxPersons = 64;                       // one 64-bit row mask
struct __align__(16) SpmFusedPerson {
    int cx, cy;                 // centre (box = [cx-half, cx+half] x [cy-half, cy+half])
    int ulx, uly;               // template origin in map coordinates
    int px0, px1, py0, py1;     // Gaussian patch window clipped to the map and to the template, [x0,x1) x [y0,y1); empty: all 0
};
__host__ __device__ inline SpmGeomLayout spm_geom_layout(int R) {
    SpmGeomLayout L;
    const int wpr = (R / 4 + 31) / 32;
    unsigned long long off = 0;
    L.off_persons = (unsigned)off; off += (unsigned long long)kSpmFusedMaxPersons * sizeof(SpmFusedPerson);
    L.off_rowmask = (unsigned)off; off += (unsigned long long)R * 8;
    L.off_covq = (unsigned)off; off += (unsigned long long)R * wpr * 4;
    off = (off + 15) / 16 * 16;
    L.use_map = (long long)R * R <= 65536;
    L.off_map = (unsigned)off; off += L.use_map ? (unsigned long long)R * R : 0ull;
    L.stride = (off + 255) / 256 * 256;
    return L;
}

struct SpmFusedParams {
    const float* logits; float* dlogits; float* target_out;
    const long long* centers;   // [N][Pmax][2]
    const long long* joints;    // [N][Pmax][K][2]
    const int* counts;          // [N]
    const float* lut; int lut_n;
    double three_sigma; int half; double z;
    double* partials;           // [units][2]  (S_root, S_disp) per work unit
    unsigned int* ticket;       // two-level loss reduction counter (zeroed by the geometry kernel)
    unsigned char* geom;        // [N] per-image records (SpmGeomLayout)
    double* div_tab;            // [div_n] (i - R) / z, or unused when div_n == 0
    SpmGeomLayout gl;
    int N, Pmax, K, R;
    int quads; FastDiv div_qpr; // float4 per plane; quads per row (R/4)
    int wpr;                    // 32-bit words of covered-quad bits per row: ceil(R/4/32)
    int div_n;                  // 2R+1 entries of the quotient table, or 0 (R too large: divide directly)
    float groot, gdisp;         // 2*lambda_root*inv_norm, lambda_disp*inv_norm
};

constexpr int kSpmGeomThreads = 512;
__host__ __device__ inline size_t spm_geom_smem_bytes(int R, int lut_n) {
    const int wpr = (R / 4 + 31) / 32;
    return (size_t)R * 8 + (size_t)R * wpr * 4 + (size_t)lut_n * lut_n * 4;
}

__global__ void __launch_bounds__(kSpmGeomThreads) spm_geometry_kernel(SpmFusedParams P) {
    // dynamic shared memory: [R] u64 row masks | [R*wpr] u32 covered quads | template
    extern __shared__ __align__(16) unsigned char spm_geom_smem[];
    unsigned long long* rowmask_s = reinterpret_cast<unsigned long long*>(spm_geom_smem);
    unsigned int* covq_s = reinterpret_cast<unsigned int*>(rowmask_s + P.R);
    float* lut_s = reinterpret_cast<float*>(covq_s + P.R * P.wpr);
    __shared__ SpmFusedPerson s_p[kSpmFusedMaxPersons];
    pdl_launch_dependents();                                          // the unit kernel's CTAs may be scheduled: they wait for us before reading the records
    const int img = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (img == 0) {
        if (threadIdx.x == 0) *P.ticket = 0u;
        for (int i = threadIdx.x; i < P.div_n; i += blockDim.x) P.div_tab[i] = (double)(i - P.R) / P.z;
    }
    unsigned char* rec = P.geom + (unsigned long long)img * P.gl.stride;
    const int np = min(max(P.counts[img], 0), min(P.Pmax, kSpmFusedMaxPersons));
    for (int i = threadIdx.x; i < P.lut_n * P.lut_n; i += blockDim.x) lut_s[i] = P.lut[i];
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        SpmFusedPerson sp;
        const long long pi = (long long)img * P.Pmax + i;
        sp.cx = (int)max(min(P.centers[pi * 2], 1ll << 30), -(1ll << 30));
        sp.cy = (int)max(min(P.centers[pi * 2 + 1], 1ll << 30), -(1ll << 30));
        sp.ulx = (int)rint(((double)sp.cx - P.three_sigma) - 1.0);
        sp.uly = (int)rint(((double)sp.cy - P.three_sigma) - 1.0);
        const int brx = (int)rint(((double)sp.cx + P.three_sigma) + 2.0);
        const int bry = (int)rint(((double)sp.cy + P.three_sigma) + 2.0);
        sp.px0 = max(0, sp.ulx); sp.px1 = min(min(brx, P.R), sp.ulx + P.lut_n);
        sp.py0 = max(0, sp.uly); sp.py1 = min(min(bry, P.R), sp.uly + P.lut_n);
        if (sp.px1 <= sp.px0 || sp.py1 <= sp.py0) sp.px0 = sp.px1 = sp.py0 = sp.py1 = 0;
        s_p[i] = sp;
        reinterpret_cast<SpmFusedPerson*>(rec + P.gl.off_persons)[i] = sp;
    }
    for (int i = threadIdx.x; i < P.R; i += blockDim.x) rowmask_s[i] = 0ull;
    for (int i = threadIdx.x; i < P.R * P.wpr; i += blockDim.x) covq_s[i] = 0u;
    __syncthreads();
    // one thread per (person, row of the union of its box and patch): row mask bit + covered-quad bits
    const int span = 2 * P.half + 1 + P.lut_n;                 // upper bound on the rows of the union
    for (int t = threadIdx.x; t < np * span; t += blockDim.x) {
        const int p = t / span;
        const SpmFusedPerson sp = s_p[p];
        if (sp.cx <= 0 && sp.cy <= 0) continue;                 // skipped by all three generators
        const bool patch = sp.px1 > sp.px0;
        const int ylo = patch ? min(sp.cy - P.half, sp.py0) : sp.cy - P.half;
        const int yhi = patch ? max(sp.cy + P.half, sp.py1 - 1) : sp.cy + P.half;
        const int row = ylo + (t - p * span);
        if (row < 0 || row >= P.R || row > yhi) continue;
        const int xlo = max(0, patch ? min(sp.cx - P.half, sp.px0) : sp.cx - P.half);
        const int xhi = min(P.R - 1, patch ? max(sp.cx + P.half, sp.px1 - 1) : sp.cx + P.half);
        if (xhi < xlo) continue;
        atomicOr(&rowmask_s[row], 1ull << p);
        const int q0 = xlo >> 2, q1 = xhi >> 2;
        for (int w = q0 >> 5; w <= (q1 >> 5); ++w) {
            const int lo = max(q0 - 32 * w, 0), hi = min(q1 - 32 * w, 31);
            const unsigned int bits = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
            atomicOr(&covq_s[row * P.wpr + w], bits);
        }
    }
    __syncthreads();
    unsigned long long* rowmask_g = reinterpret_cast<unsigned long long*>(rec + P.gl.off_rowmask);
    unsigned int* covq_g = reinterpret_cast<unsigned int*>(rec + P.gl.off_covq);
    for (int i = threadIdx.x; i < P.R; i += blockDim.x) rowmask_g[i] = rowmask_s[i];
    for (int i = threadIdx.x; i < P.R * P.wpr; i += blockDim.x) covq_g[i] = covq_s[i];
    if (P.gl.use_map) {
        // one warp per row; only rows that some person touches are ever looked up, the others are not even written
        unsigned char* map_g = rec + P.gl.off_map;
        for (int row = wid; row < P.R; row += kSpmGeomThreads / 32) {
            const unsigned long long rm = rowmask_s[row];
            if (rm == 0ull) continue;                           // warp-uniform
            for (int col = lane; col < P.R; col += 32) {
                unsigned long long m = rm;
                unsigned int code = 0u, nbox = 0u;
                bool mk = false;
                while (m) {
                    const int p = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const SpmFusedPerson sp = s_p[p];
                    if (row >= sp.py0 && row < sp.py1 && col >= sp.px0 && col < sp.px1 &&
                        lut_s[(row - sp.uly) * P.lut_n + (col - sp.ulx)] > 0.0f) mk = true;
                    if (row >= sp.cy - P.half && row <= sp.cy + P.half && col >= sp.cx - P.half && col <= sp.cx + P.half) {
                        if (nbox++ == 0u) code = (unsigned)p + 1u;
                    }
                }
                if (nbox > 1u) code = 127u;
                map_g[row * P.R + col] = (unsigned char)(code | (mk ? 128u : 0u));
            }
        }
    }
}

// target of one pixel: (root value t0 = max of the covering Gaussian patches, displacement te of the plane's joint / axis);
// s_j[p] = person p's joint of this plane
__device__ __forceinline__ void spm_pixel_target(const SpmFusedParams& P, const SpmFusedPerson* __restrict__ s_p, const int2* __restrict__ s_j,
                                                 unsigned long long m, int row, int col, bool disp, int axis, float& t0, float& te) {
    t0 = 0.0f;
    te = 0.0f;
    while (m) {
        const int p = __ffsll((long long)m) - 1;
        m &= m - 1;
        const SpmFusedPerson sp = s_p[p];
        if (row >= sp.py0 && row < sp.py1 && col >= sp.px0 && col < sp.px1)
            t0 = fmaxf(t0, __ldg(P.lut + (row - sp.uly) * P.lut_n + (col - sp.ulx)));
        if (disp && row >= sp.cy - P.half && row <= sp.cy + P.half && col >= sp.cx - P.half && col <= sp.cx + P.half) {
            const int2 jv = s_j[p];
            if (!(jv.x <= 0 && jv.y <= 0)) {
                const int dd = axis ? jv.y - row : jv.x - col;
                const double qd = (P.div_n && dd >= -P.R && dd <= P.R) ? __ldg(P.div_tab + dd + P.R) : (double)dd / P.z;
                te = (float)((double)te + qd);                       // fp32(fp64(acc) + q): numpy's mixed-precision +=
            }
        }
    }
}

// float4 per work unit (512 ... 4096 = 8 ... 64 KB of one channel plane), per variant class (tools/tune_spm.py,
// profiles/r02_tune_spm_unit_*.log): the kernels that read logits want SMALL units -- a CTA cannot start before its whole
// bulk copy has landed, so 32 KB units cost 9 % and 64 KB units 60 % against 16 KB -- the render-only write stream wants 32 KB
// (fewer CTAs, the per-CTA prologue amortised: 100 vs 109 us per 256 images).
#ifndef POSE_SPM_UNIT_QUADS_LOSS
#define POSE_SPM_UNIT_QUADS_LOSS 1024
#endif
#ifndef POSE_SPM_UNIT_QUADS_RENDER
#define POSE_SPM_UNIT_QUADS_RENDER 2048
#endif
#ifndef POSE_SPM_RO_SCREEN_ALL
#define POSE_SPM_RO_SCREEN_ALL 1
#endif
#ifndef POSE_SPM_UNIT_THREADS
#define POSE_SPM_UNIT_THREADS 128
#endif
#ifndef POSE_SPM_UNIT_MINB_LOSS
#define POSE_SPM_UNIT_MINB_LOSS 10
#endif
#ifndef POSE_SPM_UNIT_MINB_RENDER
#define POSE_SPM_UNIT_MINB_RENDER 6
#endif
__host__ __device__ constexpr int spm_unit_quads(bool loss) { return loss ? POSE_SPM_UNIT_QUADS_LOSS : POSE_SPM_UNIT_QUADS_RENDER; }
static_assert(POSE_SPM_UNIT_QUADS_LOSS % 512 == 0 && POSE_SPM_UNIT_QUADS_LOSS <= 4096 && POSE_SPM_UNIT_QUADS_RENDER % 512 == 0 &&
              POSE_SPM_UNIT_QUADS_RENDER <= 4096, "unit size");
constexpr int kSpmUnitThreads = POSE_SPM_UNIT_THREADS;
constexpr int kSpmUnitWarps = kSpmUnitThreads / 32;
__host__ __device__ inline size_t spm_unit_smem_bytes(bool loss) { return loss ? (size_t)spm_unit_quads(true) * 16 : 0; }
__host__ __device__ inline int spm_units_per_plane(int R, bool loss) { return (R * R / 4 + spm_unit_quads(loss) - 1) / spm_unit_quads(loss); }
// (the loss pairs: one per unit of the kernels that read logits)
__host__ __device__ inline long long spm_units(int N, int K, int R) { return (long long)N * (1 + 2 * K) * spm_units_per_plane(R, true); }

// LOSS = false is the render-only form (pose_spm_render for <= 64 persons per image): no logits are read.
template <bool LOSS, bool GRAD, bool WTGT>
__global__ void __launch_bounds__(kSpmUnitThreads, LOSS ? POSE_SPM_UNIT_MINB_LOSS : POSE_SPM_UNIT_MINB_RENDER) spm_unit_kernel(SpmFusedParams P) {
    constexpr int kSpmUnitQuads = spm_unit_quads(LOSS);
    constexpr int kSpmUnitWords = kSpmUnitQuads / 32;                  // coverage words per unit (<= 128: up to four per lane in the scan)
    extern __shared__ __align__(128) float tile[];                     // LOSS: the unit's logits
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ SpmFusedPerson s_p[kSpmFusedMaxPersons];
    __shared__ int2 s_j[kSpmFusedMaxPersons];
    __shared__ unsigned int s_cov[kSpmUnitWords];                      // coverage of the unit's quads, 32 consecutive quads per word
    __shared__ unsigned short s_list[kSpmUnitQuads];                   // covered quads of the unit (unit-relative), ascending
    __shared__ float s_acc[kSpmUnitWarps];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // grid = (units per plane, 1+2K channels, N images): the hardware hands the CTAs out in memory order, no index divisions
    const int chunk = blockIdx.x, c = blockIdx.y, img = blockIdx.z;
    const long long plane = (long long)img * gridDim.y + c;
    const long long unit = plane * gridDim.x + chunk;
    const int q_lo = chunk * kSpmUnitQuads;
    const int nq = min(kSpmUnitQuads, P.quads - q_lo);                  // quads of this unit
    const long long off = plane * P.quads + q_lo;                       // first quad of the unit in the tensor
    const bool disp = c != 0;
    const int jn = (c - 1) >> 1, axis = (c - 1) & 1;                    // displacement plane: joint and axis (0 = x, 1 = y)
    const int qpr = (int)P.div_qpr.d;

    if (LOSS && tid == 0) {
        mbar_init(smem_u32(&s_bar), 1);
        mbar_init_fence();
        mbar_arrive_expect_tx(smem_u32(&s_bar), (uint32_t)nq * 16u);
        bulk_load(smem_u32(tile), reinterpret_cast<const float4*>(P.logits) + off, (uint32_t)nq * 16u, smem_u32(&s_bar));
    }
    // inputs of the call (not produced by the geometry kernel): the persons' joint of this plane.  All Pmax slots are fetched:
    // waiting for counts[img] first would put one more L2 round trip in front of everything; slots beyond the image's count
    // are never referenced by the geometry.  (Fetching them only in units that turn out to contain covered quads -- 40 % --
    // was measured: no gain for the loss kernels, 9 % slower for the render-only one, which has no bulk copy to hide it under.)
    const int np = min(P.Pmax, kSpmFusedMaxPersons);
    // Read-only form (SCREEN_ALL): phase A needs nothing but the logits, so the prologue's global loads (joints, coverage word,
    // person record -- at most one of each per thread) are only ISSUED here, phase A runs under their latency, and they are
    // parked in shared memory afterwards.  ncu before: 24 % of all stall samples sat behind the barrier that waited for them.
    constexpr bool SCREEN_ALL = LOSS && !GRAD && !WTGT && POSE_SPM_RO_SCREEN_ALL;
    const bool early_a = SCREEN_ALL && qpr == 32;
    float acc = 0.f;
    auto screen_quad = [&](int q) {
        // (the sum of the four is NaN whenever one of them is; +inf + -inf also lands here and adds nothing)
        const float4 v = reinterpret_cast<const float4*>(tile)[q];
        const float sum4 = (v.x + v.y) + (v.z + v.w);
        if (sum4 != sum4) acc += (v.x != v.x ? v.x : 0.f) + (v.y != v.y ? v.y : 0.f) + (v.z != v.z ? v.z : 0.f) + (v.w != v.w ? v.w : 0.f);
    };
    if (early_a) {
        static_assert(kSpmFusedMaxPersons <= kSpmUnitThreads || !SCREEN_ALL, "one person per thread");
        static_assert(kSpmUnitWords <= kSpmUnitThreads || !SCREEN_ALL, "one coverage word per thread");
        longlong2 jv = make_longlong2(0, 0);
        if (disp && tid < np) jv = __ldg(reinterpret_cast<const longlong2*>(P.joints + ((long long)img * P.Pmax * P.K + jn) * 2 + (long long)tid * P.K * 2));
        pdl_wait();                                                     // the image records are complete and visible
        const unsigned char* rec0 = P.geom + (unsigned long long)img * P.gl.stride;
        unsigned cw = 0u;
        if (tid < kSpmUnitWords && 32 * tid < nq) cw = __ldcg(reinterpret_cast<const unsigned int*>(rec0 + P.gl.off_covq) + (q_lo >> 5) + tid);
        uint4 sp0 = make_uint4(0, 0, 0, 0), sp1 = sp0;
        if (tid < np) {
            const uint4* ps = reinterpret_cast<const uint4*>(rec0 + P.gl.off_persons) + 2 * tid;
            sp0 = __ldcg(ps);
            sp1 = __ldcg(ps + 1);
        }
        __syncthreads();                                                // the initialised mbarrier
        mbar_wait_parity(&s_bar, 0);
#pragma unroll 4
        for (int q = tid; q < nq; q += kSpmUnitThreads) screen_quad(q);
        if (disp && tid < np) s_j[tid] = make_int2((int)max(min(jv.x, 1ll << 30), -(1ll << 30)), (int)max(min(jv.y, 1ll << 30), -(1ll << 30)));
        if (tid < kSpmUnitWords) s_cov[tid] = cw;
        if (tid < np) {
            reinterpret_cast<uint4*>(s_p)[2 * tid] = sp0;
            reinterpret_cast<uint4*>(s_p)[2 * tid + 1] = sp1;
        }
    } else {
        if (disp) {
            const long long* jimg = P.joints + ((long long)img * P.Pmax * P.K + jn) * 2;
            for (int p = tid; p < np; p += kSpmUnitThreads) {
                const longlong2 jv = __ldg(reinterpret_cast<const longlong2*>(jimg + (long long)p * P.K * 2));
                s_j[p] = make_int2((int)max(min(jv.x, 1ll << 30), -(1ll << 30)), (int)max(min(jv.y, 1ll << 30), -(1ll << 30)));
            }
        }
        pdl_wait();                                                     // the image records are complete and visible
    }
    const unsigned char* rec = P.geom + (unsigned long long)img * P.gl.stride;
    const unsigned int* covq_g = reinterpret_cast<const unsigned int*>(rec + P.gl.off_covq);
    const unsigned long long* rowmask_g = reinterpret_cast<const unsigned long long*>(rec + P.gl.off_rowmask);
    const unsigned char* map_g = rec + P.gl.off_map;
    if (!early_a) {
        // coverage words of the unit: word i = quads [32 i, 32 i + 32) of the unit
        if (qpr == 32) {
            for (int i = tid; i < kSpmUnitWords; i += kSpmUnitThreads) s_cov[i] = (32 * i < nq) ? __ldcg(covq_g + (q_lo >> 5) + i) : 0u;
        } else {
#pragma unroll 2
            for (int i = wid; i < kSpmUnitWords; i += kSpmUnitWarps) {
                const int q = q_lo + 32 * i + lane;
                bool covered = false;
                if (32 * i + lane < nq) {
                    const int row = (int)fdiv((uint32_t)q, P.div_qpr), cq = q - row * qpr;
                    covered = (__ldcg(covq_g + row * P.wpr + (cq >> 5)) >> (cq & 31)) & 1u;
                }
                const unsigned w = __ballot_sync(FULL_MASK, covered);
                if (lane == 0) s_cov[i] = w;
            }
        }
        for (int p = tid; p < np; p += kSpmUnitThreads) s_p[p] = reinterpret_cast<const SpmFusedPerson*>(rec + P.gl.off_persons)[p];
    }
    __syncthreads();                                                    // s_cov, s_p, s_j (and, not early_a, the initialised mbarrier)
    // list of the covered quads: every warp scans the word pop-counts itself (no extra barrier; a lane takes WPL consecutive
    // words), then fills the part of the list that belongs to its own words
    int ncov;
    {
        constexpr int WPL = (kSpmUnitWords + 31) / 32;                  // words per lane: 1, 2 or 4
        unsigned w[WPL];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
            w[k] = (lane * WPL + k < kSpmUnitWords) ? s_cov[lane * WPL + k] : 0u;
            cnt += __popc(w[k]);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        ncov = __shfl_sync(FULL_MASK, incl, 31);
        if (ncov) {                                                     // warp-uniform
            // lane l of warp `wid` lists the words l*WPL .. l*WPL+WPL-1 when l % warps == wid (every word exactly once)
            if ((lane % kSpmUnitWarps) == wid) {
                int pos = incl - cnt;
#pragma unroll
                for (int k = 0; k < WPL; ++k) {
                    unsigned bits = w[k];
                    const int q0 = 32 * (lane * WPL + k);
                    while (bits) {
                        const int b = __ffs((int)bits) - 1;
                        bits &= bits - 1u;
                        s_list[pos++] = (unsigned short)(q0 + b);
                    }
                }
            }
        }
    }
    if (ncov) __syncthreads();                                          // the list is complete (CTA-uniform branch)

    // target of covered pixel i of the unit (4 pixels per listed quad): nothing here depends on the logits
    auto target_of = [&](int i, float& t0, float& te, bool& mk) {
        const int qu = (int)s_list[i >> 2], e = i & 3;                  // unit-relative quad, element
        const int q = q_lo + qu;                                        // plane-relative quad
        const int row = (int)fdiv((uint32_t)q, P.div_qpr), col = (q - row * qpr) * 4 + e;
        t0 = 0.0f;
        te = 0.0f;
        unsigned int code = 127u;
        if (P.gl.use_map && disp) code = __ldg(map_g + row * P.R + col);
        if (P.gl.use_map && disp && (code & 127u) != 127u) {
            mk = code >> 7;
            if (code & 127u) {
                const int2 jv = s_j[(int)(code & 127u) - 1];
                if (!(jv.x <= 0 && jv.y <= 0)) {
                    const int dd = axis ? jv.y - row : jv.x - col;
                    te = (float)((P.div_n && dd >= -P.R && dd <= P.R) ? __ldg(P.div_tab + dd + P.R) : (double)dd / P.z);
                }
            }
        } else {
            spm_pixel_target(P, s_p, s_j, __ldg(rowmask_g + row), row, col, disp, axis, t0, te);
            mk = t0 > 0.0f;
        }
        if (!disp) te = t0;                                             // (root plane: the target value itself)
    };
    // loss term + gradient of covered pixel i given its target; stores the results
    auto finish = [&](int i, float t0, float te, bool mk) {
        const int qu = (int)s_list[i >> 2], e = i & 3;
        float ge = 0.0f;
        if (LOSS) {
            const float pe = tile[qu * 4 + e];
            if (!disp) {
                const float sg = sigmoid_fast(pe);
                const float d = (mk ? sg : sg * 0.0f) - t0;
                acc = fmaf(d, d, acc);
                ge = mk ? P.groot * d * ((1.0f - sg) * sg) : 0.0f;
            } else {
                // tanh only where the root mask is set; elsewhere tanh(p)*0 == 0 for every finite or infinite p
                float th = 0.0f, pm = pe != pe ? pe : 0.0f;          // NaN logits propagate as in the reference
                if (mk) { th = tanhf(pe); pm = th; }
                const float d = pm - te;
                const float ad = fabsf(d);
                acc += ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
                ge = mk ? P.gdisp * fminf(fmaxf(d, -1.0f), 1.0f) * (1.0f - th * th) : 0.0f;
            }
        }
        const long long ei = (off + qu) * 4 + e;
        if (GRAD) __stcs(P.dlogits + ei, ge);
        if (WTGT) __stcs(P.target_out + ei, te);
    };
    // render-only form: the targets of this thread's first covered pixels are requested BEFORE the zero stores of phase A, so
    // their dependent chain (geometry byte from L2 -> joint -> quotient table) runs under the store stream: 100 vs 108 us per
    // 256 images.  (With logits to wait for the same trick gains nothing: 123 vs 119 us read-only.)
    constexpr int NPRE = LOSS ? 0 : 2;
    float pt0[NPRE + 1], pte[NPRE + 1];
    bool pmk[NPRE + 1];
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
        const int i = tid + k * kSpmUnitThreads;
        pt0[k] = 0.0f; pte[k] = 0.0f; pmk[k] = false;
        if (i < 4 * ncov) target_of(i, pt0[k], pte[k], pmk[k]);
    }
    float4* G4 = reinterpret_cast<float4*>(P.dlogits) + off;
    float4* T4 = reinterpret_cast<float4*>(P.target_out) + off;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (LOSS && !early_a) mbar_wait_parity(&s_bar, 0);
    // phase A: the quads no person touches
    // (read-only form: the screen runs over EVERY quad, covered or not -- a NaN under a person makes the sum NaN in phase B anyway,
    // every other covered logit adds 0 here -- so the stream carries no coverage test: SCREEN_ALL; done above when early_a)
    if (!early_a) {
#pragma unroll 4
        for (int q = tid; q < nq; q += kSpmUnitThreads) {
            if (!SCREEN_ALL && ((s_cov[q >> 5] >> (q & 31)) & 1u)) continue;
            if (LOSS) screen_quad(q);
            if (GRAD) __stcs(G4 + q, z4);
            if (WTGT) __stcs(T4 + q, z4);
        }
    }
    // phase B: one pixel of a covered quad per thread.  (Measured and dropped: working the targets out before the wait for the
    // bulk copy -- 123 vs 119 us read-only; a per-image hash table pixel -> covering persons for the pixels in several boxes
    // instead of the person walk -- 131 vs 123 us; profiles/r02_tune_spm_experiments.log.)
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
        const int i = tid + k * kSpmUnitThreads;
        if (i < 4 * ncov) finish(i, pt0[k], pte[k], pmk[k]);
    }
    for (int i = tid + NPRE * kSpmUnitThreads; i < 4 * ncov; i += kSpmUnitThreads) {
        float t0, te;
        bool mk;
        target_of(i, t0, te, mk);
        finish(i, t0, te, mk);
    }
    if (!LOSS) return;                                                   // render-only: no loss partials
    acc = warp_sum(acc);
    if (lane == 0) s_acc[wid] = acc;
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < kSpmUnitWarps; ++w) a += (double)s_acc[w];
        reinterpret_cast<double2*>(P.partials)[unit] = disp ? make_double2(0.0, a) : make_double2(a, 0.0);
    }
}

// two-level deterministic reduction of the per-unit (S_root, S_disp) pairs (common.cuh); grid = R slice CTAs
__global__ void __launch_bounds__(256) spm_loss_reduce_kernel(const double* __restrict__ pairs, long long n, double* __restrict__ slices,
                                                              unsigned int* __restrict__ ticket, int R, double w0, double w1, double inv_norm,
                                                              float* __restrict__ loss_out, double* __restrict__ num_out) {
    pdl_wait();
    if (reduce_slice_and_elect(pairs, n, slices, ticket, R, (int)blockIdx.x)) reduce_pairs_cta(slices, R, 2, w0, w1, inv_norm, loss_out, num_out);
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
    int sig_ref;                 // apply_act: which torch.sigmoid gives the root confidences (kSigmoidAtenCpu / kSigmoidAtenCuda)
    float x_lo;                  // apply_act: logits <= x_lo cannot reach sigmoid > thr under either reference (host: logit(thr) minus a margin)
    float zf;                    // fp32(sqrt(2 R^2))
    float input_size;            // DecodeSPM.input_size
    long long s_min;             // smallest integer s with sqrt((double)s) > dist_thr: the radius test on integer offsets
    float q_lt;                  // smallest fp32 q with sqrt((double)q) >= dist_thr: the joint test `sqrt_fp64(q) < dist_thr` is q < q_lt
};

// One body joint of one root (get_spm_keypoints utils/spm_utils.py:187-197 + the rescale at :247-248):
//   kx = disp[2k][y,x] * fp32(z) + x   (fp32 multiply, then fp32 add: two roundings, no FMA)
//   d  = sqrt_fp64( fp32((x-kx)^2 + (y-ky)^2) );  d < dist_thr -> (0,0,0) else (kx, ky, conf) * num / den
// (sqrt_fp64 of an fp32 is monotone, so the host turns `sqrt((double)q) < dist_thr` into the fp32 bound q_lt: no DSQRT per joint)
__device__ __forceinline__ void spm_joint(const float* __restrict__ disp, long long plane, int pix, int rx, int ry, float conf,
                                          int k, int apply_act, float zf, float q_lt, float num, float den, float* o) {
    float dx = __ldg(disp + (2 * k) * plane + pix);
    float dy = __ldg(disp + (2 * k + 1) * plane + pix);
    if (apply_act) { dx = tanhf(dx); dy = tanhf(dy); }
    const float fx = (float)rx, fy = (float)ry;
    const float kx = __fadd_rn(__fmul_rn(dx, zf), fx);
    const float ky = __fadd_rn(__fmul_rn(dy, zf), fy);
    const float ex = __fsub_rn(fx, kx), ey = __fsub_rn(fy, ky);
    const float q = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    if (q < q_lt) {
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

// Hierarchical displacement chaining (NOT in the reference, which reads every joint at the root pixel -- single hop,
// utils/spm_utils.py:187-189; opt-in, parity unpinned; the hierarchical SPR decode of the SPM paper): joint k hangs off
// parent[k] (-1 = the root); its displacement is read at the PARENT's decoded position (truncated to a pixel, like the
// reference's .long() on the root) and added to that position with the reference's arithmetic (fp32 multiply by fp32(z), fp32
// add; fp64 sqrt of the fp32 squared distance to the parent, `< dist_thr` -> the joint is absent).  An absent or off-map parent
// makes every descendant absent (0,0,0).  One thread per (root, joint) walks its own ancestor path from the root down, so
// `parent` needs no particular order; with parent[k] = -1 for every k the result is bit-identical to spm_gather_kernel.
constexpr int kSpmChainDepth = 16;
__global__ void __launch_bounds__(128) spm_gather_chain_kernel(const float* __restrict__ roots, const float* __restrict__ disp,
                                                               const int* __restrict__ parent, float* __restrict__ kps, int n, int K, int R,
                                                               float zf, double dist_thr) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * K) return;
    const int r = t / K, k = t - r * K;
    float* o = kps + 3 * t;
    int path[kSpmChainDepth];
    int depth = 0;
    for (int j = k; j >= 0 && depth < kSpmChainDepth; j = __ldg(parent + j)) path[depth++] = j;
    float px = roots[3 * r], py = roots[3 * r + 1];
    const float c = roots[3 * r + 2];
    bool ok = depth < kSpmChainDepth || __ldg(parent + path[depth - 1]) < 0;       // a longer chain (or a cycle) is rejected
    for (int i = depth - 1; i >= 0 && ok; --i) {
        const int j = path[i];
        const int xi = (int)px, yi = (int)py;
        if (xi < 0 || xi >= R || yi < 0 || yi >= R) { ok = false; break; }
        const float dx = __ldg(disp + (long long)(2 * j) * R * R + yi * R + xi);
        const float dy = __ldg(disp + (long long)(2 * j + 1) * R * R + yi * R + xi);
        const float kx = __fadd_rn(__fmul_rn(dx, zf), px), ky = __fadd_rn(__fmul_rn(dy, zf), py);
        const float ex = __fsub_rn(px, kx), ey = __fsub_rn(py, ky);
        const float q = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
        if (sqrt((double)q) < dist_thr) { ok = false; break; }
        px = kx; py = ky;
    }
    if (ok) { o[0] = px; o[1] = py; o[2] = c; }
    else { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
}

// Tail of SPMmAPCOCO.update_state (utils/spm_utils.py:302-304): x *= img_w / input_size, y *= img_h / input_size -- the ratio is
// an integer tensor divided by a python int, i.e. one fp32 division, then an in-place fp32 multiply.  Rows >= counts[i] are
// not touched (they were never written by the decode kernel).
__global__ void __launch_bounds__(256) spm_rescale_kernel(const float* __restrict__ kps, const int* __restrict__ counts,
                                                          const long long* __restrict__ image_w, const long long* __restrict__ image_h,
                                                          float* __restrict__ out, int N, int Pmax, int K, float input_size) {
    const long long per_img = (long long)Pmax * K;
    const long long total = (long long)N * per_img;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int img = (int)(t / per_img);
        const int p = (int)((t - (long long)img * per_img) / K);
        if (p >= counts[img]) continue;
        const float rx = __fdiv_rn((float)image_w[img], input_size), ry = __fdiv_rn((float)image_h[img], input_size);
        out[3 * t] = __fmul_rn(kps[3 * t], rx);
        out[3 * t + 1] = __fmul_rn(kps[3 * t + 1], ry);
        out[3 * t + 2] = kps[3 * t + 2];
    }
}

// Root NMS + joint gather of one image per CTA (128 threads; ~9 CTAs per SM, so 1024 images are one wave).
//  1. The root plane is streamed ONCE (128-bit loads, 8 in flight per lane); pixels with conf > thr are appended to a
//     candidate list in shared memory (packed (y, x), value) with one shared atomic each -- real maps have a few dozen.
//  2. Greedy NMS on the LIST: best remaining candidate (value desc, row-major index asc), then every listed candidate
//     within the radius is struck out.  <= kSpmWarpList candidates: one warp runs the whole loop on shuffles, no block
//     barrier.  Picks are only RECORDED here.
//  3. The joints of all recorded roots are gathered by the whole CTA with every load in flight at once (2K scattered
//     loads per root, ~1 us of DRAM latency each if done inside the loop).
//  Dense maps (> kSpmCandCap candidates; adversarial) keep no list: every iteration re-reads the plane through L2 and a
//  suppressed-pixel bitmap in shared memory replaces the strike-out.  Slow, bounded, same picks.
// History: v1 kept the whole activated map in shared memory (64 KB: 3 CTAs per SM), scanned it with the CTA for every
// root (3 barriers per root) and gathered joints inside the loop: 58 us per 256 images; this version 13 us.
// CTA size: 128 threads when the launch has enough images to fill the GPU (1024 images = one wave of 7 CTAs per SM); with fewer
// images than ~4 per SM more threads per image stream its plane faster (N=256: 16.6 us at 128, 14.4 at 256, 12.5 at 512;
// N=1024: 22.7 / 24.9 / 27.6 -- profiles/r02_tune_spm_decode.log).  Loads in flight per thread: 8 at 128 threads, 4 above.
constexpr int kSpmDecThreadsMax = 512;
constexpr int kSpmCandCap = 2048;
constexpr int kSpmWarpList = 256;
constexpr int kSpmRootCap = 256;              // picked roots recorded in shared memory before their joints are gathered

template <int kSpmDecThreads>
__global__ void __launch_bounds__(kSpmDecThreads) spm_decode_kernel(SpmDecodeParams P) {
    extern __shared__ __align__(16) unsigned int sup_bits[];   // dense fallback only: R*R bits, 1 = suppressed
    __shared__ unsigned int s_cand[kSpmCandCap];               // (y << 16) | x  == row-major order for ties
    __shared__ __align__(8) float s_val[kSpmCandCap];           // activated confidence, -inf once struck out
    __shared__ int s_root_i[kSpmRootCap];
    __shared__ float s_root_c[kSpmRootCap];
    __shared__ float s_v[kSpmDecThreads / 32];
    __shared__ unsigned int s_i[kSpmDecThreads / 32];
    __shared__ float s_best;
    __shared__ unsigned int s_besti;
    __shared__ int s_ncand, s_found;
    const int img = blockIdx.x;
    const int RR = P.R * P.R;
    const long long plane = RR;
    const float* base = P.x + (long long)img * P.C * plane;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = kSpmDecThreads / 32;
    if (threadIdx.x == 0) s_ncand = 0;
    __syncthreads();

    // root confidence = the reference's sigmoid, bit for bit (the greedy order and the threshold test depend on it).  The stream
    // only compares the raw value with a bound (apply_act: x_lo, the logit below which neither reference can exceed thr; else thr
    // itself) and lists what passes; the reference sigmoid is evaluated afterwards, on the list -- a few dozen entries of a real
    // map, all lanes busy -- so the streaming loop holds no SFU work, no divisions and no divergent branch.  (Evaluating it
    // inside the stream behind `v > x_lo` cost 8-10 us per launch: 18.6 -> 26.7 us per 256 images.)
    const float pass_lo = P.apply_act ? P.x_lo : P.thr;
    auto conf_of = [&](float v) -> float {                     // dense fallback only
        if (!P.apply_act) return v;
        return v > P.x_lo ? sigmoid_ref(v, P.sig_ref) : -INFINITY;
    };
    auto append = [&](int i, float h) {
        const int pos = atomicAdd(&s_ncand, 1);
        if (pos < kSpmCandCap) {
            const int y = i / P.R;
            s_cand[pos] = ((unsigned)y << 16) | (unsigned)(i - y * P.R);
            s_val[pos] = h;
        }
    };
    if ((RR & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0) {
        const float4* b4 = reinterpret_cast<const float4*>(base);
        const int nq = RR >> 2;
        constexpr int U = kSpmDecThreads == 128 ? 8 : 4;
        for (int q0 = threadIdx.x; q0 < nq; q0 += kSpmDecThreads * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = q0 + u * kSpmDecThreads;
                if (q < nq) v[u] = ldg_stream(b4 + q);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = q0 + u * kSpmDecThreads;
                if (q >= nq) break;
                const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (e[k] > pass_lo) append(4 * q + k, e[k]);
            }
        }
    } else {
        for (int i = threadIdx.x; i < RR; i += blockDim.x) {
            const float v = ldg_stream(base + i);
            if (v > pass_lo) append(i, v);
        }
    }
    __syncthreads();
    const int ncand = s_ncand;
    if (P.apply_act && ncand <= kSpmCandCap) {
        // listed logits -> reference confidences; entries that do not pass the threshold after all are struck out
        for (int e = threadIdx.x; e < ncand; e += blockDim.x) {
            const float h = sigmoid_ref(s_val[e], P.sig_ref);
            s_val[e] = h > P.thr ? h : -INFINITY;
        }
        __syncthreads();
    }

    int found = 0, flushed = 0;
    float* roots = P.roots + (long long)img * P.Pmax * 3;
    float* kps = P.kps + (long long)img * P.Pmax * P.K * 3;

    // roots [flushed, upto) -> global rows + joints, by the whole CTA (callers synchronise around it)
    auto flush = [&](int upto) {
        const int n = min(upto, P.Pmax) - flushed;
        for (int t = threadIdx.x; t < n * (P.K + 1); t += blockDim.x) {
            const int r = t / (P.K + 1), k = t - r * (P.K + 1) - 1;
            const int ry = s_root_i[r] >> 16, rx = s_root_i[r] & 0xffff;
            const float best = s_root_c[r];
            const int slot = flushed + r;
            if (k < 0) {
                roots[slot * 3 + 0] = __fdiv_rn(__fmul_rn((float)rx, P.input_size), (float)P.R);
                roots[slot * 3 + 1] = __fdiv_rn(__fmul_rn((float)ry, P.input_size), (float)P.R);
                roots[slot * 3 + 2] = best;
            } else {
                spm_joint(base + plane, plane, ry * P.R + rx, rx, ry, best, k, P.apply_act, P.zf, P.q_lt, P.input_size, (float)P.R,
                          kps + ((long long)slot * P.K + k) * 3);
            }
        }
    };
    // best remaining listed candidate among entries t0, t0+stride, ...: larger value, then smaller (y, x) key
    auto pick_listed = [&](int t0, int stride, float& best, unsigned& bkey) {
        best = -INFINITY;
        bkey = 0xffffffffu;
        for (int e = t0; e < ncand; e += stride) {
            const float v = s_val[e];
            const unsigned key = s_cand[e];
            if (v > best || (v == best && key < bkey)) { best = v; bkey = key; }
        }
    };
    auto warp_best = [&](float& v, unsigned& key) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL_MASK, v, o);
            const unsigned ok = __shfl_xor_sync(FULL_MASK, key, o);
            if (ov > v || (ov == v && ok < key)) { v = ov; key = ok; }
        }
    };
    // survivors satisfy sqrt(dx^2+dy^2) > dist_thr  <=>  dx^2+dy^2 >= s_min (s_min found on the host with the same sqrt)
    auto strike_listed = [&](unsigned bkey, int t0, int stride) {
        const int by = (int)(bkey >> 16), bx = (int)(bkey & 0xffffu);
        for (int e = t0; e < ncand; e += stride) {
            const unsigned key = s_cand[e];
            const long long dy = (int)(key >> 16) - by, dx = (int)(key & 0xffffu) - bx;
            if (dx * dx + dy * dy < P.s_min) s_val[e] = -INFINITY;
        }
    };

    if (ncand <= 64 && P.R <= 32768) {
        // The usual case (a few dozen candidates; config 4: 23 on average, 40 at most): the greedy loop "best remaining, strike its
        // neighbourhood" costs ~1000 cycles per ROOT as a chain of warp reductions (measured with clock64 stamps: 5 700 cycles per
        // image, 11 600 for an 8-person image -- more than the 64 KB stream of the plane at N = 256).  Here one warp does the same
        // selection in closed form, two candidates per lane:
        //   1. rank of every candidate under (value desc, row-major index asc) by comparison with all others -- ONE unsigned 64-bit
        //      comparison of (ordered value bits, ~index) keys per pair (broadcast LDS.64);
        //   2. candidates re-read in rank order; for each, the 64-bit mask of the ranks it would strike (the same integer radius
        //      test, in 32 bits: R <= 32768 keeps dx^2 + dy^2 below 2^31);
        //   3. the greedy walk over a warp-uniform `alive` mask: lowest set bit = next root, alive &= ~its mask -- four shuffles and
        //      three logic operations per root, no reduction.
        // Same picks in the same order as the loop below.  A single warp runs this, so its time is its instruction count times the
        // issue latency: the 64-bit keys and 32-bit distances took the two loops from 12 + 30 to 7 + 14 instructions per candidate.
        if (wid == 0) {
            unsigned int* s_skey = reinterpret_cast<unsigned int*>(s_root_i) + kSpmRootCap / 2;      // scratch: upper half of the root record
            float* s_sval = s_root_c + kSpmRootCap / 2;                                             // (a warp-list image records <= 64 roots)
            unsigned long long* s_k64 = reinterpret_cast<unsigned long long*>(s_val + kSpmCandCap / 2);   // scratch: 64 keys behind the <= 64 listed values
            const int n = ncand;
            const int i0 = lane, i1 = lane + 32;
            const float v0 = i0 < n ? s_val[i0] : -INFINITY, v1 = i1 < n ? s_val[i1] : -INFINITY;
            const unsigned k0 = i0 < n ? s_cand[i0] : 0xffffffffu, k1 = i1 < n ? s_cand[i1] : 0xffffffffu;
            // larger key = better candidate: value first ((v + 0) folds -0 into +0; struck-out entries are -inf), then the SMALLER (y, x)
            const unsigned long long key0 = ((unsigned long long)float_key(v0 + 0.0f) << 32) | (unsigned long long)(0xffffffffu - k0);
            const unsigned long long key1 = ((unsigned long long)float_key(v1 + 0.0f) << 32) | (unsigned long long)(0xffffffffu - k1);
            if (i0 < n) s_k64[i0] = key0;
            if (i1 < n) s_k64[i1] = key1;
            __syncwarp();
            int r0 = 0, r1 = 0;
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
                const unsigned long long ki = s_k64[i];
                r0 += ki > key0 ? 1 : 0;
                r1 += ki > key1 ? 1 : 0;
            }
            if (i0 < n) { s_skey[r0] = k0; s_sval[r0] = v0; }
            if (i1 < n) { s_skey[r1] = k1; s_sval[r1] = v1; }
            __syncwarp();
            const unsigned sk0 = i0 < n ? s_skey[i0] : 0u, sk1 = i1 < n ? s_skey[i1] : 0u;         // lane owns ranks `lane` and `lane + 32`
            const float sv0 = i0 < n ? s_sval[i0] : -INFINITY, sv1 = i1 < n ? s_sval[i1] : -INFINITY;
            const int y0 = (int)(sk0 >> 16), x0 = (int)(sk0 & 0xffffu), y1 = (int)(sk1 >> 16), x1 = (int)(sk1 & 0xffffu);
            const unsigned smin = (unsigned)min(P.s_min, 0xffffffffll);
            unsigned m0lo = 0u, m0hi = 0u, m1lo = 0u, m1hi = 0u;
            auto strike = [&](int q, unsigned& ma, unsigned& mb) {
                const unsigned kq = s_skey[q];
                const int yq = (int)(kq >> 16), xq = (int)(kq & 0xffffu);
                const unsigned a = (unsigned)((yq - y0) * (yq - y0) + (xq - x0) * (xq - x0));
                const unsigned b = (unsigned)((yq - y1) * (yq - y1) + (xq - x1) * (xq - x1));
                const unsigned bit = 1u << (q & 31);
                if (a < smin) ma |= bit;
                if (b < smin) mb |= bit;
            };
            const int nlo = min(n, 32);
#pragma unroll 4
            for (int q = 0; q < nlo; ++q) strike(q, m0lo, m1lo);
#pragma unroll 4
            for (int q = 32; q < n; ++q) strike(q, m0hi, m1hi);
            __syncwarp();                                                   // the scratch is read; the record may be written
            unsigned long long alive = (unsigned long long)__ballot_sync(FULL_MASK, sv0 > -INFINITY) |
                                       ((unsigned long long)__ballot_sync(FULL_MASK, sv1 > -INFINITY) << 32);
            while (alive) {                                                 // warp-uniform
                const int p = __ffsll((long long)alive) - 1, src = p & 31;
                const bool hi = p >= 32;
                const unsigned key = __shfl_sync(FULL_MASK, hi ? sk1 : sk0, src);
                const float val = __shfl_sync(FULL_MASK, hi ? sv1 : sv0, src);
                const unsigned mlo = __shfl_sync(FULL_MASK, hi ? m1lo : m0lo, src), mhi = __shfl_sync(FULL_MASK, hi ? m1hi : m0hi, src);
                if (lane == 0 && found < P.Pmax) { s_root_i[found] = (int)key; s_root_c[found] = val; }
                alive &= ~(((unsigned long long)mhi << 32) | (unsigned long long)mlo);
                alive &= ~(1ull << p);
                ++found;
            }
            if (lane == 0) s_found = found;
        }
        __syncthreads();
        found = s_found;
        flush(found);
    } else if (ncand <= kSpmWarpList) {
        // one warp, no block barriers; at most ncand <= kSpmRootCap roots, so the record never overflows
        if (wid == 0) {
            while (true) {
                float best; unsigned bkey;
                pick_listed(lane, 32, best, bkey);
                warp_best(best, bkey);
                if (!(best > -INFINITY)) break;
                if (lane == 0 && found < P.Pmax) { s_root_i[found] = (int)bkey; s_root_c[found] = best; }
                strike_listed(bkey, lane, 32);
                ++found;
                __syncwarp();
            }
            if (lane == 0) s_found = found;
        }
        __syncthreads();
        found = s_found;
        flush(found);
    } else {
        const bool list = ncand <= kSpmCandCap;
        const int rad = (int)floor(P.dist_thr);
        const int side = 2 * rad + 1;
        if (!list) {
            for (int i = threadIdx.x; i < (RR + 31) / 32; i += blockDim.x) sup_bits[i] = 0u;
            __syncthreads();
        }
        while (true) {
            float best; unsigned bkey;
            if (list) {
                pick_listed((int)threadIdx.x, kSpmDecThreads, best, bkey);
            } else {
                // dense map: re-read the plane (L2), skip suppressed pixels; ascending index per thread keeps the first of equals
                best = -INFINITY;
                bkey = 0xffffffffu;
                for (int i = threadIdx.x; i < RR; i += blockDim.x) {
                    if ((sup_bits[i >> 5] >> (i & 31)) & 1u) continue;
                    const float h = conf_of(__ldg(base + i));
                    if (h > P.thr && h > best) { best = h; const int y = i / P.R; bkey = ((unsigned)y << 16) | (unsigned)(i - y * P.R); }
                }
            }
            warp_best(best, bkey);
            if (lane == 0) { s_v[wid] = best; s_i[wid] = bkey; }
            __syncthreads();
            if (wid == 0) {
                best = lane < NW ? s_v[lane] : -INFINITY;
                bkey = lane < NW ? s_i[lane] : 0xffffffffu;
                warp_best(best, bkey);
                if (lane == 0) { s_best = best; s_besti = bkey; }
            }
            __syncthreads();
            best = s_best; bkey = s_besti;
            if (!(best > -INFINITY)) break;
            if (threadIdx.x == 0 && found < P.Pmax) { s_root_i[found - flushed] = (int)bkey; s_root_c[found - flushed] = best; }
            if (list) {
                strike_listed(bkey, (int)threadIdx.x, kSpmDecThreads);
            } else {
                const int ry = (int)(bkey >> 16), rx = (int)(bkey & 0xffffu);
                for (int t = threadIdx.x; t < side * side; t += blockDim.x) {
                    const int oy = t / side - rad, ox = t - (t / side) * side - rad;
                    const int y = ry + oy, x = rx + ox;
                    if (y < 0 || y >= P.R || x < 0 || x >= P.R) continue;
                    if ((long long)(ox * ox + oy * oy) < P.s_min) atomicOr(&sup_bits[(y * P.R + x) >> 5], 1u << ((y * P.R + x) & 31));
                }
            }
            ++found;
            __syncthreads();
            if (found - flushed == kSpmRootCap) {                        // CTA-uniform: the record is full
                flush(found);
                flushed = found;
                __syncthreads();
            }
        }
        flush(found);
    }
    if (threadIdx.x == 0) {
        P.counts[img] = min(found, P.Pmax);
        if (P.counts_total) P.counts_total[img] = found;
    }
}

}  // namespace pose
