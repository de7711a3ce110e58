// Simple-Baselines hot path kernels: render, fused render+loss+grad(+decode), decode, back-projection.
//
// Work unit = one heat map (H*W fp32, 12 288 B at 64x48).  One warp owns one map at a time and a
// persistent grid (SM count x resident CTAs) strides over all N*K maps, so every global access is
// a fully coalesced 128-bit load/store of 512 contiguous bytes per warp instruction, the per-map
// reductions (argmax, loss partials) are warp shuffles only, and no __syncthreads() appears in
// the streaming loop.  All stages are HBM-bound; nothing here is a contraction (no tensor cores).
#pragma once
#include "common.cuh"

namespace pose {

#ifndef POSE_FUSED_U
#define POSE_FUSED_U 6          // independent 128-bit loads in flight per lane in the fused kernel (tuned: tools/tune_fused.py)
#endif
#ifndef POSE_FUSED_MINB
#define POSE_FUSED_MINB 3       // resident CTAs per SM the fused kernel is compiled for (register cap 80)
#endif
// The read-only render variants (validation: loss and/or decode from keypoints, no dlogits) have no store stream to carry half of the traffic:
// they need more loads in flight per SM to hide the same latency, so they get their own knobs.
// (tools/tune_fused.py --which ng, B=4096, loss only: U6/M4 144.9 us, U8/M4 143.5, U4/M5 144.7, U6/M3 150.8, U12/M3 148.3.)
#ifndef POSE_FUSED_U_NG
#define POSE_FUSED_U_NG 8
#endif
#ifndef POSE_FUSED_MINB_NG
#define POSE_FUSED_MINB_NG 4    // register cap 64
#endif
// ... and the read-only variant that also decodes (running argmax: more registers per element) its own again
// (tools/tune_fused.py --which ngd: U8/M3 158.3 us, U4/M4 158.5, U6/M3 163.8, U6/M4 169.8, U8/M2 182.7.  Tracking the maximum per
//  128-bit vector and resolving the element once per map with a 16-byte re-read was measured too: 2 % slower in every shape.)
#ifndef POSE_FUSED_U_NGD
#define POSE_FUSED_U_NGD 8
#endif
#ifndef POSE_FUSED_MINB_NGD
#define POSE_FUSED_MINB_NGD 3   // register cap 80
#endif
constexpr int kSbpThreads = 256;               // 8 warps per CTA
constexpr int kSbpWarps = kSbpThreads / 32;
constexpr int kMaxPartialBlocks = 148 * 16;    // upper bound on the persistent grid (workspace sizing)

// ---------------------------------------------------------------- vector helpers
template <int V> struct Vec;
template <> struct Vec<4> {
    using T = float4;
    static __device__ __forceinline__ void load(const float* base, int vi, float (&o)[4]) {
        float4 v = ldg_stream(reinterpret_cast<const float4*>(base) + vi);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    static __device__ __forceinline__ void load_cached(const float* base, int vi, float (&o)[4]) {
        float4 v = ldg_cached(reinterpret_cast<const float4*>(base) + vi);
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
    }
    static __device__ __forceinline__ void store(float* base, int vi, const float (&o)[4]) {
        __stcs(reinterpret_cast<float4*>(base) + vi, make_float4(o[0], o[1], o[2], o[3]));
    }
};
template <> struct Vec<1> {
    using T = float;
    static __device__ __forceinline__ void load(const float* base, int vi, float (&o)[1]) { o[0] = ldg_stream(base + vi); }
    static __device__ __forceinline__ void load_cached(const float* base, int vi, float (&o)[1]) { o[0] = ldg_cached(base + vi); }
    static __device__ __forceinline__ void store(float* base, int vi, const float (&o)[1]) { __stcs(base + vi, o[0]); }
};

// ---------------------------------------------------------------- Gaussian patch geometry of one map
// SBPHeatmapGenerator.__call__ (utils/sbp_utils.py:36-51): skip if x<0 or y<0; centre = clip(int(x));
// corners = round-half-even(c - 3s - 1), round-half-even(c + 3s + 2) evaluated in double exactly as
// numpy does ((c - 3s) - 1); destination window = corners clipped to the map.
struct Patch {
    int ulx, uly;            // template origin in map coordinates (may be negative)
    int px0, px1, py0, py1;  // clipped destination window, empty when the joint is invisible
    // flat element range [e0, e0 + ecnt) of the rows the window touches (ecnt == 0: nothing): the cheap per-vector test
    int e0, ecnt;
    int nrows;               // py1 - uly (0 when empty): template rows [0, nrows) minus the ones above the map are rendered
};

// Shared-memory template of the fused kernels: n+1 rows of n+8 floats -- 4 zeros, the n template values, 4 zeros -- the last row all
// zeros, so a lane fetches the targets of 4 consecutive pixels of one row with two clamped indices instead of 4 x (row test,
// column test) behind divergent branches.
constexpr int kLutPad = 4;
__host__ __device__ inline size_t lut_padded_floats(int n) { return (size_t)(n + 1) * (size_t)(n + 2 * kLutPad); }
__device__ __forceinline__ void stage_lut_padded(float* __restrict__ lut_s, const float* __restrict__ lut, int n) {
    const int pw = n + 2 * kLutPad;
    for (int i = threadIdx.x; i < (n + 1) * pw; i += blockDim.x) {
        const int r = i / pw, c = i - r * pw - kLutPad;
        lut_s[i] = (r < n && c >= 0 && c < n) ? lut[r * n + c] : 0.0f;
    }
}

__device__ __forceinline__ Patch make_patch(double x, double y, int H, int W, double three_sigma, int lut_n) {
    Patch p;
    const bool visible = (x >= 0.0) && (y >= 0.0);   // == not (x<0 or y<0) for non-NaN input; -0.0 is visible
    // x >= 0 here, so clip(int(x), 0, W-1) == int(min(x, W-1)) -- also safe for huge / inf inputs
    const int cx = (int)fmin(x, (double)(W - 1));
    const int cy = (int)fmin(y, (double)(H - 1));
    p.ulx = (int)rint(((double)cx - three_sigma) - 1.0);
    p.uly = (int)rint(((double)cy - three_sigma) - 1.0);
    const int brx = (int)rint(((double)cx + three_sigma) + 2.0);
    const int bry = (int)rint(((double)cy + three_sigma) + 2.0);
    p.px0 = max(0, p.ulx);
    p.py0 = max(0, p.uly);
    p.px1 = min(min(brx, W), p.ulx + lut_n);
    p.py1 = min(min(bry, H), p.uly + lut_n);
    if (!visible) { p.px0 = p.px1 = p.py0 = p.py1 = 0; p.ulx = p.uly = 0; }
    p.e0 = p.py0 * W;
    p.ecnt = (p.py1 > p.py0 && p.px1 > p.px0) ? (p.py1 - p.py0) * W : 0;
    p.nrows = p.ecnt ? p.py1 - p.uly : 0;
    return p;
}

__device__ __forceinline__ void load_kp(const void* kp, int kp_f64, long long map, double& x, double& y) {
    if (kp_f64) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(kp) + map);
        x = v.x; y = v.y;
    } else {
        const float2 v = __ldg(reinterpret_cast<const float2*>(kp) + map);
        x = (double)v.x; y = (double)v.y;
    }
}

// target values of V consecutive elements starting at flat index e0 of a map
template <int V>
__device__ __forceinline__ bool patch_values(const Patch& p, const float* __restrict__ lut_s, int lut_n, int e0,
                                             int W, FastDiv divW, float (&t)[V]) {
    int row = (int)fdiv((uint32_t)e0, divW);
    int col = e0 - row * W;
#pragma unroll
    for (int j = 0; j < V; ++j) t[j] = 0.0f;
    // whole vector outside the patch rows: the overwhelmingly common case (zero target)
    const int row_last = row + ((col + V - 1) >= W ? 1 : 0);   // V <= 4 <= W on the vector path: at most one wrap
    if (row_last < p.py0 || row >= p.py1) return false;
    bool any = false;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        if (row >= p.py0 && row < p.py1 && col >= p.px0 && col < p.px1) {
            t[j] = lut_s[(row - p.uly) * lut_n + (col - p.ulx)];
            any = true;
        }
        if (++col >= W) { col = 0; ++row; }
    }
    return any;
}

// ---------------------------------------------------------------- argmax of sigmoid(x): search in logit space, rank with the reference's sigmoid
// nms_sbp (utils/sbp_utils.py:71-80) takes the first row-major index of the largest sigmoid VALUE, and fp32 sigmoid is
// many-to-one, so which neighbours tie depends on the sigmoid implementation to the last bit (common.cuh).  The streaming
// loop therefore never evaluates a sigmoid for the decode: each lane tracks, per 128-bit vector, the largest logit of its
// share of the map (value, vector index, the vector's 4 values) and the second-largest vector maximum; at the end of the map
// the warp derives the candidate window [lo, m] from the map maximum m (sigmoid_window_lo: nothing below lo can reach
// sigmoid_ref(m)), and ranks the candidates -- almost always exactly one element -- with the reference's own sigmoid
// (sigmoid_ref, bit-exact, also the reported confidence).  Only when some lane holds two candidate vectors (it kept one) or
// the map is degenerate does the warp look at the map a second time (L2 hit).  2 FMNMX3/FMNMX + 1 FSETP + 6 predicated
// moves per vector, no MUFU.
template <int V>
struct ArgTrack {
    float best, second;     // largest / second-largest vector maximum seen by this lane (NaNs ignored)
    int bestvi;             // vector index of `best` (first occurrence)
    float keep[V];          // the elements of that vector
    __device__ __forceinline__ void reset() {
        best = second = -INFINITY; bestvi = 0;
#pragma unroll
        for (int j = 0; j < V; ++j) keep[j] = -INFINITY;
    }
    template <bool SIG>
    __device__ __forceinline__ void push(const float (&x)[V], int vi) {
        float vm = x[0];
#pragma unroll
        for (int j = 1; j < V; ++j) vm = fmaxf(vm, x[j]);
        if (SIG) second = fmaxf(second, fminf(vm, best));
        if (vm > best) {
            best = vm; bestvi = vi;
#pragma unroll
            for (int j = 0; j < V; ++j) keep[j] = x[j];
        }
    }
};

// -> (conf, idx) in the lane(s) named by the return value: the reference's activation value at its argmax and the flat index
// (0x7fffffff: nothing comparable in the map).  Returns the lane that holds the result when exactly one lane holds exactly one
// candidate (the overwhelmingly common case: no shuffle reduction, the winner evaluates the reference sigmoid once on its own),
// else 0 after a warp reduction (every lane then holds the result).
// SIG == false (heat maps that are already activated, DecodeSBP.pred == False): candidates are the elements equal to m.
template <int V, bool SIG>
__device__ __forceinline__ int resolve_argmax(const ArgTrack<V>& a, const float* __restrict__ src, int nvec, int lane, int sig_ref,
                                              float& conf, int& idx) {
    const float m = warp_max(a.best);
    float lo = m;
    bool again = false;
    if (SIG) {
        lo = sigmoid_window_lo(m);
        // m <= -80 (or no finite element at all): the references' results are denormal / zero there and the error model of
        // the window does not hold -- rank every element
        const bool degenerate = !(m > -80.0f);
        if (degenerate) lo = -INFINITY;
        again = degenerate || __any_sync(FULL_MASK, a.second >= lo);
    }
    float fb = -INFINITY;
    int fi = 0x7fffffff;
    if (!again) {
        // candidates of this lane: elements of its best vector inside the window
        int ncand = 0, jc = 0;
        float xc = 0.0f;
        if (a.best >= lo) {
#pragma unroll
            for (int j = V - 1; j >= 0; --j)
                if (a.keep[j] >= lo) { ++ncand; jc = j; xc = a.keep[j]; }   // (jc, xc): the first of them
        }
        const unsigned holders = __ballot_sync(FULL_MASK, ncand > 0);
        const bool single = __popc(holders) == 1 && !__any_sync(FULL_MASK, ncand > 1);
        if (single) {                                                     // warp-uniform
            const int owner = __ffs(holders) - 1;
            if (lane == owner) {
                conf = SIG ? sigmoid_ref(xc, sig_ref) : xc;
                idx = a.bestvi * V + jc;
            }
            return owner;
        }
        if (ncand > 0) {
#pragma unroll
            for (int j = 0; j < V; ++j)
                if (a.keep[j] >= lo) {
                    const float f = SIG ? sigmoid_ref(a.keep[j], sig_ref) : a.keep[j];
                    if (f > fb) { fb = f; fi = a.bestvi * V + j; }
                }
        }
    } else {
        for (int vi = lane; vi < nvec; vi += 32) {
            float x[V];
            Vec<V>::load_cached(src, vi, x);
#pragma unroll
            for (int j = 0; j < V; ++j)
                if (x[j] >= lo) {
                    const float f = sigmoid_ref(x[j], sig_ref);
                    if (f > fb) { fb = f; fi = vi * V + j; }
                }
        }
    }
    warp_argmax_first(fb, fi);
    conf = fb; idx = fi;
    return 0;
}

// ---------------------------------------------------------------- render only
struct SbpRenderParams {
    const void* kp; int kp_f64;
    float* target;
    const float* lut; int lut_n;
    double three_sigma;
    long long n_maps; int H, W, HW; FastDiv divW;
};

template <int V>
__global__ void __launch_bounds__(kSbpThreads) sbp_render_kernel(SbpRenderParams P) {
    extern __shared__ float lut_s[];
    for (int i = threadIdx.x; i < P.lut_n * P.lut_n; i += blockDim.x) lut_s[i] = P.lut[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kSbpWarps + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kSbpWarps;
    const int nvec = P.HW / V;
    for (long long map = warp0; map < P.n_maps; map += nwarps) {
        double x, y;
        load_kp(P.kp, P.kp_f64, map, x, y);
        const Patch pt = make_patch(x, y, P.H, P.W, P.three_sigma, P.lut_n);
        float* out = P.target + map * P.HW;
#pragma unroll 4
        for (int vi = lane; vi < nvec; vi += 32) {
            float t[V];
            patch_values<V>(pt, lut_s, P.lut_n, vi * V, P.W, P.divW, t);
            Vec<V>::store(out, vi, t);
        }
    }
}

// ---------------------------------------------------------------- fused loss (+render, +grad, +decode)
struct SbpFusedParams {
    const float* logits;
    const float* target_in;      // TGT == 1
    const void* kp; int kp_f64;  // TGT == 2
    const float* lut; int lut_n; double three_sigma;
    float* dlogits;              // GRAD
    float* target_out;           // WTGT
    float* joints;               // DEC
    double* partials;            // [gridDim.x][2]  (S_pos, S_neg)
    float thr, scale;
    float gpos, gneg;            // 2*lambda*inv_norm
    long long n_maps; int H, W, HW; FastDiv divW;
    int sig_ref;                 // DEC: which torch sigmoid ranks near-ties and gives the confidence (kSigmoidAtenCpu / kSigmoidAtenCuda)
    ExchangePub xpub;            // multi-GPU in-band exchange: block 0 publishes the previous step's flag (world == 0: off)
};

constexpr int TGT_DENSE = 1;
constexpr int TGT_RENDER = 2;

// One element of the joints-MSE loss in closed form (models/loss/sbp_loss.py:44-47):
//   t > 0 : S_pos += (s-t)^2                      dL/dp = gpos (s-t) s (1-s)
//   t <= 0: S_pos += t^2, S_neg += (s-t)^2        dL/dp = gneg (s-t) s (1-s)
template <bool GRAD>
__device__ __forceinline__ float loss_elem(float s, float t, float gpos, float gneg, float& apos, float& aneg) {
    const float d = s - t;
    const bool pos = t > 0.0f;
    apos = fmaf(pos ? d : t, pos ? d : t, apos);
    aneg = pos ? aneg : fmaf(d, d, aneg);
    if (!GRAD) return 0.0f;
    return (pos ? gpos : gneg) * d * ((1.0f - s) * s);
}

// One vector (V consecutive elements, flat index e = vi*V) of the fused render+loss(+grad) pass.
// Every lane first takes the zero-target result (the target is zero on ~93% of a map):
//   S_neg += s^2,  dL/dp = gneg s^2 (1-s);
// then one unsigned compare decides whether the vector can touch the rows of the joint's Gaussian patch; only those
// lanes compute (row, col), look the template up and replace their elements' results (`arem` collects the s^2 terms
// that have to be taken out of S_neg again, so the common path stays a single FFMA per element).
template <int V, bool GRAD, bool WTGT>
__device__ __forceinline__ void render_loss_vec(const float (&x)[V], float (&g)[V], float (&tout)[V], int vi, const Patch& pt,
                                                const float* __restrict__ lut_s, int lut_n, int W, FastDiv divW, float gpos, float gneg,
                                                float& apos, float& aneg, float& arem) {
    float sg[V];
    sigmoid_vec<V, !GRAD>(x, sg);           // read-only variants: SFU-bound, one reciprocal per vector
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const float sj = sg[j];
        if (GRAD) {
            const float c = sj * sj;
            aneg += c;
            g[j] = gneg * fmaf(-c, sj, c);
        } else {
            aneg = fmaf(sj, sj, aneg);
        }
        if (WTGT) tout[j] = 0.0f;
    }
    const int e = vi * V;
    if ((unsigned)(e + (V - 1) - pt.e0) < (unsigned)(pt.ecnt + (V - 1))) {
        // (the V elements of a vector lie in one row: the launcher takes V = 4 only when W % 4 == 0)
        const int r = (int)fdiv((uint32_t)e, divW);
        const int cc = e - r * W;
        const int pw = lut_n + 2 * kLutPad;
        if constexpr (GRAD) {
            // training kernel (HBM-bound with issue slots to spare): per-element tests, only lanes inside the window do the
            // extra gradient math.  Measured against the branch-free form below: 280 vs 284 us per 69 632 maps; the running
            // (r2, c2) pair schedules better than `cc + j` (278.6 vs 280.1 us, profiles/r01_tune_readonly_b200.log).
            int r2 = r, c2 = cc;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if (r2 >= pt.py0 && r2 < pt.py1 && c2 >= pt.px0 && c2 < pt.px1) {
                    const float t = lut_s[(r2 - pt.uly) * pw + (c2 - pt.ulx) + kLutPad];
                    if (WTGT) tout[j] = t;
                    if (t > 0.0f) {      // t == 0 (underflowed template tail): the zero-target result stands
                        const float d = sg[j] - t;
                        arem = fmaf(sg[j], sg[j], arem);
                        apos = fmaf(d, d, apos);
                        g[j] = gpos * d * ((1.0f - sg[j]) * sg[j]);
                    }
                }
                if (++c2 >= W) { c2 = 0; ++r2; }
            }
        } else {
            // read-only kernels (SFU / issue-bound): rows outside the window -> the all-zero row; columns clamped into the zero
            // pads; `lim` cuts a window that is narrower than the template (non-integer sigma): no per-element branch.
            // 104 -> 55 instructions on this path, 125 M -> 100 M warp instructions per launch; 158.5 -> 145.5 us (loss only).
            const int ri = ((unsigned)(r - pt.uly) < (unsigned)pt.nrows) ? r - pt.uly : lut_n;
            const int ci = min(max(cc - pt.ulx, -kLutPad), lut_n);
            const float* lp = lut_s + ri * pw + ci + kLutPad;
            const int lim = pt.px1 - cc;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float t = lp[j];
                const bool pos = (j < lim) && t > 0.0f;   // t == 0 (pad, or underflowed template tail): the zero-target result stands
                if (WTGT) tout[j] = (j < lim) ? t : 0.0f;
                const float d = sg[j] - t;
                arem = fmaf(pos ? sg[j] : 0.0f, sg[j], arem);
                apos = fmaf(pos ? d : 0.0f, d, apos);
            }
        }
    }
}

// end of one map: flush the fp32 partials into fp64 (keeps the 2e8-term sum accurate and deterministic) and, when decoding,
// resolve the argmax and write the joint row
template <int V, bool DEC>
__device__ __forceinline__ void finish_map(const SbpFusedParams& P, long long map, int lane, float apos, float aneg, float arem,
                                           const ArgTrack<V>& arg, double& dpos, double& dneg) {
    dpos += (double)apos;
    dneg += (double)aneg - (double)arem;
    if (DEC) {
        float conf = -INFINITY;
        int idx = 0x7fffffff;
        const int owner = resolve_argmax<V, true>(arg, P.logits + map * P.HW, P.HW / V, lane, P.sig_ref, conf, idx);
        if (lane == owner) {
            float jx = -1.0f, jy = -1.0f, jc = -1.0f;
            if (conf > P.thr && idx != 0x7fffffff) {
                const int row = (int)fdiv((uint32_t)idx, P.divW);
                jx = (float)(idx - row * P.W);
                jy = (float)row;
                jc = conf;
            }
            float* jo = P.joints + map * 3;
            jo[0] = __fmul_rn(jx, P.scale);
            jo[1] = __fmul_rn(jy, P.scale);
            jo[2] = jc;
        }
    }
}

template <int V, int TGT, bool GRAD, bool WTGT, bool DEC>
__global__ void __launch_bounds__(kSbpThreads, (GRAD || WTGT || TGT != 2) ? POSE_FUSED_MINB : DEC ? POSE_FUSED_MINB_NGD : POSE_FUSED_MINB_NG)
sbp_fused_kernel(SbpFusedParams P) {
    extern __shared__ float lut_s[];
    __shared__ double red[kSbpWarps][2];
    pdl_launch_dependents();      // the epilogue grid may be scheduled as our CTAs retire; it waits for our completion itself
    if (P.xpub.world > 0 && blockIdx.x == 0 && threadIdx.x == 0) exchange_open_step(P.xpub);
    if (TGT == TGT_RENDER) {
        stage_lut_padded(lut_s, P.lut, P.lut_n);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const long long warp0 = (long long)blockIdx.x * kSbpWarps + wid;
    const long long nwarps = (long long)gridDim.x * kSbpWarps;
    const int nvec = P.HW / V;
    constexpr int U = (V == 4) ? ((GRAD || WTGT || TGT != TGT_RENDER) ? POSE_FUSED_U : DEC ? POSE_FUSED_U_NGD : POSE_FUSED_U_NG) : 8;
    double dpos = 0.0, dneg = 0.0;

    // the keypoint of the NEXT map is fetched while the current map streams, so its latency is off the per-map
    // dependency chain (kp -> patch geometry -> first batch of loads)
    double kx = -1.0, ky = -1.0;
    if (TGT == TGT_RENDER && warp0 < P.n_maps) load_kp(P.kp, P.kp_f64, warp0, kx, ky);

    // One stream of (map, batch of U vectors per lane) items per warp.  The loads of the NEXT item -- the next batch of this map
    // or the first batch of the warp's next map -- are issued right after the current batch has been consumed and BEFORE the
    // per-map tail (loss flush, argmax resolution with its shuffles, votes and the reference sigmoid), so that tail runs
    // under memory latency instead of in front of it.  Same registers as the plain "load U, compute U" loop.
    float xv[U][V], tv[U][V];
    auto issue = [&](long long m, int b0) {
        const float* lg = P.logits + m * P.HW;
        const float* tg = (TGT == TGT_DENSE) ? P.target_in + m * P.HW : nullptr;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int vi = b0 + 32 * u;
            if (vi < nvec) {
                Vec<V>::load(lg, vi, xv[u]);
                if (TGT == TGT_DENSE) Vec<V>::load(tg, vi, tv[u]);
            }
        }
    };
    long long map = warp0;
    int base = lane;
    Patch pt;
    float apos = 0.0f, aneg = 0.0f, arem = 0.0f;
    ArgTrack<V> arg;
    arg.reset();
    if (map < P.n_maps) {
        issue(map, base);
        if (TGT == TGT_RENDER) {
            pt = make_patch(kx, ky, P.H, P.W, P.three_sigma, P.lut_n);
            if (map + nwarps < P.n_maps) load_kp(P.kp, P.kp_f64, map + nwarps, kx, ky);
        }
    }
    while (map < P.n_maps) {
        float* dl = GRAD ? P.dlogits + map * P.HW : nullptr;
        float* to = WTGT ? P.target_out + map * P.HW : nullptr;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int vi = base + 32 * u;
            if (vi >= nvec) break;
            float g[V];
            if (DEC) arg.template push<true>(xv[u], vi);
            if (TGT == TGT_RENDER) {
                render_loss_vec<V, GRAD, WTGT>(xv[u], g, tv[u], vi, pt, lut_s, P.lut_n, P.W, P.divW, P.gpos, P.gneg, apos, aneg, arem);
            } else {
                float sg[V];
                sigmoid_vec<V, !GRAD>(xv[u], sg);
#pragma unroll
                for (int j = 0; j < V; ++j) g[j] = loss_elem<GRAD>(sg[j], tv[u][j], P.gpos, P.gneg, apos, aneg);
            }
            if (GRAD) Vec<V>::store(dl, vi, g);
            if (WTGT) Vec<V>::store(to, vi, tv[u]);
        }
        int nbase = base + 32 * U;
        long long nmap = map;
        const bool last = nbase >= nvec;
        if (last) { nbase = lane; nmap = map + nwarps; }
        if (nmap < P.n_maps) issue(nmap, nbase);
        if (last) {
            finish_map<V, DEC>(P, map, lane, apos, aneg, arem, arg, dpos, dneg);
            apos = aneg = arem = 0.0f;
            arg.reset();
            if (TGT == TGT_RENDER && nmap < P.n_maps) {
                pt = make_patch(kx, ky, P.H, P.W, P.three_sigma, P.lut_n);
                if (nmap + nwarps < P.n_maps) load_kp(P.kp, P.kp_f64, nmap + nwarps, kx, ky);
            }
        }
        map = nmap;
        base = nbase;
    }

    dpos = warp_sum(dpos);
    dneg = warp_sum(dneg);
    if (lane == 0) { red[wid][0] = dpos; red[wid][1] = dneg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < kSbpWarps; ++w) { a += red[w][0]; b += red[w][1]; }
        P.partials[2 * blockIdx.x] = a;
        P.partials[2 * blockIdx.x + 1] = b;
    }
}

// Deterministic second stage of the loss: sum `n` (a, b) fp64 pairs, `stride` doubles apart, in a fixed order (no float
// atomics), then loss = (w0*A + w1*B) * inv_norm.  Runs in one CTA.
__device__ __forceinline__ void reduce_pairs_cta(const double* __restrict__ pairs, int n, long long stride, double w0, double w1,
                                                 double inv_norm, float* __restrict__ loss_out, double* __restrict__ num_out) {
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { a += pairs[i * stride]; b += pairs[i * stride + 1]; }
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

__global__ void __launch_bounds__(256) loss_reduce_kernel(const double* __restrict__ pairs, int n, long long stride, double w0, double w1,
                                                          double inv_norm, float* __restrict__ loss_out, double* __restrict__ num_out) {
    pdl_wait();
    reduce_pairs_cta(pairs, n, stride, w0, w1, inv_norm, loss_out, num_out);
}

// back-projection + COCO row fields of one sample by one warp.
// SBPmAPCOCO.update_state (utils/sbp_utils.py:141-163): ratio in fp64 -> fp32, fp32 multiply, fp32 add of
// fp32(bbox origin) (two roundings, no FMA); conf<0 -> (0,0,0); score = left-to-right fp32 sum / K.
// Lane k handles joint k, lane 0 then folds the K confidences in index order through shuffles so the sum is
// bit-identical to the reference's python sum().  Output is packed [N][3K+1] = K rows of (x_img, y_img, flag) followed
// by the score -- ready for one all-gather / D2H.
__device__ __forceinline__ void backproject_sample(const float* __restrict__ joints, const double* __restrict__ bbox,
                                                   float* __restrict__ packed, int n, int K, double in_h, double in_w, int lane) {
    const double bx = __ldg(bbox + 4 * n), by = __ldg(bbox + 4 * n + 1), bw = __ldg(bbox + 4 * n + 2), bh = __ldg(bbox + 4 * n + 3);
    const float rx = (float)(bw / in_w), ry = (float)(bh / in_h);
    const float ox = (float)bx, oy = (float)by;
    const int stride = 3 * K + 1;
    float sum = 0.0f;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        float c = -1.0f;
        if (k < K) {
            const float* j = joints + ((long long)n * K + k) * 3;
            float* o = packed + (long long)n * stride + 3 * k;
            c = j[2];
            if (c < 0.0f) {
                o[0] = 0.0f; o[1] = 0.0f; o[2] = 0.0f;
            } else {
                o[0] = __fadd_rn(__fmul_rn(j[0], rx), ox);
                o[1] = __fadd_rn(__fmul_rn(j[1], ry), oy);
                o[2] = 1.0f;
            }
        }
        const int cnt = min(32, K - k0);
        for (int i = 0; i < cnt; ++i) {
            const float ci = __shfl_sync(FULL_MASK, c, i);
            if (!(ci < 0.0f)) sum = __fadd_rn(sum, ci);
        }
    }
    if (lane == 0) packed[(long long)n * stride + 3 * K] = __fdiv_rn(sum, (float)K);
}

// Epilogue of the fused step, one launch: the LAST CTA reduces the loss partials, the others back-project the decoded
// joints (8 samples per CTA).  Launched with programmatic stream serialisation: it is scheduled while the fused kernel
// drains and blocks in griddepcontrol.wait until that grid has completed and its writes are visible.
struct SbpEpilogueParams {
    const double* partials; int nblocks; double w0, w1, inv_norm; float* loss_out; double* num_out;
    const float* joints; const double* bbox; float* packed; int N, K; double in_h, in_w;
};

__global__ void __launch_bounds__(256) sbp_epilogue_kernel(SbpEpilogueParams P) {
    pdl_wait();
    if (blockIdx.x == gridDim.x - 1) {
        reduce_pairs_cta(P.partials, P.nblocks, 2, P.w0, P.w1, P.inv_norm, P.loss_out, P.num_out);
        return;
    }
    const int n = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (n < P.N) backproject_sample(P.joints, P.bbox, P.packed, n, P.K, P.in_h, P.in_w, threadIdx.x & 31);
}

// dlogits *= *g, whole launch is a no-op when *g == 1 (the usual loss.backward())
__global__ void __launch_bounds__(256) scale_grad_kernel(float* __restrict__ d, const float* __restrict__ g, unsigned long long n) {
    const float s = __ldg(g);
    if (s == 1.0f) return;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        const unsigned long long n4 = n >> 2;
        float4* d4 = reinterpret_cast<float4*>(d);
        for (unsigned long long k = i; k < n4; k += stride) {
            float4 v = d4[k];
            v.x *= s; v.y *= s; v.z *= s; v.w *= s;
            d4[k] = v;
        }
        for (unsigned long long k = (n4 << 2) + i; k < n; k += stride) d[k] *= s;
    } else {
        for (unsigned long long k = i; k < n; k += stride) d[k] *= s;
    }
}

// ---------------------------------------------------------------- decode only
struct SbpDecodeParams {
    const float* x;
    float* joints;
    float thr, scale;
    long long n_maps; int H, W, HW; FastDiv divW;
    int refine;
    int sig_ref;             // SIG: kSigmoidAtenCpu / kSigmoidAtenCuda
};

template <bool SIG>
__device__ __forceinline__ float act(float v) { return SIG ? sigmoid_fast(v) : v; }

template <int V, bool SIG>
__global__ void __launch_bounds__(kSbpThreads) sbp_decode_kernel(SbpDecodeParams P) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kSbpWarps + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kSbpWarps;
    const int nvec = P.HW / V;
    constexpr int U = 8;

    // (map, batch) items as one stream per warp; the next item's loads are issued before the per-map tail (see sbp_fused_kernel)
    float xv[U][V];
    auto issue = [&](long long m, int b0) {
        const float* src = P.x + m * P.HW;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int vi = b0 + 32 * u;
            if (vi < nvec) Vec<V>::load(src, vi, xv[u]);
        }
    };
    long long map = warp0;
    int base = lane;
    ArgTrack<V> arg;
    arg.reset();
    if (map < P.n_maps) issue(map, base);
    while (map < P.n_maps) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int vi = base + 32 * u;
            if (vi >= nvec) break;
            arg.template push<SIG>(xv[u], vi);
        }
        int nbase = base + 32 * U;
        long long nmap = map;
        const bool last = nbase >= nvec;
        if (last) { nbase = lane; nmap = map + nwarps; }
        if (nmap < P.n_maps) issue(nmap, nbase);
        if (last) {
            const float* src = P.x + map * P.HW;
            float best = -INFINITY;
            int besti = 0x7fffffff;
            const int owner = resolve_argmax<V, SIG>(arg, src, nvec, lane, P.sig_ref, best, besti);
            if (lane == owner) {
                float jx = -1.0f, jy = -1.0f, jc = -1.0f;
                if (best > P.thr && besti != 0x7fffffff) {
                    const int row = (int)fdiv((uint32_t)besti, P.divW);
                    const int col = besti - row * P.W;
                    jx = (float)col; jy = (float)row; jc = best;
                    if (P.refine && col > 1 && col < P.W - 1 && row > 1 && row < P.H - 1) {
                        // quarter-pixel shift toward the higher neighbour (NOT in the reference; opt-in)
                        const float dx = act<SIG>(__ldg(src + besti + 1)) - act<SIG>(__ldg(src + besti - 1));
                        const float dy = act<SIG>(__ldg(src + besti + P.W)) - act<SIG>(__ldg(src + besti - P.W));
                        jx += dx > 0.0f ? 0.25f : (dx < 0.0f ? -0.25f : 0.0f);
                        jy += dy > 0.0f ? 0.25f : (dy < 0.0f ? -0.25f : 0.0f);
                    }
                }
                float* jo = P.joints + map * 3;
                jo[0] = __fmul_rn(jx, P.scale);
                jo[1] = __fmul_rn(jy, P.scale);
                jo[2] = jc;
            }
            arg.reset();
        }
        map = nmap;
        base = nbase;
    }
}

// ---------------------------------------------------------------- decode with flip-test averaging (NOT in the reference; opt-in)
// heat[n,k,h,w] = 0.5 * (act(x[n,k,h,w]) + act(xf[n,perm[k],h,W-1-w])): the second forward pass saw the mirrored image, so
// its maps are mirrored back (columns reversed) and left/right joints swapped (perm) before averaging -- the published
// Simple-Baselines test-time rule.  Both maps stream through once; the averaged map is never materialised.
struct SbpDecodeFlipParams {
    const float* x;
    const float* xf;
    const int* perm;         // [K] channel of xf that holds joint k's mirrored map
    float* joints;
    float thr, scale;
    long long n_maps; int K, H, W, HW; FastDiv divW, divK;
    int refine;
};

template <int V, bool SIG>
__global__ void __launch_bounds__(kSbpThreads) sbp_decode_flip_kernel(SbpDecodeFlipParams P) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kSbpWarps + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kSbpWarps;
    const int nvec = P.HW / V;
    const int vec_per_row = P.W / V;             // V == 4 only when W % 4 == 0
    constexpr int U = 4;

    for (long long map = warp0; map < P.n_maps; map += nwarps) {
        const long long n = map / P.K;
        const int k = (int)(map - n * P.K);
        const float* src = P.x + map * P.HW;
        const float* srf = P.xf + (n * P.K + __ldg(P.perm + k)) * P.HW;
        float best = -INFINITY;
        int besti = 0x7fffffff;
        for (int base = lane; base < nvec; base += 32 * U) {
            float xv[U][V], fv[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int vi = base + 32 * u;
                if (vi < nvec) {
                    const int row = (int)fdiv((uint32_t)(vi * V), P.divW);
                    const int cv = vi - row * vec_per_row;                         // vector column inside the row
                    Vec<V>::load(src, vi, xv[u]);
                    Vec<V>::load(srf, row * vec_per_row + (vec_per_row - 1 - cv), fv[u]);     // mirrored chunk, reversed below
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int vi = base + 32 * u;
                if (vi >= nvec) break;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float s = __fmul_rn(0.5f, __fadd_rn(act<SIG>(xv[u][j]), act<SIG>(fv[u][V - 1 - j])));
                    if (s > best) { best = s; besti = vi * V + j; }
                }
            }
        }
        warp_argmax_first(best, besti);

        if (lane == 0) {
            float jx = -1.0f, jy = -1.0f, jc = -1.0f;
            if (best > P.thr && besti != 0x7fffffff) {
                const int row = (int)fdiv((uint32_t)besti, P.divW);
                const int col = besti - row * P.W;
                jx = (float)col; jy = (float)row; jc = best;
                if (P.refine && col > 1 && col < P.W - 1 && row > 1 && row < P.H - 1) {
                    auto heat = [&](int r, int c) {
                        return __fmul_rn(0.5f, __fadd_rn(act<SIG>(__ldg(src + r * P.W + c)), act<SIG>(__ldg(srf + r * P.W + (P.W - 1 - c)))));
                    };
                    const float dx = heat(row, col + 1) - heat(row, col - 1);
                    const float dy = heat(row + 1, col) - heat(row - 1, col);
                    jx += dx > 0.0f ? 0.25f : (dx < 0.0f ? -0.25f : 0.0f);
                    jy += dy > 0.0f ? 0.25f : (dy < 0.0f ? -0.25f : 0.0f);
                }
            }
            float* jo = P.joints + map * 3;
            jo[0] = __fmul_rn(jx, P.scale);
            jo[1] = __fmul_rn(jy, P.scale);
            jo[2] = jc;
        }
    }
}

// ---------------------------------------------------------------- stand-alone back-projection (one warp per sample)
__global__ void __launch_bounds__(256) sbp_backproject_kernel(const float* __restrict__ joints, const double* __restrict__ bbox,
                                                              float* __restrict__ packed, int N, int K, double in_h, double in_w) {
    const int n = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (n < N) backproject_sample(joints, bbox, packed, n, K, in_h, in_w, threadIdx.x & 31);
}

// ---------------------------------------------------------------- the reference sigmoids (diagnostics)
// y[i] = sigmoid_ref(x[i]): lets a test compare the device restatements with torch.sigmoid bit for bit
__global__ void sigmoid_ref_eval_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned long long n, int sig_ref) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = sigmoid_ref(x[i], sig_ref);
}

// For fp32 m in (-80, +inf], every 61st float: nothing below sigmoid_window_lo(m) may reach sigmoid_ref(m) -- probed at the 64 floats just below
// the window and at 64 geometrically spaced points further down (the references are monotone up to a few ulp, so these are
// where a violation would be).  Counts violations.
__global__ void sigmoid_window_check_kernel(unsigned long long* violations, int sig_ref) {
    const uint32_t k0 = float_key(-80.0f), k1 = float_key(INFINITY);
    const uint64_t total = (uint64_t)(k1 - k0);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        // sampled m: all of them would be 2^32 x 128 evaluations; every 61st float still covers every binade densely
        if (i % 61 != 0) continue;
        const float m = key_float(k0 + (uint32_t)i);
        const float fm = sigmoid_ref(m, sig_ref);
        const float lo = sigmoid_window_lo(m);
        const uint32_t kl = float_key(lo);
        for (uint32_t d = 1; d <= 64; ++d) {
            if (kl < d) break;
            if (!(sigmoid_ref(key_float(kl - d), sig_ref) < fm)) ++bad;
        }
        float w = m - lo;
        for (int d = 0; d < 64; ++d) {
            w *= 1.25f;
            if (!(sigmoid_ref(lo - w, sig_ref) < fm)) ++bad;
        }
    }
    bad = __reduce_add_sync(FULL_MASK, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(violations, (unsigned long long)bad);
}

}  // namespace pose
