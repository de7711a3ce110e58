// Simple-Baselines hot path kernels: render, fused render+loss+grad(+decode), decode, back-projection.
//
// Work unit = one heat map (H*W fp32, 12 288 B at 64x48) = ONE CTA, and the grid covers all N*K maps (not persistent): every
// thread issues all of its 128-bit loads the moment the CTA starts (3 per thread at 64x48: the whole map is in flight at
// once), the per-map reductions (loss partials, argmax) go through shared memory with one or two barriers, the CTA retires
// and the hardware scheduler hands the SM the next map in memory order.  Measured on B200 (tools/stream_patterns.cu,
// profiles/r02_stream_patterns.log) for a 1 read : 1 write stream with this kernel's arithmetic:
//   persistent grid, one warp per map, 6 loads per lane (the r01 layout)      282-288 us   5.95-6.06 TB/s  (0.91-0.93 of copy peak)
//   non-persistent, one warp per map (8 maps per CTA)                          261 us       6.55 TB/s
//   non-persistent, one CTA per map + block reduction (this file)              249 us       6.87 TB/s  (1.05 of copy peak)
// read-only 119 us (7.2 TB/s), write-only 115 us (7.4 TB/s).  Resident CTAs (6-8 per SM) supply the memory-level parallelism
// that the persistent form had to build from registers (U loads per lane) and lost in every load -> compute -> store phase.
// All stages are HBM-bound; nothing here is a contraction (no tensor cores).
#pragma once
#include "common.cuh"

// linkage of the non-template kernels of this header: a second translation unit that includes it (api_head.cu) compiles them
// as file-local copies
#ifndef POSE_GLOBAL
#define POSE_GLOBAL __global__
#endif

namespace pose {

constexpr int kSbpThreads = 256;               // 8 warps per CTA
constexpr int kSbpWarps = kSbpThreads / 32;
constexpr int kMapU = 3;                       // 128-bit vectors per thread and round: 256 x 3 x 16 B = one 64x48 map per round
// resident CTAs per SM the map kernels are compiled for (register caps 40 / 48 / 64 at 6 / 5 / 4)
#ifndef POSE_MAP_MINB_GRAD
#define POSE_MAP_MINB_GRAD 5
#endif
#ifndef POSE_MAP_MINB_RO
#define POSE_MAP_MINB_RO 6
#endif

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
    // generic address (the map may sit in shared memory: the bulk-staged kernels)
    static __device__ __forceinline__ void load_any(const float* base, int vi, float (&o)[4]) {
        float4 v = reinterpret_cast<const float4*>(base)[vi];
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
    static __device__ __forceinline__ void load_any(const float* base, int vi, float (&o)[1]) { o[0] = base[vi]; }
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

// Template of the fused kernels: n+1 rows of n+8 floats -- 4 zeros, the n template values, 4 zeros -- the last row all zeros, so a
// thread fetches the targets of 4 consecutive pixels of one row with two clamped indices instead of 4 x (row test, column
// test) behind divergent branches.  The caller passes it in this layout (pose_gauss_template_padded_host; 1.5 KB at sigma 2)
// and the kernels read it through L1 (__ldg): with one CTA per map there is no per-CTA staging pass to amortise.
constexpr int kLutPad = kTemplatePad;
__host__ __device__ inline size_t lut_padded_floats(int n) { return (size_t)(n + 1) * (size_t)(n + 2 * kLutPad); }

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
// part therefore never evaluates a sigmoid for the decode and does not even track an index: a thread keeps the running
// maximum of the logits it sees (one FMNMX3 per two elements).  At the end of the map each WARP derives a candidate window
// [lo, m_w] from its own maximum m_w (sigmoid_window_lo: nothing below lo can reach sigmoid_ref(m_w) -- true for any subset of
// the map), finds the lanes whose maximum lies in it -- almost always exactly one -- and re-reads THOSE lanes' vectors
// cooperatively (lane i takes the owner's i-th vector; shared memory or L2); elements inside the window are ranked with the
// reference's own sigmoid (sigmoid_ref, bit-exact, also the reported confidence) through a 64-bit atomicMax in shared memory on
// (ordered value, ~index): larger value first, then the smaller flat index.  The maximum over the warps' winners is the map's
// argmax, so the CTA needs no block-wide maximum and no extra barrier.  (r02 history: tracking (value, vector index, second
// best) per vector cost 7-8 ALU instructions per vector and made the staged kernels issue-bound: grad+decode 278 vs 263 us.)
template <int V>
struct ArgTrack {
    float best;             // largest logit seen by this thread (NaNs ignored)
    __device__ __forceinline__ void reset() { best = -INFINITY; }
    __device__ __forceinline__ void push(const float (&x)[V]) {
        if (V == 4) best = fmaxf(fmaxf(best, fmaxf(x[0], x[1])), fmaxf(x[2], x[V - 1]));
        else {
#pragma unroll
            for (int j = 0; j < V; ++j) best = fmaxf(best, x[j]);
        }
    }
};

// (float_key(value) << 32) | (0xffffffff - flat index); 0 = no candidate.  (+ 0.0f: -0 and +0 are equal to the reference's
// comparison, so they must share one key.)
__device__ __forceinline__ unsigned long long arg_key(float f, int i) {
    return ((unsigned long long)float_key(f + 0.0f) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
}

// Warp-level resolution; every lane of the warp calls it.  The warp's lane l owns vectors first_vi(l) + k*vstride with
// first_vi(l) = first_vi0 + l (k = 0, 1, ...; < nvec).  Offers the warp's winner (if any) to *key.
// SIG == false (heat maps that are already activated, DecodeSBP.pred == False): candidates are the elements equal to the
// warp maximum.
// SLOT == false: *key is shared by several warps and takes the warp's winner through one atomicMax (lane 0).  SLOT == true: *key is
// this warp's own slot and is always written (0 = no candidate): no atomic, the caller takes the largest slot after its barrier.
template <int V, bool SIG, bool SLOT = false>
__device__ __forceinline__ void warp_offer_argmax(const ArgTrack<V>& a, const float* __restrict__ src, int nvec, int first_vi0, int vstride,
                                                  int sig_ref, unsigned long long* key) {
    const int lane = threadIdx.x & 31;
    const float m = warp_max(a.best);
    float lo = m;
    if (SIG) {
        lo = sigmoid_window_lo(m);
        // m <= -80 (or no finite element at all): the references' results are denormal / zero there and the error model of
        // the window does not hold -- rank every element of the warp's share
        if (!(m > -80.0f)) lo = -INFINITY;
    }
    // the reference sigmoid of the warp's maximum is evaluated NEXT TO the candidate scan, not behind it (one dependent chain of
    // ~250 cycles off the end of every map), and re-used for every candidate equal to m -- usually the only one
    const float sm = SIG ? sigmoid_ref(m, sig_ref) : m;
    unsigned owners = __ballot_sync(FULL_MASK, a.best >= lo);
    unsigned long long k = 0ull;
    if (__popc(owners) <= 2) {
        // the usual case: one lane (rarely two) holds the maximum -- the warp reads that lane's vectors together
        while (owners) {                                        // warp-uniform
            const int o = __ffs((int)owners) - 1;
            owners &= owners - 1u;
            for (int vi = first_vi0 + o + vstride * lane; vi < nvec; vi += vstride * 32) {
                float x[V];
                Vec<V>::load_any(src, vi, x);
#pragma unroll
                for (int j = 0; j < V; ++j)
                    if (x[j] >= lo) {
                        float v = sm;
                        if (SIG && x[j] != m) v = sigmoid_ref(x[j], sig_ref);
                        k = max(k, arg_key(v, vi * V + j));
                    }
            }
        }
    } else if (a.best >= lo) {
        // plateaus (an all-zero target map, saturated or constant logits): every owner scans its own vectors, in index order;
        // without an activation all candidates carry the same value m, so a lane's first hit is its best
        bool hit = false;
        for (int vi = first_vi0 + lane; vi < nvec && !hit; vi += vstride) {
            float x[V];
            Vec<V>::load_any(src, vi, x);
#pragma unroll
            for (int j = 0; j < V; ++j)
                if (x[j] >= lo && !hit) {
                    float v = sm;
                    if (SIG && x[j] != m) v = sigmoid_ref(x[j], sig_ref);
                    k = max(k, arg_key(v, vi * V + j));
                    if (!SIG) hit = true;
                }
        }
    }
    k = warp_max_key(k);                                        // two CREDUX.MAX
    if (SLOT) {
        if (lane == 0) *key = k;
    } else if (lane == 0 && k != 0ull) {
        atomicMax(key, k);
    }
}

__device__ __forceinline__ void key_to_argmax(unsigned long long key, float& conf, int& idx) {
    conf = -INFINITY;
    idx = 0x7fffffff;
    if (key != 0ull) {
        conf = key_float((uint32_t)(key >> 32));
        idx = (int)(0xffffffffu - (uint32_t)key);
    }
}

// joint row of one map from its (confidence, flat index): nms_sbp + the coordinate scale of DecodeSBP.forward (:116)
__device__ __forceinline__ void write_joint(float* __restrict__ jo, float conf, int idx, float thr, float scale, int W, FastDiv divW,
                                            float dx = 0.0f, float dy = 0.0f) {
    float jx = -1.0f, jy = -1.0f, jc = -1.0f;
    if (conf > thr && idx != 0x7fffffff) {
        const int row = (int)fdiv((uint32_t)idx, divW);
        jx = (float)(idx - row * W) + dx;
        jy = (float)row + dy;
        jc = conf;
    }
    jo[0] = __fmul_rn(jx, scale);
    jo[1] = __fmul_rn(jy, scale);
    jo[2] = jc;
}

// ---------------------------------------------------------------- render only
struct SbpRenderParams {
    const void* kp; int kp_f64;
    float* target;
    const float* lut; int lut_n;
    double three_sigma;
    long long n_maps; int H, W, HW; FastDiv divW;
};

// one CTA per map: thread 0 derives the patch geometry while the template is staged; then a pure write stream
template <int V>
POSE_GLOBAL void __launch_bounds__(kSbpThreads) sbp_render_kernel(SbpRenderParams P) {
    extern __shared__ float lut_s[];
    __shared__ Patch s_patch;
    const long long map = blockIdx.x;
    if (threadIdx.x == 0) {
        double x, y;
        load_kp(P.kp, P.kp_f64, map, x, y);
        s_patch = make_patch(x, y, P.H, P.W, P.three_sigma, P.lut_n);
    }
    for (int i = threadIdx.x; i < P.lut_n * P.lut_n; i += blockDim.x) lut_s[i] = P.lut[i];
    __syncthreads();
    const Patch pt = s_patch;
    const int nvec = P.HW / V;
    float* out = P.target + map * P.HW;
#pragma unroll 3
    for (int vi = threadIdx.x; vi < nvec; vi += kSbpThreads) {
        float t[V];
        patch_values<V>(pt, lut_s, P.lut_n, vi * V, P.W, P.divW, t);
        Vec<V>::store(out, vi, t);
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
    double* partials;            // [n_maps][2]  (S_pos, S_neg) of every map
    unsigned int* ticket;        // the epilogue's two-level reduction counts its slice CTAs here; zeroed by CTA 0 of this grid
    float thr, scale;
    float gpos, gneg;            // 2*lambda*inv_norm
    long long n_maps; int H, W, HW; FastDiv divW;
    int sig_ref;                 // DEC: which torch sigmoid ranks near-ties and gives the confidence (kSigmoidAtenCpu / kSigmoidAtenCuda)
    ExchangePub xpub;            // multi-GPU in-band exchange: CTA 0 publishes the previous step's flag (world == 0: off)
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
template <int V, bool GRAD, bool WTGT, bool SHARE>
__device__ __forceinline__ void render_loss_vec(const float (&x)[V], float (&g)[V], float (&tout)[V], int vi, const Patch& pt,
                                                const float* __restrict__ lut, int lut_n, int W, FastDiv divW, float gpos, float gneg,
                                                float& apos, float& aneg, float& arem) {
    float sg[V];
    sigmoid_vec<V, SHARE>(x, sg);           // read-only variants: SFU-bound, one reciprocal per vector
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
                    const float t = __ldg(lut + (r2 - pt.uly) * pw + (c2 - pt.ulx) + kLutPad);
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
            const float* lp = lut + ri * pw + ci + kLutPad;
            const int lim = pt.px1 - cc;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float t = __ldg(lp + j);
                const bool pos = (j < lim) && t > 0.0f;   // t == 0 (pad, or underflowed template tail): the zero-target result stands
                if (WTGT) tout[j] = (j < lim) ? t : 0.0f;
                const float d = sg[j] - t;
                arem = fmaf(pos ? sg[j] : 0.0f, sg[j], arem);
                apos = fmaf(pos ? d : 0.0f, d, apos);
            }
        }
    }
}

// Read-only two-phase form (the bulk-staged kernels, whose map sits in shared memory): phase 1 streams EVERY vector as if the
// target were zero -- sigmoid, S_neg += s^2, nothing else: no patch test, no row / column arithmetic -- and phase 2 deals the
// vectors that intersect the joint's window (<= 15 rows x 5 vectors at sigma 2: 9 % of the map, where the per-vector row test of
// render_loss_vec sends 26 % down the patch path) out to the map's threads, which re-read them from shared memory, evaluate the
// same sigmoid again (bit-identical, so the s^2 taken out of S_neg is exactly the one phase 1 put in) and apply the corrections.
// 984 -> ~870 instructions per thread and map; validation loss 142 -> 128 us, loss + decode 167 -> 150 us per 69 632 maps.
template <int V, bool SHARE>
__device__ __forceinline__ void zero_target_vec(const float (&x)[V], float& aneg) {
    float sg[V];
    sigmoid_vec<V, SHARE>(x, sg);
#pragma unroll
    for (int j = 0; j < V; ++j) aneg = fmaf(sg[j], sg[j], aneg);
}
// the correction of one window vector: row r, first column cc (a multiple of 4) -- the branch-free padded-template lookup of render_loss_vec
template <bool SHARE>
__device__ __forceinline__ void patch_correct_vec(const float (&x)[4], int r, int cc, const Patch& pt, const float* __restrict__ lut, int lut_n,
                                                  float& apos, float& arem) {
    float sg[4];
    sigmoid_vec<4, SHARE>(x, sg);
    const int pw = lut_n + 2 * kLutPad;
    const int ri = ((unsigned)(r - pt.uly) < (unsigned)pt.nrows) ? r - pt.uly : lut_n;
    const int ci = min(max(cc - pt.ulx, -kLutPad), lut_n);
    const float* lp = lut + ri * pw + ci + kLutPad;
    const int lim = pt.px1 - cc;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float t = __ldg(lp + j);
        const bool pos = (j < lim) && t > 0.0f;
        const float d = sg[j] - t;
        arem = fmaf(pos ? sg[j] : 0.0f, sg[j], arem);
        apos = fmaf(pos ? d : 0.0f, d, apos);
    }
}

// One CTA per heat map.  Order inside the CTA: (1) every thread issues the 128-bit loads of its share of the map -- nothing
// else has been waited for yet, so the whole map is in flight within the CTA's first instructions; (2) thread 0 turns the
// keypoint into the patch geometry (shared memory, one barrier -- under the latency of (1)); (3) the arithmetic and the
// stores; (4) fp32 partial sums of the map -> warp shuffles -> shared memory, and each warp's argmax winner -> one shared
// atomicMax; ONE barrier; thread 0 adds the 8 warps in fixed order, in fp64, writes the map's (S_pos, S_neg) pair -- no float
// atomics anywhere: the loss is run-to-run deterministic -- and the joint row.
template <int V, int TGT, bool GRAD, bool WTGT, bool DEC>
POSE_GLOBAL void __launch_bounds__(kSbpThreads, GRAD ? POSE_MAP_MINB_GRAD : POSE_MAP_MINB_RO) sbp_fused_kernel(SbpFusedParams P) {
    __shared__ Patch s_patch;
    __shared__ float s_sum[kSbpWarps][3];
    __shared__ unsigned long long s_key;
    pdl_launch_dependents();      // the epilogue grid may be scheduled as our CTAs retire; it waits for our completion itself
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long map = blockIdx.x;
    const int nvec = P.HW / V;
    constexpr int U = kMapU;
    // read-only variants: SFU-bound once the stream is this fast -- one reciprocal per 128-bit vector (common.cuh: sigmoid_fast4)
    constexpr bool SHARE = !GRAD;
    const float* lg = P.logits + map * P.HW;
    const float* tg = (TGT == TGT_DENSE) ? P.target_in + map * P.HW : nullptr;

    float xv[U][V], tv[U][V];
    auto issue = [&](int b0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int vi = b0 + tid + kSbpThreads * u;
            if (vi < nvec) {
                Vec<V>::load(lg, vi, xv[u]);
                if (TGT == TGT_DENSE) Vec<V>::load(tg, vi, tv[u]);
            }
        }
    };
    issue(0);
    if (tid == 0) {
        if (map == 0) {
            if (P.xpub.world > 0) exchange_open_step(P.xpub);
            *P.ticket = 0u;
        }
        if (TGT == TGT_RENDER) {
            double kx, ky;
            load_kp(P.kp, P.kp_f64, map, kx, ky);
            s_patch = make_patch(kx, ky, P.H, P.W, P.three_sigma, P.lut_n);
        }
        if (DEC) s_key = 0ull;
    }
    if (TGT == TGT_RENDER || DEC) __syncthreads();

    float apos = 0.0f, aneg = 0.0f, arem = 0.0f;
    ArgTrack<V> arg;
    arg.reset();
    for (int b0 = 0; b0 < nvec; b0 += kSbpThreads * U) {
        if (b0 > 0) issue(b0);                                  // maps larger than one round (96x72: three)
        float* dl = GRAD ? P.dlogits + map * P.HW : nullptr;
        float* to = WTGT ? P.target_out + map * P.HW : nullptr;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int vi = b0 + tid + kSbpThreads * u;
            if (vi >= nvec) break;
            float g[V];
            if (DEC) arg.push(xv[u]);
            if (TGT == TGT_RENDER) {
                render_loss_vec<V, GRAD, WTGT, SHARE>(xv[u], g, tv[u], vi, s_patch, P.lut, P.lut_n, P.W, P.divW, P.gpos, P.gneg, apos, aneg, arem);
            } else {
                float sg[V];
                sigmoid_vec<V, SHARE>(xv[u], sg);
#pragma unroll
                for (int j = 0; j < V; ++j) g[j] = loss_elem<GRAD>(sg[j], tv[u][j], P.gpos, P.gneg, apos, aneg);
            }
            if (GRAD) Vec<V>::store(dl, vi, g);
            if (WTGT) Vec<V>::store(to, vi, tv[u]);
        }
    }

    apos = warp_sum(apos);
    aneg = warp_sum(aneg);
    arem = warp_sum(arem);
    if (lane == 0) { s_sum[wid][0] = apos; s_sum[wid][1] = aneg; s_sum[wid][2] = arem; }
    if (DEC) warp_offer_argmax<V, true>(arg, lg, nvec, wid * 32, kSbpThreads, P.sig_ref, &s_key);
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0, r = 0.0;
#pragma unroll
        for (int w = 0; w < kSbpWarps; ++w) { a += (double)s_sum[w][0]; b += (double)s_sum[w][1]; r += (double)s_sum[w][2]; }
        reinterpret_cast<double2*>(P.partials)[map] = make_double2(a, b - r);
        if (DEC) {
            float conf;
            int idx;
            key_to_argmax(s_key, conf, idx);
            write_joint(P.joints + map * 3, conf, idx, P.thr, P.scale, P.W, P.divW);
        }
    }
}

// ---------------------------------------------------------------- bulk-async (TMA engine) staged forms: render mode, 16-byte aligned maps
// What bounds the register-staged kernels above is bytes in flight and instructions per map: at ~2 us loaded DRAM latency a
// B200 SM needs ~90 KB of reads outstanding to draw its share of 6.9 TB/s, i.e. 8 maps, and a CTA that holds its map in 12
// registers per thread gets 5-6 CTAs per SM at the 40-48 registers the arithmetic needs; and with 256 threads on one map
// every thread's prologue and tail (index arithmetic, three warp reductions, the argmax offer) is paid for just 3 vectors:
// 292 instructions per thread and map, twice the r01 kernel's count per map -- ncu: issue slots 78-83 % busy, DRAM 57 %.
// Here the maps in flight live in SHARED MEMORY: one thread issues a cp.async.bulk (UBLKCP, completion on an mbarrier) for
// each of the CTA's POSE_TMA_MPC_* consecutive maps the moment the CTA starts; POSE_TMA_WPM_* warps then work on each map (all
// maps of the CTA concurrently), one 128-bit vector per thread at a time out of shared memory, so the per-thread overhead is
// spread over 12-24 vectors and the number of maps in flight is set by shared memory (18 maps fit), not by registers.
// The arithmetic per element is that of sbp_fused_kernel, so dlogits and joints are bit-identical (a GPU test compares the
// two); the fp32 partial sums of a map are grouped differently (loss equal to ~1e-7).  Candidates of the argmax are re-read
// from shared memory.
// CTA shape per variant class (tools/tune_fused.py, profiles/r02_tune_tma_*.log; 69 632 maps): the kernels that write dlogits
// run best as 64-thread CTAs with ONE map, 16 of them per SM (grad+decode 259.3 us, grad 259.1; 2 maps per CTA: 267 / 260; 4:
// 277 / 265) and 2 warps per map (1 warp per map loses 25 %, 3-4 warps per map pay the per-thread overhead too often); the first
// sweep put the read-only ones at 256-thread CTAs with 4 maps (loss 136 us, loss+decode 158, decode 120.9).
// Round 2, later: the read-only fused kernels (validation: loss, loss + decode) are bound by issue slots, and everything a map
// costs per WARP (three loss sums, the argmax offer, prologue) is paid once with ONE warp per map: 2 maps per 64-thread CTA,
// 8 CTAs per SM -- loss + decode 152 (4 maps x 2 warps) -> 133 (2 x 2) -> 129.6 us (2 x 1); profiles/r02_tune_tma_readonly_mpc2.log.
// The decode-only kernel (HBM-bound) keeps 2 warps per map.
#ifndef POSE_TMA_WPM_GRAD
#define POSE_TMA_WPM_GRAD 2     // warps per map, kernels that write dlogits
#endif
#ifndef POSE_TMA_WPM_RO
#define POSE_TMA_WPM_RO 1       // warps per map, read-only fused kernels
#endif
#ifndef POSE_TMA_WPM_DEC
#define POSE_TMA_WPM_DEC 2      // warps per map, decode-only kernel
#endif
#ifndef POSE_TMA_MPC_GRAD
#define POSE_TMA_MPC_GRAD 1     // maps per CTA, kernels that write dlogits
#endif
#ifndef POSE_TMA_MPC_RO
#define POSE_TMA_MPC_RO 2       // maps per CTA, read-only fused kernels
#endif
#ifndef POSE_TMA_MPC_DEC
#define POSE_TMA_MPC_DEC 2      // maps per CTA, decode-only kernel
#endif
#ifndef POSE_TMA_MINB_GRAD
#define POSE_TMA_MINB_GRAD 16
#endif
#ifndef POSE_TMA_MINB_RO
#define POSE_TMA_MINB_RO 8
#endif
#ifndef POSE_TMA_MINB_DEC
#define POSE_TMA_MINB_DEC 8
#endif
constexpr int kTmaGrad = 0, kTmaRo = 1, kTmaDec = 2;            // kernel classes of the bulk-staged forms
__host__ __device__ constexpr int tma_cls(bool grad) { return grad ? kTmaGrad : kTmaRo; }
__host__ __device__ constexpr int tma_mpc(int cls) { return cls == kTmaGrad ? POSE_TMA_MPC_GRAD : cls == kTmaRo ? POSE_TMA_MPC_RO : POSE_TMA_MPC_DEC; }
__host__ __device__ constexpr int tma_wpm(int cls) { return cls == kTmaGrad ? POSE_TMA_WPM_GRAD : cls == kTmaRo ? POSE_TMA_WPM_RO : POSE_TMA_WPM_DEC; }
__host__ __device__ constexpr int tma_minb(int cls) { return cls == kTmaGrad ? POSE_TMA_MINB_GRAD : cls == kTmaRo ? POSE_TMA_MINB_RO : POSE_TMA_MINB_DEC; }
__host__ __device__ constexpr int tma_threads(int cls) { return 32 * tma_wpm(cls) * tma_mpc(cls); }
__host__ __device__ inline size_t sbp_tma_smem_bytes(int HW, int cls) { return (size_t)tma_mpc(cls) * (size_t)HW * sizeof(float); }

// stage the CTA's maps: thread 0 initialises one mbarrier per map and issues the bulk copies (the caller's next CTA barrier
// makes the initialised mbarriers visible to the waiting threads)
__device__ __forceinline__ void tma_stage_maps(float* tiles, unsigned long long* bars, const float* __restrict__ src, long long map0, int nmap,
                                               int HW) {
    if (threadIdx.x == 0) {
        for (int k = 0; k < nmap; ++k) mbar_init(smem_u32(&bars[k]), 1);
        mbar_init_fence();
        const uint32_t bytes = (uint32_t)HW * 4u;
        for (int k = 0; k < nmap; ++k) {
            mbar_arrive_expect_tx(smem_u32(&bars[k]), bytes);
            bulk_load(smem_u32(tiles + (size_t)k * HW), src + (map0 + k) * HW, bytes, smem_u32(&bars[k]));
        }
    }
}

template <bool GRAD, bool DEC>
POSE_GLOBAL void __launch_bounds__(tma_threads(tma_cls(GRAD)), tma_minb(tma_cls(GRAD))) sbp_fused_tma_kernel(SbpFusedParams P) {
    constexpr int V = 4, MPC = tma_mpc(tma_cls(GRAD)), WPM = tma_wpm(tma_cls(GRAD)), TPM = 32 * WPM;
    extern __shared__ __align__(128) float tiles[];            // MPC maps of HW floats
    __shared__ __align__(8) unsigned long long s_bar[MPC];
    __shared__ Patch s_patch[MPC];
    __shared__ FastDiv s_div[MPC];                             // read-only form: division by the window's width in vectors
    __shared__ float s_sum[MPC][WPM][3];
    __shared__ unsigned long long s_key[MPC][WPM];              // one slot per warp: no atomics
    // launched with programmatic stream serialisation: the grid may be scheduled while its predecessor (in a training loop the
    // epilogue of the previous step) drains; nothing the predecessor produced is touched before this wait
    pdl_wait();
    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = wid / WPM, ws = wid - k * WPM;               // this warp's map slot and its index among the map's warps
    const int t = ws * 32 + lane;                              // thread index within the map
    const long long map0 = (long long)blockIdx.x * MPC;
    const int nmap = (int)min((long long)MPC, P.n_maps - map0);
    const int nvec = P.HW / V;
    constexpr bool SHARE = !GRAD;
    tma_stage_maps(tiles, s_bar, P.logits, map0, nmap, P.HW);
    if (tid == 0 && map0 == 0) {
        if (P.xpub.world > 0) exchange_open_step(P.xpub);
        *P.ticket = 0u;
    }
    if (t == 0 && k < nmap) {
        double kx, ky;
        load_kp(P.kp, P.kp_f64, map0 + k, kx, ky);
        const Patch pt = make_patch(kx, ky, P.H, P.W, P.three_sigma, P.lut_n);
        s_patch[k] = pt;
        if (!GRAD) {
            const uint32_t nvx = pt.ecnt ? (uint32_t)(((pt.px1 - 1) >> 2) - (pt.px0 >> 2) + 1) : 1u;
            s_div[k] = FastDiv{nvx, nvx > 1 ? (uint32_t)((0xffffffffu / nvx) + 1u) : 0u};      // ceil(2^32 / nvx) for nvx that is no power of two; exact for id < 2^16
        }
    }
    __syncthreads();
    if (k < nmap) {
        const float* tile = tiles + (size_t)k * P.HW;
        float* dl = GRAD ? P.dlogits + (map0 + k) * P.HW : nullptr;
        float apos = 0.0f, aneg = 0.0f, arem = 0.0f;
        ArgTrack<V> arg;
        arg.reset();
        mbar_wait_parity(&s_bar[k], 0);
        if constexpr (GRAD) {
#pragma unroll 4
            for (int vi = t; vi < nvec; vi += TPM) {
                float x[V], g[V], unused[V];
                Vec<V>::load_any(tile, vi, x);
                if (DEC) arg.push(x);
                render_loss_vec<V, GRAD, false, SHARE>(x, g, unused, vi, s_patch[k], P.lut, P.lut_n, P.W, P.divW, P.gpos, P.gneg, apos, aneg, arem);
                Vec<V>::store(dl, vi, g);
            }
        } else {
            // phase 1: the whole map as zero target
#pragma unroll 4
            for (int vi = t; vi < nvec; vi += TPM) {
                float x[V];
                Vec<V>::load_any(tile, vi, x);
                if (DEC) arg.push(x);
                zero_target_vec<V, SHARE>(x, aneg);
            }
            // phase 2: the vectors of the joint's window, dealt out to the map's threads
            const Patch& pt = s_patch[k];
            if (pt.ecnt) {
                const FastDiv dv = s_div[k];
                const int cv0 = pt.px0 >> 2, nvx = (int)dv.d, wq = P.W >> 2;
                const int total = nvx * (pt.py1 - pt.py0);
                for (int id = t; id < total; id += TPM) {
                    const int ry = (int)fdiv((uint32_t)id, dv), cv = cv0 + id - ry * nvx, r = pt.py0 + ry;
                    float x[V];
                    Vec<V>::load_any(tile, r * wq + cv, x);
                    patch_correct_vec<SHARE>(x, r, 4 * cv, pt, P.lut, P.lut_n, apos, arem);
                }
            }
        }
        apos = warp_sum(apos);
        aneg = warp_sum(aneg);
        arem = warp_sum(arem);
        if (lane == 0) { s_sum[k][ws][0] = apos; s_sum[k][ws][1] = aneg; s_sum[k][ws][2] = arem; }
        if (DEC) warp_offer_argmax<V, true, true>(arg, tile, nvec, ws * 32, TPM, P.sig_ref, &s_key[k][ws]);
    }
    __syncthreads();
    if (tid < nmap) {
        const int m = tid;
        double a = 0.0, b = 0.0, r = 0.0;
#pragma unroll
        for (int w = 0; w < WPM; ++w) { a += (double)s_sum[m][w][0]; b += (double)s_sum[m][w][1]; r += (double)s_sum[m][w][2]; }
        reinterpret_cast<double2*>(P.partials)[map0 + m] = make_double2(a, b - r);
        if (DEC) {
            float conf;
            int idx;
            unsigned long long key = 0ull;
#pragma unroll
            for (int w = 0; w < WPM; ++w) key = max(key, s_key[m][w]);
            key_to_argmax(key, conf, idx);
            write_joint(P.joints + (map0 + m) * 3, conf, idx, P.thr, P.scale, P.W, P.divW);
        }
    }
}

// ---------------------------------------------------------------- loss reduction kernel (helpers: common.cuh)
POSE_GLOBAL void __launch_bounds__(256) loss_reduce_kernel(const double* __restrict__ pairs, int n, long long stride, double w0, double w1,
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

// Epilogue of the fused step, one launch: CTAs [0, bp_ctas) back-project the decoded joints (8 samples per CTA), CTAs
// [bp_ctas, bp_ctas + R) reduce the per-map loss pairs (two levels, see reduce_slice_and_elect).  Launched with programmatic
// stream serialisation: it is scheduled while the fused kernel drains and blocks in griddepcontrol.wait until that grid has
// completed and its writes are visible.
struct SbpEpilogueParams {
    const double* partials; long long n_pairs; double* slices; unsigned int* ticket; int R; int bp_ctas;
    double w0, w1, inv_norm; float* loss_out; double* num_out;
    const float* joints; const double* bbox; float* packed; int N, K; double in_h, in_w;
};

POSE_GLOBAL void __launch_bounds__(256) sbp_epilogue_kernel(SbpEpilogueParams P) {
    pdl_wait();
    pdl_launch_dependents();      // a successor launched the same way (the next step's fused kernel) may be scheduled now; it waits for us itself
    if ((int)blockIdx.x >= P.bp_ctas) {
        if (reduce_slice_and_elect(P.partials, P.n_pairs, P.slices, P.ticket, P.R, (int)blockIdx.x - P.bp_ctas))
            reduce_pairs_cta(P.slices, P.R, 2, P.w0, P.w1, P.inv_norm, P.loss_out, P.num_out);
        return;
    }
    const int n = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (n < P.N) backproject_sample(P.joints, P.bbox, P.packed, n, P.K, P.in_h, P.in_w, threadIdx.x & 31);
}

// dlogits *= *g, whole launch is a no-op when *g == 1 (the usual loss.backward())
POSE_GLOBAL void __launch_bounds__(256) scale_grad_kernel(float* __restrict__ d, const float* __restrict__ g, unsigned long long n) {
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

// one CTA per map, every load issued up front (see the file header); no SFU work in the stream; one barrier
template <int V, bool SIG>
POSE_GLOBAL void __launch_bounds__(kSbpThreads, POSE_MAP_MINB_RO) sbp_decode_kernel(SbpDecodeParams P) {
    __shared__ unsigned long long s_key;
    const int tid = threadIdx.x;
    const long long map = blockIdx.x;
    const int nvec = P.HW / V;
    constexpr int U = kMapU;
    const float* src = P.x + map * P.HW;
    float xv[U][V];
    auto issue = [&](int b0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int vi = b0 + tid + kSbpThreads * u;
            if (vi < nvec) Vec<V>::load(src, vi, xv[u]);
        }
    };
    issue(0);
    if (tid == 0) s_key = 0ull;
    __syncthreads();
    ArgTrack<V> arg;
    arg.reset();
    for (int b0 = 0; b0 < nvec; b0 += kSbpThreads * U) {
        if (b0 > 0) issue(b0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int vi = b0 + tid + kSbpThreads * u;
            if (vi >= nvec) break;
            arg.push(xv[u]);
        }
    }
    warp_offer_argmax<V, SIG>(arg, src, nvec, (tid >> 5) * 32, kSbpThreads, P.sig_ref, &s_key);
    __syncthreads();
    if (tid == 0) {
        float best;
        int besti;
        key_to_argmax(s_key, best, besti);
        float dx = 0.0f, dy = 0.0f;
        if (P.refine && best > P.thr && besti != 0x7fffffff) {
            const int row = (int)fdiv((uint32_t)besti, P.divW);
            const int col = besti - row * P.W;
            if (col > 1 && col < P.W - 1 && row > 1 && row < P.H - 1) {
                // quarter-pixel shift toward the higher neighbour (NOT in the reference; opt-in)
                const float ddx = act<SIG>(__ldg(src + besti + 1)) - act<SIG>(__ldg(src + besti - 1));
                const float ddy = act<SIG>(__ldg(src + besti + P.W)) - act<SIG>(__ldg(src + besti - P.W));
                dx = ddx > 0.0f ? 0.25f : (ddx < 0.0f ? -0.25f : 0.0f);
                dy = ddy > 0.0f ? 0.25f : (ddy < 0.0f ? -0.25f : 0.0f);
            }
        }
        write_joint(P.joints + map * 3, best, besti, P.thr, P.scale, P.W, P.divW, dx, dy);
    }
}

// bulk-async staged form of sbp_decode_kernel (16-byte aligned maps): POSE_TMA_MPC_DEC maps per CTA, POSE_TMA_WPM_DEC warps per map,
// see sbp_fused_tma_kernel
template <bool SIG>
POSE_GLOBAL void __launch_bounds__(tma_threads(kTmaDec), tma_minb(kTmaDec)) sbp_decode_tma_kernel(SbpDecodeParams P) {
    constexpr int V = 4, MPC = tma_mpc(kTmaDec), WPM = tma_wpm(kTmaDec), TPM = 32 * WPM;
    extern __shared__ __align__(128) float tiles[];
    __shared__ __align__(8) unsigned long long s_bar[MPC];
    __shared__ unsigned long long s_key[MPC][WPM];              // one slot per warp: no atomics
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = wid / WPM, ws = wid - k * WPM, t = ws * 32 + lane;
    const long long map0 = (long long)blockIdx.x * MPC;
    const int nmap = (int)min((long long)MPC, P.n_maps - map0);
    const int nvec = P.HW / V;
    tma_stage_maps(tiles, s_bar, P.x, map0, nmap, P.HW);
    __syncthreads();
    if (k < nmap) {
        const float* tile = tiles + (size_t)k * P.HW;
        ArgTrack<V> arg;
        arg.reset();
        mbar_wait_parity(&s_bar[k], 0);
#pragma unroll 4
        for (int vi = t; vi < nvec; vi += TPM) {
            float x[V];
            Vec<V>::load_any(tile, vi, x);
            arg.push(x);
        }
        warp_offer_argmax<V, SIG, true>(arg, tile, nvec, ws * 32, TPM, P.sig_ref, &s_key[k][ws]);
    }
    __syncthreads();
    if (tid < nmap) {
        const int m = tid;
        const float* tile = tiles + (size_t)m * P.HW;
        float best;
        int besti;
        unsigned long long key = 0ull;
#pragma unroll
        for (int w = 0; w < WPM; ++w) key = max(key, s_key[m][w]);
        key_to_argmax(key, best, besti);
        float dx = 0.0f, dy = 0.0f;
        if (P.refine && best > P.thr && besti != 0x7fffffff) {
            const int row = (int)fdiv((uint32_t)besti, P.divW);
            const int col = besti - row * P.W;
            if (col > 1 && col < P.W - 1 && row > 1 && row < P.H - 1) {
                const float ddx = act<SIG>(tile[besti + 1]) - act<SIG>(tile[besti - 1]);
                const float ddy = act<SIG>(tile[besti + P.W]) - act<SIG>(tile[besti - P.W]);
                dx = ddx > 0.0f ? 0.25f : (ddx < 0.0f ? -0.25f : 0.0f);
                dy = ddy > 0.0f ? 0.25f : (ddy < 0.0f ? -0.25f : 0.0f);
            }
        }
        write_joint(P.joints + (map0 + m) * 3, best, besti, P.thr, P.scale, P.W, P.divW, dx, dy);
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
POSE_GLOBAL void __launch_bounds__(kSbpThreads) sbp_decode_flip_kernel(SbpDecodeFlipParams P) {
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
POSE_GLOBAL void __launch_bounds__(256) sbp_backproject_kernel(const float* __restrict__ joints, const double* __restrict__ bbox,
                                                              float* __restrict__ packed, int N, int K, double in_h, double in_w) {
    const int n = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (n < N) backproject_sample(joints, bbox, packed, n, K, in_h, in_w, threadIdx.x & 31);
}

// ---------------------------------------------------------------- the reference sigmoids (diagnostics)
// y[i] = sigmoid_ref(x[i]): lets a test compare the device restatements with torch.sigmoid bit for bit
POSE_GLOBAL void sigmoid_ref_eval_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned long long n, int sig_ref) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = sigmoid_ref(x[i], sig_ref);
}

// For fp32 m in (-80, +inf], every 61st float: nothing below sigmoid_window_lo(m) may reach sigmoid_ref(m) -- probed at the 64 floats just below
// the window and at 64 geometrically spaced points further down (the references are monotone up to a few ulp, so these are
// where a violation would be).  Counts violations.
POSE_GLOBAL void sigmoid_window_check_kernel(unsigned long long* violations, int sig_ref) {
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
