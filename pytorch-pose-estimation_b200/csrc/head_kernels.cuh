// Head fusion: the detector's last layer -- Conv2d(C, K, 1, bias=False), models/detector/sbp.py:35-37 -- fused
// with the heat-map loss / gradient / decode that consume its output (models/loss/sbp_loss.py:29-49, utils/sbp_utils.py:56-82),
// so the logits [B,K,H,W] never exist in HBM.
//
// The 1x1 convolution is, per image, the GEMM  logits[p, k] = sum_c X[c, p] * Wt[k, c]  (p = pixel, X = features [C][H*W] in
// NCHW, i.e. PIXEL-contiguous) and the only dense contraction next to the hot path, so it runs on the 5th-generation tensor
// cores: tcgen05.mma, kind::tf32, M = 128 pixels, N = 48, K = 8 channels per instruction, accumulators in TMEM.
//
// fp32 parity with TF32 tensor cores ("3xTF32"): kind::tf32 reads the upper 19 bits of each fp32 operand, so one product loses
// 13 mantissa bits on either side (relative error 2^-11: the loss would be off by 1e-4, 10x the 1e-5 parity bar).  Both operands
// are therefore split into a tf32-exact head and a residual:
//   W  = Wh + Wl   done once per call on the device (head_split_weights_kernel): rows [0,K) of the B operand hold Wh, rows
//                  [24, 24+K) hold Wl -- both sets ride in ONE N = 48 instruction, and the epilogue adds column k and column 24+k;
//   X  = Xh + Xl   Xh is what the tensor core sees when it reads the raw fp32 tile; Xl = X - (X & 0xffffe000) is computed by four
//                  "residual" warps from the staged tile into a second shared-memory tile and multiplied by the same B operand.
// D = Xh Wh + Xh Wl (main accumulator) + Xl Wh + Xl Wl (correction accumulator), every product exact in fp32: what is left is the
// rounding of the accumulation itself (measured: tests/test_head_gpu.py compares with an fp64 convolution next to cuDNN's fp32 result).
//
// Data flow of one CTA (persistent, one per SM, whole images: the per-map reductions of loss and argmax stay inside the CTA):
//   warp 0       producer: TMA tensor loads (cp.async.bulk.tensor.4d, 128-byte swizzle with 32-byte atoms) of [32 channels x 128
//                pixels] fp32 tiles (16 KB) of X into a ring of shared-memory stages, completion on mbarriers; W (hi|lo, 48 x C) is
//                loaded once and stays in shared memory (96 KB at C = 512);
//   warps 6-13   residual warps, two groups of four on alternate stages: lane = pixel, 32 conflict-free LDS.32 bring one pixel's 32
//                channel values into registers, tcgen05.st writes them and their residuals into a ring of A tiles in TENSOR MEMORY
//                (128 lanes x 64 columns per stage), tcgen05.wait::st, fence, mbarrier arrive; the shared-memory stage is released
//                as soon as it is in registers;
//   warps 1, 14  MMA issuers (one thread each): 4 tcgen05.mma kind::tf32 (M = 128, N = 48, K = 8) per stage with A from TMEM and
//                B = a K-major SW128 descriptor on the W chunk -- raw products into the main accumulator (warp 1), residual products
//                into the correction accumulator (warp 14); tcgen05.commit frees the A slot and, after the last channel chunk,
//                hands the double-buffered accumulators to the epilogue;
//   warps 2-5    epilogue: tcgen05.ld (lane = pixel, 48 columns per accumulator), logit = D[k] + D[24+k] of both accumulators, then
//                the arithmetic of sbp_fused_kernel per (pixel, joint): sigmoid, Gaussian target from the joint's patch, loss pair,
//                dL/dlogit written as 128-byte rows (32 consecutive pixels of one map per warp store), running argmax; per image:
//                warp shuffles + a 4-way fixed-order combine -> the map's (S_pos, S_neg) fp64 pair and its joint row.
// (`tuning` bit 24 keeps the first version for comparison: residual tile written back to shared memory, MN-major A descriptors.)
// HBM traffic per image: C*H*W*4 bytes of features read once (+ K*H*W*4 of dlogits written when training) -- the logits' write
// and re-read (2 x K*H*W*4) of the unfused pair "conv kernel -> fused loss kernel" are gone.
#pragma once
#include <cuda.h>

#include "head_ptx.cuh"
#include "sbp_kernels.cuh"

namespace pose {

constexpr int kHeadM = 128;                      // pixels per tile = UMMA M
constexpr int kHeadKC = 32;                      // channels per stage = one 128-byte swizzle span of a W row
constexpr int kHeadN = 48;                       // UMMA N: rows [0,K) tf32 heads of W, rows [24,24+K) residuals
constexpr int kHeadLoRow = 24;
constexpr int kHeadMaxK = 17;                    // joints supported (17 COCO, 11 PIS); sizes the per-thread accumulators of the epilogue
constexpr int kHeadStageBytes = kHeadM * kHeadKC * 4;        // 16 KB
constexpr int kHeadWChunkBytes = kHeadN * 128;               // 6 KB: 48 rows x 32 channels
constexpr int kHeadConvGroups = 2;               // groups of four residual warps (tensor-memory variant)
constexpr int kHeadIssuer2 = 6 + 4 * kHeadConvGroups;        // warp that issues the residual MMAs (tensor-memory variant)
constexpr int kHeadThreads = (kHeadIssuer2 + 1) * 32;
constexpr int kHeadMaxStages = 8;
constexpr int kHeadAccCols = 64;                 // TMEM columns reserved per accumulator (48 used)
constexpr int kHeadTmemRing = 4;                 // A tiles (raw + residual, 64 columns each) in tensor memory
constexpr int kHeadMaxLut = 24;                  // template side supported (sigma <= 3.5)
constexpr int kHeadSmemCap = 227 * 1024;         // static + dynamic shared memory a CTA may own on sm_100

constexpr unsigned kHeadGrad = 1u, kHeadDecode = 4u, kHeadLogitsOut = 32u, kHeadNoResidual = 64u;

struct SbpHeadParams {
    int N, K, C, H, W, HW;
    int tiles_per_img, n_kc, raw_stages, lo_stages;
    FastDiv divW;
    const void* kp; int kp_f64;
    const float* lut; int lut_n; double three_sigma;       // UNPADDED n x n template
    float* dlogits; float* logits_out; float* joints;
    double* partials; unsigned int* ticket;
    float thr, scale, gpos, gneg;
    int sig_ref;
    unsigned flags;
};

// W [K][C] fp32 -> Wcat [48][C]: row k = tf32 head (round to nearest, ties away: cvt.rna.tf32), row 24+k = the residual
// W - head cut to tf32 (so nothing depends on how the tensor core treats the low 13 bits), all other rows zero.
__global__ void __launch_bounds__(256) head_split_weights_kernel(const float* __restrict__ w, float* __restrict__ wcat, int K, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kHeadN * C) return;
    const int r = i / C, c = i - r * C;
    float v = 0.0f;
    const int k = r < kHeadLoRow ? r : r - kHeadLoRow;
    if (k < K) {
        const float x = __ldg(w + (size_t)k * C + c);
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
        const float hi = __uint_as_float(hb & 0xffffe000u);
        v = r < kHeadLoRow ? hi : __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u);
    }
    wcat[i] = v;
}

__device__ __forceinline__ float head_residual(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// RT (residual through TMEM, the default): the residual warps read a staged tile ONCE from shared memory, pixel-per-thread
// (lane = pixel, 32 channel values in registers; one conflict-free 128-byte row per warp load), and store both the raw values
// and the residuals into a ring of A tiles in TENSOR MEMORY (tcgen05.st, 128 lanes x 64 columns per stage); both MMAs then take
// A from TMEM (tcgen05.mma [d], [a], b-desc: K-major by construction) and the shared-memory stage is free again as soon as the
// loads have landed in registers.  Shared-memory traffic per 16 KB tile: 16 KB TMA write + 16 KB read + 12 KB of W operand reads.
// RT = false (kept for the record, `tuning` bit 24): residuals go to a second shared-memory tile in the same swizzled layout and
// both MMAs read MN-major (transposed) A operands from shared memory -- 16 + 16 + 16 + 32 + 12 KB per tile: bound by the
// shared-memory pipe at half the HBM rate (profiles/r02_head_*).
template <bool RT>
__global__ void __launch_bounds__(kHeadThreads, 1)
sbp_head_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const SbpHeadParams P) {
    extern __shared__ uint8_t head_smem[];
    __shared__ __align__(8) uint64_t bar_full_raw[kHeadMaxStages], bar_empty_raw[kHeadMaxStages];
    __shared__ __align__(8) uint64_t bar_full_lo[kHeadMaxStages], bar_empty_lo[kHeadMaxStages];      // RT: the TMEM A ring
    __shared__ __align__(8) uint64_t bar_w, bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_lut[kHeadMaxLut * kHeadMaxLut];
    __shared__ Patch s_patch[2][kHeadMaxK];
    __shared__ float s_sum[4][kHeadMaxK][2];
    __shared__ float s_bv[4][kHeadMaxK];
    __shared__ int s_bi[4][kHeadMaxK];

    pdl_launch_dependents();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* const base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(head_smem) + 1023) & ~(uintptr_t)1023);
    uint8_t* const w_s = base;
    uint8_t* const raw_s = w_s + (size_t)P.n_kc * kHeadWChunkBytes;
    uint8_t* const lo_s = raw_s + (size_t)P.raw_stages * kHeadStageBytes;
    const bool residual = !(P.flags & kHeadNoResidual);
    const int tiles = P.tiles_per_img;
    // TMEM: columns [0,256) two buffers x (main accumulator | correction accumulator); RT: columns [256, 256 + 64*kHeadTmemRing) the A
    // ring (32 raw + 32 residual columns per stage).  The residual products go to their OWN accumulator: the small terms do not
    // suffer the rounding of the large running sum (logit error halved), and two threads can issue into the two accumulators.
    constexpr uint32_t kTmemCols = RT ? 512 : 256;
    constexpr uint32_t kARing0 = 4 * kHeadAccCols;

    if (tid == 0) {
        for (int s = 0; s < kHeadMaxStages; ++s) {
            head::mbar_init(&bar_full_raw[s], 1);
            head::mbar_init(&bar_empty_raw[s], RT ? 4 : 1);
            head::mbar_init(&bar_full_lo[s], 4);
            head::mbar_init(&bar_empty_lo[s], (RT && residual) ? 2 : 1);        // RT: one commit per issuing thread
        }
        head::mbar_init(&bar_w, 1);
        for (int b = 0; b < 2; ++b) { head::mbar_init(&bar_acc_full[b], (RT && residual) ? 2 : 1); head::mbar_init(&bar_acc_empty[b], 4); }
        head::mbar_init_fence();
        head::tma_prefetch_desc(&tmX);
        head::tma_prefetch_desc(&tmW);
        if (blockIdx.x == 0) *P.ticket = 0u;
    }
    for (int i = tid; i < P.lut_n * P.lut_n; i += kHeadThreads) s_lut[i] = __ldg(P.lut + i);
    if (warp == 1) head::tmem_alloc(&s_tmem, kTmemCols);
    head::tc_fence_before();
    __syncthreads();
    head::tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            head::mbar_arrive_expect_tx(&bar_w, (uint32_t)(P.n_kc * kHeadWChunkBytes));
            for (int c = 0; c < P.n_kc; ++c) head::tma_load_2d(w_s + (size_t)c * kHeadWChunkBytes, &tmW, &bar_w, c * kHeadKC, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int img = blockIdx.x; img < P.N; img += gridDim.x)
                for (int t = 0; t < tiles; ++t)
                    for (int kc = 0; kc < P.n_kc; ++kc) {
                        head::mbar_wait(&bar_empty_raw[s], ph ^ 1u);
                        head::mbar_arrive_expect_tx(&bar_full_raw[s], (uint32_t)kHeadStageBytes);
                        head::tma_load_4d(raw_s + (size_t)s * kHeadStageBytes, &tmX, &bar_full_raw[s], 0, kc * kHeadKC, t * (kHeadM / 32), img);
                        if (++s == P.raw_stages) { s = 0; ph ^= 1u; }
                    }
        }
    } else if (warp == 1 || warp == kHeadIssuer2) {
        // ------------------------------------------------------------ MMA issuers (one thread each)
        // What bounds an issuing thread is its own instruction stream (a lone thread on the uniform datapath: ~45 cycles per
        // tcgen05.mma in the tightest loop, tools/umma_bench.cu; 125 measured with descriptors rebuilt per instruction), not the
        // tensor pipe (N = 48: 24 cycles of work).  So the descriptors are advanced with one 32-bit add, barriers are addressed
        // by 32-bit shared addresses, and -- RT -- the raw and the residual MMAs are issued by TWO threads in different warps into
        // separate accumulators.
        const bool second = warp == kHeadIssuer2;
        if (lane == 0 && (!second || (RT && residual))) {
            head::mbar_wait(&bar_w, 0);
            head::tc_fence_after();
            constexpr uint32_t idesc_smem_a = head::umma_idesc_tf32(kHeadN, /*A MN-major*/ 1, /*B K-major*/ 0);
            constexpr uint32_t idesc_tmem_a = head::umma_idesc_tf32(kHeadN, 0, 0);
            const uint32_t w_addr = head::smem_u32(w_s), raw_addr = head::smem_u32(raw_s), lo_addr = head::smem_u32(lo_s);
            const uint32_t full_lo0 = head::smem_u32(&bar_full_lo[0]), empty_lo0 = head::smem_u32(&bar_empty_lo[0]);
            const uint32_t acc_full0 = head::smem_u32(&bar_acc_full[0]), acc_empty0 = head::smem_u32(&bar_acc_empty[0]);
            // B descriptor of W chunk 0, k-step 0 (K-major SW128: LBO 16, SBO 1024): low word carries the address, +384 per chunk, +2 per k-step
            const uint64_t bdesc0 = head::umma_smem_desc(w_addr, 16, 1024, head::kUmmaSw128);
            const uint32_t bd_hi = (uint32_t)(bdesc0 >> 32), bd_lo0 = (uint32_t)bdesc0;
            int rs = 0, ls = 0;
            uint32_t rph = 0, lph = 0;
            int it = 0;
            for (int img = blockIdx.x; img < P.N; img += gridDim.x)
                for (int t = 0; t < tiles; ++t, ++it) {
                    const int buf = it & 1;
                    head::mbar_wait_u32(acc_empty0 + 8u * buf, (((uint32_t)it >> 1) & 1u) ^ 1u);      // epilogue has drained this buffer
                    head::tc_fence_after();
                    const uint32_t d = tmem + (uint32_t)(buf * 2 * kHeadAccCols), dc = d + kHeadAccCols;
                    if (RT) {
                        const uint32_t dd = second ? dc : d;
                        const uint32_t a_off = tmem + kARing0 + (second ? 32u : 0u);
                        uint32_t bd_lo = bd_lo0;
                        uint32_t acc = 0u;
                        for (int kc = 0; kc < P.n_kc; ++kc, bd_lo += kHeadWChunkBytes / 16) {
                            head::mbar_wait_u32(full_lo0 + 8u * ls, lph);
                            head::tc_fence_after();
                            const uint32_t a_t = a_off + (uint32_t)(ls * 64);
                            head::umma_tf32_ts(dd, a_t, ((uint64_t)bd_hi << 32) | bd_lo, idesc_tmem_a, acc);
                            head::umma_tf32_ts(dd, a_t + 8, ((uint64_t)bd_hi << 32) | (bd_lo + 2), idesc_tmem_a, 1u);
                            head::umma_tf32_ts(dd, a_t + 16, ((uint64_t)bd_hi << 32) | (bd_lo + 4), idesc_tmem_a, 1u);
                            head::umma_tf32_ts(dd, a_t + 24, ((uint64_t)bd_hi << 32) | (bd_lo + 6), idesc_tmem_a, 1u);
                            head::umma_commit_u32(empty_lo0 + 8u * ls);
                            acc = 1u;
                            if (++ls == kHeadTmemRing) { ls = 0; lph ^= 1u; }
                        }
                    } else {
                        for (int kc = 0; kc < P.n_kc; ++kc) {
                            const uint32_t wb = w_addr + (uint32_t)(kc * kHeadWChunkBytes);
                            head::mbar_wait(&bar_full_raw[rs], rph);
                            head::tc_fence_after();
                            const uint32_t a0 = raw_addr + (uint32_t)(rs * kHeadStageBytes);
#pragma unroll
                            for (int ks = 0; ks < kHeadKC / 8; ++ks)
                                head::umma_tf32(d, head::umma_smem_desc(a0 + ks * 1024, 32 * 128, 512, head::kUmmaSw128Base32),
                                                head::umma_smem_desc(wb + ks * 32, 16, 1024, head::kUmmaSw128), idesc_smem_a, (uint32_t)((kc | ks) != 0));
                            if (residual) {
                                head::mbar_wait(&bar_full_lo[ls], lph);
                                head::tc_fence_after();
                                const uint32_t l0 = lo_addr + (uint32_t)(ls * kHeadStageBytes);
#pragma unroll
                                for (int ks = 0; ks < kHeadKC / 8; ++ks)
                                    head::umma_tf32(dc, head::umma_smem_desc(l0 + ks * 1024, 32 * 128, 512, head::kUmmaSw128Base32),
                                                    head::umma_smem_desc(wb + ks * 32, 16, 1024, head::kUmmaSw128), idesc_smem_a, (uint32_t)((kc | ks) != 0));
                                head::umma_commit(&bar_empty_lo[ls]);
                                if (++ls == P.lo_stages) { ls = 0; lph ^= 1u; }
                            }
                            head::umma_commit(&bar_empty_raw[rs]);
                            if (++rs == P.raw_stages) { rs = 0; rph ^= 1u; }
                        }
                    }
                    head::umma_commit_u32(acc_full0 + 8u * buf);
                }
        }
    } else if (warp >= 6 && warp < kHeadIssuer2) {
        // ------------------------------------------------------------ residual warps
        const int n_stage_img = tiles * P.n_kc;
        if (RT) {
            // lane = pixel of the tile (TMEM lane quarter = warp % 4, which is also the 32-pixel block of the staged tile)
            const int q = warp & 3;
            // staged tile: [4 pixel blocks][32 channels][128-byte rows], 32-byte chunks XOR-swizzled with (channel & 3)
            const uint32_t lane_off = (uint32_t)(q * 4096 + (lane & 7) * 4);
            const uint32_t chunk = (uint32_t)(lane >> 3);
            // Two groups of four warps take alternate stages (group = (warp - 6) / 4): what bounds this role is the latency of its
            // serial chain per stage (two mbarrier waits, 32 loads, two tcgen05.st, wait::st, fence, arrive: ~1000 cycles measured
            // with one group, ncu source view), not any pipe, so a second group doubles its rate.
            const int grp = (warp - 6) >> 2;
            int n_img = 0;
            for (int img = blockIdx.x; img < P.N; img += gridDim.x) ++n_img;
            const long long total = (long long)n_img * n_stage_img;
            int rs = grp, ls = grp;                     // stage st uses feature stage st % raw_stages and A-ring slot st % 4
            uint32_t rph = 0, lph = 0;
            if (rs >= P.raw_stages) { rs -= P.raw_stages; rph ^= 1u; }
            const uint32_t full_raw0 = head::smem_u32(&bar_full_raw[0]), empty_raw0 = head::smem_u32(&bar_empty_raw[0]);
            const uint32_t full_lo0 = head::smem_u32(&bar_full_lo[0]), empty_lo0 = head::smem_u32(&bar_empty_lo[0]);
            const uint32_t src0 = head::smem_u32(raw_s) + lane_off;
            const uint32_t a_ring = tmem + ((uint32_t)(q * 32) << 16) + kARing0;
            for (long long st = grp; st < total; st += kHeadConvGroups) {
                head::mbar_wait_u32(full_raw0 + 8u * rs, rph);
                const uint32_t src = src0 + (uint32_t)rs * kHeadStageBytes;
                uint32_t v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = head::lds_u32(src + c * 128 + ((chunk ^ (uint32_t)(c & 3)) << 5));
                head::mbar_wait_u32(empty_lo0 + 8u * ls, lph ^ 1u);
                head::tc_fence_after();
                const uint32_t a_t = a_ring + (uint32_t)(ls * 64);
                head::tmem_st_32x32(a_t, v);
                if (residual) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(head_residual(__uint_as_float(v[c])));
                    head::tmem_st_32x32(a_t + 32, v);
                }
                head::tmem_st_wait();
                head::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    head::mbar_arrive_u32(full_lo0 + 8u * ls);
                    head::mbar_arrive_u32(empty_raw0 + 8u * rs);         // the tile is in TMEM: the stage may be refilled
                }
                rs += kHeadConvGroups;
                if (rs >= P.raw_stages) { rs -= P.raw_stages; rph ^= 1u; }
                ls += kHeadConvGroups;
                if (ls >= kHeadTmemRing) { ls -= kHeadTmemRing; lph ^= 1u; }
            }
        } else if (residual && warp < 10) {
            const int ct = tid - 6 * 32;
            int rs = 0, ls = 0;
            uint32_t rph = 0, lph = 0;
            for (int img = blockIdx.x; img < P.N; img += gridDim.x)
                for (int st = 0; st < n_stage_img; ++st) {
                    head::mbar_wait(&bar_full_raw[rs], rph);
                    head::mbar_wait(&bar_empty_lo[ls], lph ^ 1u);
                    const float4* src = reinterpret_cast<const float4*>(raw_s + (size_t)rs * kHeadStageBytes);
                    float4* dst = reinterpret_cast<float4*>(lo_s + (size_t)ls * kHeadStageBytes);
                    float4 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = src[ct + 128 * i];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 r;
                        r.x = head_residual(v[i].x); r.y = head_residual(v[i].y); r.z = head_residual(v[i].z); r.w = head_residual(v[i].w);
                        dst[ct + 128 * i] = r;
                    }
                    head::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) head::mbar_arrive(&bar_full_lo[ls]);
                    if (++rs == P.raw_stages) { rs = 0; rph ^= 1u; }
                    if (++ls == P.lo_stages) { ls = 0; lph ^= 1u; }
                }
        }
    } else if (warp >= 2 && warp < 6) {
        // ------------------------------------------------------------ epilogue warps 2-5: TMEM lane quarter = warp % 4
        const int q = warp & 3;
        const int et = q * 32 + lane;                  // row of the tile = pixel within the tile (also the index among the 128 epilogue threads)
        const bool want_grad = P.flags & kHeadGrad, want_dec = P.flags & kHeadDecode, want_logits = P.flags & kHeadLogitsOut;
        const int K = P.K, HW = P.HW;
        int it = 0, ii = 0;
        for (int img = blockIdx.x; img < P.N; img += gridDim.x, ++ii) {
            const int pb = ii & 1;
            if (et < K) {
                double kx = -1.0, ky = -1.0;                    // no keypoints (decode-only call): every joint invisible, the target is zero
                if (P.kp) load_kp(P.kp, P.kp_f64, (long long)img * K + et, kx, ky);
                s_patch[pb][et] = make_patch(kx, ky, P.H, P.W, P.three_sigma, P.lut_n);
            }
            head::bar_sync(1, 128);
            // loss sums: every (tile, joint) contribution is summed over the warp at once and kept by lane k (two registers per
            // thread instead of 2 x 17); the running argmax stays per thread (the pixels of a thread come in increasing order)
            float my_pos = 0.0f, my_neg = 0.0f, bv[kHeadMaxK];
            int bi[kHeadMaxK];
#pragma unroll
            for (int k = 0; k < kHeadMaxK; ++k) { bv[k] = -INFINITY; bi[k] = 0x7fffffff; }
            for (int t = 0; t < tiles; ++t, ++it) {
                const int buf = it & 1;
                head::mbar_wait(&bar_acc_full[buf], ((uint32_t)it >> 1) & 1u);
                head::tc_fence_after();
                {
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * kHeadAccCols);
                    uint32_t r0[32], r1[16];
                    float xs[kHeadMaxK];
                    auto fold = [&](bool first) {              // xs[k] (+)= D[k] + D[24 + k]
#pragma unroll
                        for (int k = 0; k < kHeadMaxK; ++k) {
                            const float lo = __uint_as_float(k + kHeadLoRow < 32 ? r0[(k + kHeadLoRow) & 31] : r1[(k + kHeadLoRow - 32) & 15]);
                            const float v = __uint_as_float(r0[k]) + lo;
                            xs[k] = first ? v : xs[k] + v;
                        }
                    };
                    if (residual) {                            // the correction accumulator first: small terms are summed before the large one
                        head::tmem_ld_32x32(taddr + kHeadAccCols, r0);
                        head::tmem_ld_32x16(taddr + kHeadAccCols + 32, r1);
                        head::tmem_ld_wait();
                        fold(true);
                    }
                    head::tmem_ld_32x32(taddr, r0);
                    head::tmem_ld_32x16(taddr + 32, r1);
                    head::tmem_ld_wait();
                    fold(!residual);
                    head::tc_fence_before();                   // the accumulator is in registers: give the buffer back to the MMA warp
                    __syncwarp();
                    if (lane == 0) head::mbar_arrive(&bar_acc_empty[buf]);
                    const int p = t * kHeadM + et;
                    const int row = (int)fdiv((uint32_t)p, P.divW), col = p - row * P.W;
                    const size_t o0 = (size_t)img * K * HW + p;
#pragma unroll
                    for (int k = 0; k < kHeadMaxK; ++k) {
                        if (k < K) {
                            const float x = xs[k];
                            const Patch& pt = s_patch[pb][k];
                            float tt = 0.0f;
                            if (row >= pt.py0 && row < pt.py1 && col >= pt.px0 && col < pt.px1) tt = s_lut[(row - pt.uly) * P.lut_n + (col - pt.ulx)];
                            const float s = sigmoid_fast(x);
                            float cp = 0.0f, cn = 0.0f;
                            const float g = loss_elem<true>(s, tt, P.gpos, P.gneg, cp, cn);
                            cp = warp_sum(cp);
                            cn = warp_sum(cn);
                            if (lane == k) { my_pos += cp; my_neg += cn; }
                            if (want_grad) __stcs(P.dlogits + o0 + (size_t)k * HW, g);
                            if (want_logits) __stcs(P.logits_out + o0 + (size_t)k * HW, x);
                            const float xc = x > 17.5f ? 17.5f : x;       // sigmoid is exactly 1 in fp32 from ~17.33 on: a plateau, first index wins
                            if (xc > bv[k]) { bv[k] = xc; bi[k] = p; }
                        }
                    }
                }
            }
            // ---- per image: combine the 128 pixel threads (warp shuffles, then the 4 warps in fixed order)
#pragma unroll
            for (int k = 0; k < kHeadMaxK; ++k) {
                if (k < K) {
                    float v = bv[k];
                    int i = bi[k];
                    if (want_dec) warp_argmax_first(v, i);
                    if (lane == 0) { s_bv[q][k] = v; s_bi[q][k] = i; }
                }
            }
            if (lane < K) { s_sum[q][lane][0] = my_pos; s_sum[q][lane][1] = my_neg; }
            head::bar_sync(1, 128);
            if (et < K) {
                double a = 0.0, b = 0.0;
                float v = -INFINITY;
                int i = 0x7fffffff;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    a += (double)s_sum[w][et][0];
                    b += (double)s_sum[w][et][1];
                    const float ov = s_bv[w][et];
                    const int oi = s_bi[w][et];
                    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
                }
                reinterpret_cast<double2*>(P.partials)[(long long)img * K + et] = make_double2(a, b);
                if (want_dec) {
                    const float conf = i != 0x7fffffff ? sigmoid_ref(v, P.sig_ref) : -INFINITY;
                    write_joint(P.joints + ((long long)img * K + et) * 3, conf, i, P.thr, P.scale, P.W, P.divW);
                }
            }
            // (s_sum / s_bv / s_bi are next written after the following image's first bar_sync: no third barrier needed)
        }
    }
    head::tc_fence_before();
    __syncthreads();
    __syncwarp();
    if (warp == 1) head::tmem_dealloc(tmem, kTmemCols);
}

}  // namespace pose
