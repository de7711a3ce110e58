// Fused render+loss(+grad)(+decode) with the heat maps staged through shared memory by the TMA engine.
//
// Same arithmetic and the same "one warp owns one map" decomposition as sbp_fused_kernel, but the bytes move with
// 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP) instead of per-lane LDG/STG:
//   * every warp owns a private ring of kStages x 4 KB shared-memory tiles with one mbarrier each;
//   * lane 0 keeps kStages-1 tile loads in flight (complete_tx on the tile's mbarrier) while the warp works on one
//     tile -- memory latency is hidden by the copy engine, not by resident warps, and costs no registers;
//   * the warp reads a tile with conflict-free LDS.128, writes dlogits back IN PLACE (STS.128), fences the async
//     proxy and lane 0 sends the tile to HBM with one bulk store; a tile is refilled only after the store that last
//     read it has drained (cp.async.bulk.wait_group.read).
// A tile is a 4 KB piece of one map (3 tiles per 64x48 map); per-map state (loss partials, running argmax, patch
// geometry) lives in registers across the tiles of a map.
#pragma once
#include "sbp_kernels.cuh"

namespace pose {

#ifndef POSE_TMA_STAGES
#define POSE_TMA_STAGES 3       // tiles per warp ring (stages-1 loads in flight)
#endif
#ifndef POSE_TMA_WARPS
#define POSE_TMA_WARPS 8        // warps per CTA
#endif
#ifndef POSE_TMA_TILE_VEC
#define POSE_TMA_TILE_VEC 256   // float4 per tile (multiple of 32): 256 -> 4 KB
#endif
constexpr int kTmaStages = POSE_TMA_STAGES;
constexpr int kTmaTileVec = POSE_TMA_TILE_VEC;
constexpr int kTmaTileBytes = kTmaTileVec * 16;
constexpr int kTmaWarps = POSE_TMA_WARPS;
constexpr int kTmaThreads = kTmaWarps * 32;

__host__ __device__ inline size_t sbp_tma_smem_bytes(int lut_n) {
    size_t lut = (lut_padded_floats(lut_n) * sizeof(float) + 15) / 16 * 16;
    return (size_t)kTmaWarps * kTmaStages * kTmaTileBytes + lut + (size_t)kTmaWarps * kTmaStages * sizeof(uint64_t);
}

template <bool GRAD, bool DEC>
__global__ void __launch_bounds__(kTmaThreads) sbp_fused_tma_kernel(SbpFusedParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[kTmaWarps][2];
    pdl_launch_dependents();
    if (P.xpub.world > 0 && blockIdx.x == 0 && threadIdx.x == 0) exchange_open_step(P.xpub);

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    unsigned char* tiles = smem_raw + (size_t)wid * kTmaStages * kTmaTileBytes;
    const size_t lut_bytes = (lut_padded_floats(P.lut_n) * sizeof(float) + 15) / 16 * 16;
    float* lut_s = reinterpret_cast<float*>(smem_raw + (size_t)kTmaWarps * kTmaStages * kTmaTileBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kTmaWarps * kTmaStages * kTmaTileBytes + lut_bytes) + wid * kTmaStages;

    stage_lut_padded(lut_s, P.lut, P.lut_n);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) mbar_init(smem_u32(bars + s), 1);
        mbar_init_fence();
    }
    __syncthreads();

    const long long warp0 = (long long)blockIdx.x * kTmaWarps + wid;
    const long long nwarps = (long long)gridDim.x * kTmaWarps;
    const int nvec = P.HW / 4;
    const int tiles_per_map = (nvec + kTmaTileVec - 1) / kTmaTileVec;
    const long long my_maps = P.n_maps > warp0 ? (P.n_maps - warp0 + nwarps - 1) / nwarps : 0;
    const long long total = my_maps * tiles_per_map;
    constexpr int D = kTmaStages - 1;                 // tile loads kept in flight

    // lane 0 only: start the load of this warp's t-th tile
    auto issue_load = [&](long long t) {
        const long long mi = t / tiles_per_map;
        const int c = (int)(t - mi * tiles_per_map);
        const long long map = warp0 + mi * nwarps;
        const int v0 = c * kTmaTileVec;
        const uint32_t bytes = (uint32_t)min(kTmaTileVec, nvec - v0) * 16u;
        const int b = (int)(t % kTmaStages);
        const uint32_t bar = smem_u32(bars + b);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_load(smem_u32(tiles + b * kTmaTileBytes), P.logits + map * P.HW + (long long)v0 * 4, bytes, bar);
    };
    if (lane == 0)
        for (long long t = 0; t < D && t < total; ++t) issue_load(t);

    double dpos = 0.0, dneg = 0.0;
    float apos = 0.0f, aneg = 0.0f, arem = 0.0f;
    ArgTrack<4> arg;
    arg.reset();
    Patch pt;
    long long map = warp0;
    int c = 0;                                        // tile index inside the current map
    for (long long t = 0; t < total; ++t) {
        const int b = (int)(t % kTmaStages);
        if (c == 0) {                                 // first tile of a map: per-map state
            double x, y;
            load_kp(P.kp, P.kp_f64, map, x, y);
            pt = make_patch(x, y, P.H, P.W, P.three_sigma, P.lut_n);
            apos = aneg = arem = 0.0f;
            arg.reset();
        }
        mbar_wait(smem_u32(bars + b), (uint32_t)((t / kTmaStages) & 1));
        float4* tile = reinterpret_cast<float4*>(tiles + b * kTmaTileBytes);
        const int v0 = c * kTmaTileVec;
        const int nv = min(kTmaTileVec, nvec - v0);
#pragma unroll
        for (int k = 0; k < kTmaTileVec / 32; ++k) {
            const int li = lane + 32 * k;
            if (li < nv) {
                const float4 xv4 = tile[li];
                const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
                float g[4], unused[4];
                if (DEC) arg.template push<true>(xv, v0 + li);
                render_loss_vec<4, GRAD, false>(xv, g, unused, v0 + li, pt, lut_s, P.lut_n, P.W, P.divW, P.gpos, P.gneg, apos, aneg, arem);
                if (GRAD) tile[li] = make_float4(g[0], g[1], g[2], g[3]);
            }
        }
        if (GRAD) {
            fence_async_smem();                       // my STS must be visible to the bulk store's async-proxy reads
            __syncwarp();
            if (lane == 0) {
                bulk_store(P.dlogits + map * P.HW + (long long)v0 * 4, smem_u32(tile), (uint32_t)nv * 16u);
                bulk_commit();
            }
        } else {
            __syncwarp();                             // every lane is done reading the previous tile before it is refilled
        }
        // refill the tile processed one step ago (its store, if any, must have finished reading shared memory)
        if (lane == 0 && t + D < total) {
            if (GRAD) bulk_wait_read<1>();
            issue_load(t + D);
        }
        if (++c == tiles_per_map) {                   // last tile of the map: per-map results
            finish_map<4, DEC>(P, map, lane, apos, aneg, arem, arg, dpos, dneg);
            c = 0;
            map += nwarps;
        }
    }
    if (GRAD && lane == 0) bulk_wait_all<0>();        // all tiles delivered before the CTA (and its shared memory) retires

    dpos = warp_sum(dpos);
    dneg = warp_sum(dneg);
    if (lane == 0) { red[wid][0] = dpos; red[wid][1] = dneg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, bsum = 0.0;
#pragma unroll
        for (int w = 0; w < kTmaWarps; ++w) { a += red[w][0]; bsum += red[w][1]; }
        P.partials[2 * blockIdx.x] = a;
        P.partials[2 * blockIdx.x + 1] = bsum;
    }
}

}  // namespace pose
