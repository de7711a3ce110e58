// Multi-GPU exchange fused into the step's epilogue: the kernel that back-projects the decoded joints stores each row
// straight into EVERY rank's receive buffer over NVLink (peer-mapped symmetric memory) and the CTA that reduces the loss
// partials stores the two fp64 numerators the same way.  A one-CTA kernel on each rank then publishes "step s of rank r
// has landed" with a system-scope release store on every peer, waits (acquire) for all ranks' flags and reduces the
// gathered numerators in rank order -> the global-batch loss, bit-identical on all ranks.
// No NCCL call, no staging copy, no host synchronisation.  Receive regions form a ring of kExchangeSlots steps.  With
// defer = 0 a rank finishes step s only after every peer's step s has landed (lock-step).  With defer = 1 the wait
// kernel of step s publishes this rank's flag s but waits only for the peers' step s-1, so ranks may drift up to two
// steps apart and per-step jitter no longer adds up across GPUs; four slots make that safe: a rank can store step s
// only after its wait kernel s-1 saw every peer's flag s-2, i.e. every peer's stream is past the consumers of step s-4.
#pragma once
#include "sbp_kernels.cuh"

namespace pose {


constexpr int kMaxRowStride = 256;                 // floats per exchanged row (3K+1 padded to a multiple of 4): K <= 85

struct ExchangeDev {
    int world, rank, B, K;
    int row_stride;                                 // floats per row in the receive regions (16-byte aligned rows)
    unsigned char* peer[kMaxPeers];                 // base of every rank's exchange buffer as mapped into THIS process
    unsigned long long off_ctrl, off_flags, off_rows[kExchangeSlots], off_nums[kExchangeSlots], off_ids[kExchangeSlots];
    const long long* ids_local;                     // [B][2]
    unsigned char* mc;                              // multicast (NVLS) alias of the same buffer on ALL ranks, or nullptr
    int defer;                                      // 1: in-band mode (flags published by the next fused kernel, waits in the epilogue)
    float* loss_prev;                               // in-band mode: global loss of the previous step, written by the epilogue
    double w0, w1, inv_norm_global;                 // loss weights / normalisation for that reduction
    long long timeout_cycles;
};

// one 16-byte store replicated by the NVSwitch into every rank's copy of the buffer (multimem = NVLS multicast object)
__device__ __forceinline__ void multimem_st_v4(void* mc_addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// store 16 bytes at byte offset `off` of every rank's exchange buffer: one multicast store, or one store per peer
__device__ __forceinline__ void store_all(const ExchangeDev& X, unsigned long long off, float4 v) {
    if (X.mc) {
        multimem_st_v4(X.mc + off, v);
    } else {
        for (int r = 0; r < X.world; ++r) *reinterpret_cast<float4*>(X.peer[r] + off) = v;
    }
}

// Threads [0, world) of the calling CTA wait (acquire, system scope) until every rank's flag is >= need; all threads of the
// CTA must call it.  A peer that never signals must not hang this GPU for good: after `timeout_cycles` the wait gives up,
// records 1 + the rank it waited for in ctrl->error (sticky; PeerExchange.flush() / gathered_*() raise on it) and returns
// false -- the caller then poisons the loss of that step with NaN instead of handing out a number computed from stale rows.
__device__ __forceinline__ bool wait_for_peers(const ExchangeDev& X, ExchangeCtrl* ctrl, unsigned long long need, long long timeout_cycles) {
    __shared__ int s_timed_out;
    if (threadIdx.x == 0) s_timed_out = 0;
    __syncthreads();
    if ((int)threadIdx.x < X.world) {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(X.peer[X.rank] + X.off_flags) + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) < need) {
            if (clock64() - t0 > timeout_cycles) { atomicExch(&ctrl->error, 1u + threadIdx.x); s_timed_out = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    return s_timed_out == 0;
}

__global__ void __launch_bounds__(256) sbp_epilogue_p2p_kernel(SbpEpilogueParams P, ExchangeDev X) {
    pdl_wait();
    ExchangeCtrl* ctrl = reinterpret_cast<ExchangeCtrl*>(X.peer[X.rank] + X.off_ctrl);
    // constant while this grid runs: lock-step mode counts completed steps, in-band mode reads the step the fused
    // kernel of this call opened
    const unsigned long long step = X.defer ? ctrl->launched : ctrl->step + 1;
    const int par = (int)(step % kExchangeSlots);
    const int stride = 3 * P.K + 1;

    if ((int)blockIdx.x >= P.bp_ctas) {
        // loss partials -> local loss / numerators (two-level, fixed order: only the slice CTA that finishes last goes on),
        // then the numerators to every rank's slot [rank]
        if (!reduce_slice_and_elect(P.partials, P.n_pairs, P.slices, P.ticket, P.R, (int)blockIdx.x - P.bp_ctas)) return;
        __shared__ double nums[2];
        reduce_pairs_cta(P.slices, P.R, 2, P.w0, P.w1, P.inv_norm, P.loss_out, nums);
        __syncthreads();
        if (P.num_out && threadIdx.x == 0) { P.num_out[0] = nums[0]; P.num_out[1] = nums[1]; }
        if (threadIdx.x == 0) {
            float4 v;
            reinterpret_cast<double*>(&v)[0] = nums[0];
            reinterpret_cast<double*>(&v)[1] = nums[1];
            store_all(X, X.off_nums[par] + 16ull * X.rank, v);
        }
        if (X.defer && step >= 2) {
            // in-band completion of the PREVIOUS step: every rank published flag step-1 when its fused kernel of this
            // step started (~one kernel duration ago), so this wait normally falls straight through
            const unsigned long long need = step - 1;
            const bool ok = wait_for_peers(X, ctrl, need, X.timeout_cycles);
            const double* prev = reinterpret_cast<const double*>(X.peer[X.rank] + X.off_nums[need % kExchangeSlots]);
            reduce_pairs_cta(prev, X.world, 2, X.w0, X.w1, X.inv_norm_global, X.loss_prev, nullptr);
            __syncthreads();
            if (!ok && threadIdx.x == 0 && X.loss_prev) *X.loss_prev = __int_as_float(0x7fc00000);   // a timed-out step has no loss
        }
    } else {
        // One warp per sample.  The sample's row (K x (x_img, y_img, flag) + score, zero-padded to a 16-byte multiple) is
        // assembled in shared memory and then sent to every rank as 128-bit stores: NVLink carries a few wide, aligned
        // writes per sample instead of ~3K scalar ones.
        __shared__ __align__(16) float stage[8][kMaxRowStride];
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const int n = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
        if (n < P.N) {
            // same arithmetic as backproject_sample (SBPmAPCOCO.update_state utils/sbp_utils.py:141-163)
            const double bx = __ldg(P.bbox + 4 * n), by = __ldg(P.bbox + 4 * n + 1), bw = __ldg(P.bbox + 4 * n + 2), bh = __ldg(P.bbox + 4 * n + 3);
            const float rx = (float)(bw / P.in_w), ry = (float)(bh / P.in_h);
            const float ox = (float)bx, oy = (float)by;
            const long long grow = (long long)X.rank * X.B + n;       // row in the gathered (global, image-ordered) arrays
            float* row = stage[w];
            float sum = 0.0f;
            for (int k0 = 0; k0 < P.K; k0 += 32) {
                const int k = k0 + lane;
                float c = -1.0f;
                if (k < P.K) {
                    const float* j = P.joints + ((long long)n * P.K + k) * 3;
                    c = j[2];
                    float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f;
                    if (!(c < 0.0f)) {
                        o0 = __fadd_rn(__fmul_rn(j[0], rx), ox);
                        o1 = __fadd_rn(__fmul_rn(j[1], ry), oy);
                        o2 = 1.0f;
                    }
                    row[3 * k] = o0; row[3 * k + 1] = o1; row[3 * k + 2] = o2;
                    if (P.packed) { float* o = P.packed + (long long)n * stride + 3 * k; o[0] = o0; o[1] = o1; o[2] = o2; }
                }
                const int cnt = min(32, P.K - k0);
                for (int i = 0; i < cnt; ++i) {
                    const float ci = __shfl_sync(FULL_MASK, c, i);
                    if (!(ci < 0.0f)) sum = __fadd_rn(sum, ci);
                }
            }
            if (lane == 0) {
                const float score = __fdiv_rn(sum, (float)P.K);
                row[3 * P.K] = score;
                if (P.packed) P.packed[(long long)n * stride + 3 * P.K] = score;
            }
            if (lane >= 1 && lane < 4 && 3 * P.K + lane < X.row_stride) row[3 * P.K + lane] = 0.0f;     // padding
            __syncwarp();
            const float4* row4 = reinterpret_cast<const float4*>(row);
            for (int v = lane; v < X.row_stride / 4; v += 32) {
                store_all(X, X.off_rows[par] + 16ull * (unsigned long long)(grow * (X.row_stride / 4) + v), row4[v]);
            }
            if (lane == 0)
                store_all(X, X.off_ids[par] + 16ull * (unsigned long long)grow, *reinterpret_cast<const float4*>(X.ids_local + 2 * n));
        }
    }
    // No fence, no flag here: the grid boundary orders these stores (remote ones included) before the next grid in the
    // stream, and that grid -- exchange_wait_reduce_kernel -- publishes this rank's flag before it waits for the others.
}

// One CTA per rank.  mode 0 = finish of the lock-step mode: publish this rank's flag for step s = ctrl.step+1 (the
// epilogue grid has completed -- griddepcontrol.wait returns only after its memory operations, peer stores included,
// are performed -- so a system-scope release store suffices), wait for every rank's flag >= s, reduce the numerators
// of step s in rank order, advance the step counter.  mode 1 = flush of the in-band mode: same for the newest produced
// step (ctrl.launched), whose flag would otherwise only be published by the next fused kernel.
__global__ void __launch_bounds__(256) exchange_wait_reduce_kernel(ExchangeDev X, int mode, double w0, double w1,
                                                                   double inv_norm, float* __restrict__ loss_out, long long timeout_cycles) {
    pdl_wait();
    ExchangeCtrl* ctrl = reinterpret_cast<ExchangeCtrl*>(X.peer[X.rank] + X.off_ctrl);
    // mode 0: lock-step finish of step ctrl.step+1.  mode 1: flush of the in-band mode -- the newest produced step is
    // ctrl.launched; publish its flag (its epilogue grid has completed), wait for everybody's, reduce it.
    const unsigned long long step = mode == 0 ? ctrl->step + 1 : ctrl->launched;
    const unsigned long long need = step;
    if ((int)threadIdx.x < X.world && (mode == 0 || ctrl->step < step))
        st_release_sys(reinterpret_cast<unsigned long long*>(X.peer[threadIdx.x] + X.off_flags) + X.rank, step);
    if (need > 0) {
        const bool ok = wait_for_peers(X, ctrl, need, timeout_cycles);
        const double* nums = reinterpret_cast<const double*>(X.peer[X.rank] + X.off_nums[need % kExchangeSlots]);
        reduce_pairs_cta(nums, X.world, 2, w0, w1, inv_norm, loss_out, nullptr);
        __syncthreads();
        if (!ok && threadIdx.x == 0 && loss_out) *loss_out = __int_as_float(0x7fc00000);
    }
    if (threadIdx.x == 0) ctrl->step = step;
}

}  // namespace pose
