// tcgen05.mma issue-rate micro-benchmark: cycles per MMA for the small-N shapes of the head-fusion kernel (csrc/head_kernels.cuh).
// One CTA per SM, one thread issues `iters` MMAs (operands are whatever is in shared / tensor memory: the timing does not depend
// on the values), commits, waits; clock64 around it.  Variants: A from TMEM or from shared memory (K-major SW128 / MN-major
// SW128_32B), kind tf32 (K = 8) or bf16 (K = 16), N in {16, 32, 48, 64, 128, 256}, one or two accumulators in turn.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_bench tools/umma_bench.cu && build/umma_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#include "../pytorch-pose-estimation_b200/csrc/head_ptx.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// MODE: 0 = tf32 A in TMEM, 1 = tf32 A smem K-major, 2 = tf32 A smem MN-major (SW128_32B), 3 = bf16 A in TMEM, 4 = bf16 A smem K-major
template <int MODE, int NACC>
__global__ void __launch_bounds__(128, 1) umma_rate(int n, int m, int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 1.0f;
    if (threadIdx.x == 0) { head::mbar_init(&bar, 1); head::mbar_init_fence(); }
    if (threadIdx.x < 32) head::tmem_alloc(&s_tmem, 512);
    head::fence_proxy_async_smem();
    head::tc_fence_before();
    __syncthreads();
    head::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        const uint32_t a_s = head::smem_u32(base), b_s = head::smem_u32(base + 16384);
        constexpr bool bf16 = MODE >= 3;
        const uint32_t fmt = bf16 ? 1u : 2u;
        const uint32_t a_mn = MODE == 2 ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
        const uint64_t bdesc = head::umma_smem_desc(b_s, 16, 1024, head::kUmmaSw128);
        const uint64_t adesc = MODE == 2 ? head::umma_smem_desc(a_s, 4096, 512, head::kUmmaSw128Base32) : head::umma_smem_desc(a_s, 16, 1024, head::kUmmaSw128);
        const uint32_t at = tmem + 256;
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t d = tmem + (uint32_t)((j % NACC) * (256 / NACC));
                if (MODE == 0) head::umma_tf32_ts(d, at + (j & 3) * 8, bdesc, idesc, 1u);
                else if (MODE == 1 || MODE == 2) head::umma_tf32(d, adesc, bdesc, idesc, 1u);
                else if (MODE == 3) umma_f16_ts(d, at + (j & 3) * 8, bdesc, idesc, 1u);
                else umma_f16_ss(d, adesc, bdesc, idesc, 1u);
            }
        }
        const long long t1 = clock64();
        head::umma_commit(&bar);
        head::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        cycles[2 * blockIdx.x] = t2 - t0;
        cycles[2 * blockIdx.x + 1] = t1 - t0;
    }
    head::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) head::tmem_dealloc(tmem, 512);
}

template <int MODE, int NACC>
void run(const char* name, int sms, long long* d) {
    const int iters = 4096;
    CK(cudaFuncSetAttribute(umma_rate<MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int m : {128, 64})
        for (int n : {16, 48, 64, 128, 256}) {
            if (NACC * n > 256) continue;
            umma_rate<MODE, NACC><<<sms, 128, 64 * 1024>>>(n, m, iters, d);
            CK(cudaDeviceSynchronize());
            long long h[512];
            CK(cudaMemcpy(h, d, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost));
            long long mx = 0, mi = 0;
            for (int i = 0; i < sms; ++i) { mx = h[2 * i] > mx ? h[2 * i] : mx; mi = h[2 * i + 1] > mi ? h[2 * i + 1] : mi; }
            printf("%-32s accumulators %d  M=%3d N=%3d : %7.1f cycles per MMA to completion, %6.1f to issue\n", name, NACC, m, n, (double)mx / iters, (double)mi / iters);
        }
}

int main() {
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* d;
    CK(cudaMalloc(&d, sizeof(long long) * 2 * sms));
    run<0, 1>("tf32 A=TMEM", sms, d);
    run<0, 2>("tf32 A=TMEM", sms, d);
    run<0, 4>("tf32 A=TMEM", sms, d);
    run<1, 1>("tf32 A=smem K-major", sms, d);
    run<2, 1>("tf32 A=smem MN-major(SW128_32B)", sms, d);
    run<2, 2>("tf32 A=smem MN-major(SW128_32B)", sms, d);
    run<3, 1>("bf16 A=TMEM", sms, d);
    run<3, 2>("bf16 A=TMEM", sms, d);
    run<4, 1>("bf16 A=smem K-major", sms, d);
    return 0;
}
