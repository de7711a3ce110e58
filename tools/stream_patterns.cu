// Access-pattern micro-benchmark for a 1 read : 1 write fp32 stream (y = f(x), 12 288-byte "maps"): which mapping of work
// to warps / CTAs reaches which fraction of the copy peak on B200.  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/stream_patterns tools/stream_patterns.cu && build/stream_patterns
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float sig(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
template <bool MATH>
__device__ __forceinline__ float4 f4(float4 v, float& acc) {
    if (!MATH) return v;
    float4 o;
    float s;
    s = sig(v.x); acc += s * s; o.x = (s * s) * (1.f - s);
    s = sig(v.y); acc += s * s; o.y = (s * s) * (1.f - s);
    s = sig(v.z); acc += s * s; o.z = (s * s) * (1.f - s);
    s = sig(v.w); acc += s * s; o.w = (s * s) * (1.f - s);
    return o;
}

constexpr int MAPQ = 768;   // float4 per map (64 x 48 fp32)

// A: persistent, one warp per map, U loads in flight per lane (the shipped sbp_fused layout)
template <int U, bool MATH>
__global__ void __launch_bounds__(256) k_warp_map(const float4* __restrict__ x, float4* __restrict__ y, long long nmaps, float* sink) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (long long)gridDim.x * 8;
    float acc = 0.f;
    for (long long m = w0; m < nmaps; m += nw) {
        const float4* s = x + m * MAPQ;
        float4* d = y + m * MAPQ;
        for (int b = lane; b < MAPQ; b += 32 * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) if (b + 32 * u < MAPQ) v[u] = ldg_stream(s + b + 32 * u);
#pragma unroll
            for (int u = 0; u < U; ++u) if (b + 32 * u < MAPQ) __stcs(d + b + 32 * u, f4<MATH>(v[u], acc));
        }
    }
    if (acc == 123.456f) *sink = acc;
}

// B: persistent, one CTA (256 threads) per map: 3 float4 per thread = the whole map in one batch; CTAs stride over maps
template <int MPC, bool MATH>   // MPC maps per iteration and CTA (MPC*3 loads in flight per thread)
__global__ void __launch_bounds__(256) k_cta_map(const float4* __restrict__ x, float4* __restrict__ y, long long nmaps, float* sink) {
    float acc = 0.f;
    for (long long m = (long long)blockIdx.x * MPC; m < nmaps; m += (long long)gridDim.x * MPC) {
        const float4* s = x + m * MAPQ;
        float4* d = y + m * MAPQ;
        float4 v[3 * MPC];
        const long long lim = (nmaps - m) * MAPQ;
#pragma unroll
        for (int u = 0; u < 3 * MPC; ++u) if (threadIdx.x + 256 * u < lim) v[u] = ldg_stream(s + threadIdx.x + 256 * u);
#pragma unroll
        for (int u = 0; u < 3 * MPC; ++u) if (threadIdx.x + 256 * u < lim) __stcs(d + threadIdx.x + 256 * u, f4<MATH>(v[u], acc));
    }
    if (acc == 123.456f) *sink = acc;
}

// C: plain grid-stride over float4 (fully linear sweep), U per thread per iteration, persistent
template <int U, bool MATH>
__global__ void __launch_bounds__(256) k_linear(const float4* __restrict__ x, float4* __restrict__ y, long long nq, float* sink) {
    float acc = 0.f;
    const long long chunk = 256ll * U;
    for (long long base = (long long)blockIdx.x * chunk; base < nq; base += (long long)gridDim.x * chunk) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) v[u] = ldg_stream(x + base + threadIdx.x + 256 * u);
#pragma unroll
        for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) __stcs(y + base + threadIdx.x + 256 * u, f4<MATH>(v[u], acc));
    }
    if (acc == 123.456f) *sink = acc;
}

// D: non-persistent: one CTA per chunk of 256*U float4, grid = number of chunks (hardware CTA scheduler walks memory in order)
template <int U, bool MATH>
__global__ void __launch_bounds__(256) k_chunk(const float4* __restrict__ x, float4* __restrict__ y, long long nq, float* sink) {
    float acc = 0.f;
    const long long base = (long long)blockIdx.x * 256 * U;
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) v[u] = ldg_stream(x + base + threadIdx.x + 256 * u);
#pragma unroll
    for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) __stcs(y + base + threadIdx.x + 256 * u, f4<MATH>(v[u], acc));
    if (acc == 123.456f) *sink = acc;
}

// E: persistent, CONTIGUOUS range per CTA (each CTA owns nq/grid consecutive float4 and sweeps it linearly)
template <int U, bool MATH>
__global__ void __launch_bounds__(256) k_range(const float4* __restrict__ x, float4* __restrict__ y, long long nq, float* sink) {
    float acc = 0.f;
    const long long chunk = 256ll * U;
    const long long nchunks = (nq + chunk - 1) / chunk;
    const long long c0 = nchunks * blockIdx.x / gridDim.x, c1 = nchunks * (blockIdx.x + 1) / gridDim.x;
    for (long long c = c0; c < c1; ++c) {
        const long long base = c * chunk;
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) v[u] = ldg_stream(x + base + threadIdx.x + 256 * u);
#pragma unroll
        for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) __stcs(y + base + threadIdx.x + 256 * u, f4<MATH>(v[u], acc));
    }
    if (acc == 123.456f) *sink = acc;
}

// F: non-persistent, ONE CTA PER MAP (T threads x 768/T float4, all loads up front) + block reduction of a per-map sum and of a
// per-map argmax through shared memory (one barrier) + one partial per CTA: the realistic skeleton of a fused loss+decode kernel
template <int T, bool MATH, bool WRITE>
__global__ void __launch_bounds__(T) k_map_cta(const float4* __restrict__ x, float4* __restrict__ y, long long nmaps, double* part, float* sink) {
    constexpr int U = MAPQ / T;
    __shared__ float red[T / 32];
    __shared__ float redm[T / 32];
    const long long m = blockIdx.x;
    const float4* s = x + m * MAPQ;
    float4* d = y + m * MAPQ;
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldg_stream(s + threadIdx.x + T * u);
    float acc = 0.f, mx = -1e30f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        mx = fmaxf(mx, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
        const float4 o = f4<MATH>(v[u], acc);
        if (WRITE) __stcs(d + threadIdx.x + T * u, o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; redm[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = -1e30f;
#pragma unroll
        for (int w = 0; w < T / 32; ++w) { a += red[w]; b = fmaxf(b, redm[w]); }
        part[2 * m] = (double)a;
        part[2 * m + 1] = (double)b;
    }
    if (acc == 123.456f) *sink = acc;
}

// G: non-persistent, one WARP per map, 8 maps per CTA (the shipped per-warp loop, but the grid covers all maps)
template <int U, bool MATH, bool WRITE>
__global__ void __launch_bounds__(256) k_warp_map_np(const float4* __restrict__ x, float4* __restrict__ y, long long nmaps, double* part, float* sink) {
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= nmaps) return;
    float acc = 0.f;
    const float4* s = x + m * MAPQ;
    float4* d = y + m * MAPQ;
    for (int b = lane; b < MAPQ; b += 32 * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (b + 32 * u < MAPQ) v[u] = ldg_stream(s + b + 32 * u);
#pragma unroll
        for (int u = 0; u < U; ++u) if (b + 32 * u < MAPQ) { const float4 o = f4<MATH>(v[u], acc); if (WRITE) __stcs(d + b + 32 * u, o); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) part[2 * m] = (double)acc;
    if (acc == 123.456f) *sink = acc;
}

// H: non-persistent, one CTA per KCH consecutive chunks of 256*U float4, processed one after the other
template <int U, int KCH, bool MATH>
__global__ void __launch_bounds__(256) k_chunk_seq(const float4* __restrict__ x, float4* __restrict__ y, long long nq, float* sink) {
    float acc = 0.f;
    for (int k = 0; k < KCH; ++k) {
        const long long base = ((long long)blockIdx.x * KCH + k) * 256 * U;
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) v[u] = ldg_stream(x + base + threadIdx.x + 256 * u);
#pragma unroll
        for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) __stcs(y + base + threadIdx.x + 256 * u, f4<MATH>(v[u], acc));
    }
    if (acc == 123.456f) *sink = acc;
}

// J: write-only, non-persistent (one CTA per chunk) and persistent (grid-stride)
template <int U>
__global__ void __launch_bounds__(256) k_fill_chunk(float4* __restrict__ y, long long nq) {
    const long long base = (long long)blockIdx.x * 256 * U;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) __stcs(y + base + threadIdx.x + 256 * u, z);
}
template <int U>
__global__ void __launch_bounds__(256) k_fill_persist(float4* __restrict__ y, long long nq) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long chunk = 256ll * U;
    for (long long base = (long long)blockIdx.x * chunk; base < nq; base += (long long)gridDim.x * chunk) {
#pragma unroll
        for (int u = 0; u < U; ++u) if (base + threadIdx.x + 256 * u < nq) __stcs(y + base + threadIdx.x + 256 * u, z);
    }
}

template <typename F>
float timeit(F launch, int reps = 20) {
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    std::vector<float> ts;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        for (int i = 0; i < reps; ++i) launch();
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        ts.push_back(ms / reps);
    }
    CK(cudaGetLastError());
    return *std::min_element(ts.begin(), ts.end());
}

int main() {
    const long long nmaps = 4096ll * 17;
    const long long nq = nmaps * MAPQ;
    const size_t bytes = (size_t)nq * 16;
    float4 *x, *y;
    float* sink;
    CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&y, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(x, 0, bytes));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double gb = 2.0 * bytes / 1e9;
    auto report = [&](const char* name, float ms) { printf("%-52s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, gb / (ms * 1e-3)); fflush(stdout); };
    report("cudaMemcpyAsync D2D", timeit([&] { cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice, 0); }));
#define RUN(NAME, KERN, GRID, N) report(NAME, timeit([&] { KERN<<<(unsigned)(GRID), 256>>>(x, y, N, sink); }));
    for (int per = 3; per <= 4; ++per) {
        char nm[128];
        if (per > 4 && per != 6 && per != 8) continue;
        snprintf(nm, sizeof nm, "A warp/map U=6 copy, %d CTA/SM", per);        RUN(nm, (k_warp_map<6, false>), sms * per, nmaps)
        snprintf(nm, sizeof nm, "A warp/map U=6 math, %d CTA/SM", per);        RUN(nm, (k_warp_map<6, true>), sms * per, nmaps)
        snprintf(nm, sizeof nm, "A warp/map U=4 math, %d CTA/SM", per);        RUN(nm, (k_warp_map<4, true>), sms * per, nmaps)
        snprintf(nm, sizeof nm, "A warp/map U=12 math, %d CTA/SM", per);       RUN(nm, (k_warp_map<12, true>), sms * per, nmaps)
        snprintf(nm, sizeof nm, "B CTA/map x1 (3 ld/thr) math, %d CTA/SM", per); RUN(nm, (k_cta_map<1, true>), sms * per, nmaps)
        snprintf(nm, sizeof nm, "B CTA/map x2 (6 ld/thr) math, %d CTA/SM", per); RUN(nm, (k_cta_map<2, true>), sms * per, nmaps)
        snprintf(nm, sizeof nm, "C linear grid-stride U=4 math, %d CTA/SM", per); RUN(nm, (k_linear<4, true>), sms * per, nq)
        snprintf(nm, sizeof nm, "C linear grid-stride U=8 math, %d CTA/SM", per); RUN(nm, (k_linear<8, true>), sms * per, nq)
        snprintf(nm, sizeof nm, "E contiguous range/CTA U=4 math, %d CTA/SM", per); RUN(nm, (k_range<4, true>), sms * per, nq)
        snprintf(nm, sizeof nm, "E contiguous range/CTA U=8 math, %d CTA/SM", per); RUN(nm, (k_range<8, true>), sms * per, nq)
    }
    double* part;
    CK(cudaMalloc(&part, nmaps * 16));
    const double gb2 = gb;
    report("F CTA(256)/map, block reduce, R+W math", timeit([&] { k_map_cta<256, true, true><<<(unsigned)nmaps, 256>>>(x, y, nmaps, part, sink); }));
    report("F CTA(128)/map, block reduce, R+W math", timeit([&] { k_map_cta<128, true, true><<<(unsigned)nmaps, 128>>>(x, y, nmaps, part, sink); }));
    report("F CTA(192)/map, block reduce, R+W math", timeit([&] { k_map_cta<192, true, true><<<(unsigned)nmaps, 192>>>(x, y, nmaps, part, sink); }));
    report("G warp/map non-persistent U=6 R+W math", timeit([&] { k_warp_map_np<6, true, true><<<(unsigned)((nmaps + 7) / 8), 256>>>(x, y, nmaps, part, sink); }));
    report("G warp/map non-persistent U=4 R+W math", timeit([&] { k_warp_map_np<4, true, true><<<(unsigned)((nmaps + 7) / 8), 256>>>(x, y, nmaps, part, sink); }));
    report("G warp/map non-persistent U=12 R+W math", timeit([&] { k_warp_map_np<12, true, true><<<(unsigned)((nmaps + 7) / 8), 256>>>(x, y, nmaps, part, sink); }));
    report("H CTA per 2 x 12 KB sequential, math", timeit([&] { k_chunk_seq<3, 2, true><<<(unsigned)((nq + 1535) / 1536), 256>>>(x, y, nq, sink); }));
    report("H CTA per 4 x 12 KB sequential, math", timeit([&] { k_chunk_seq<3, 4, true><<<(unsigned)((nq + 3071) / 3072), 256>>>(x, y, nq, sink); }));
    report("H CTA per 8 x 12 KB sequential, math", timeit([&] { k_chunk_seq<3, 8, true><<<(unsigned)((nq + 6143) / 6144), 256>>>(x, y, nq, sink); }));
    printf("-- read-only (GB/s counts the %.0f MB read)\n", bytes / 1e6);
    {
        auto rep1 = [&](const char* name, float ms) { printf("%-52s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, bytes / 1e9 / (ms * 1e-3)); fflush(stdout); };
        rep1("F CTA(256)/map read-only math (loss+argmax)", timeit([&] { k_map_cta<256, true, false><<<(unsigned)nmaps, 256>>>(x, y, nmaps, part, sink); }));
        rep1("F CTA(128)/map read-only math", timeit([&] { k_map_cta<128, true, false><<<(unsigned)nmaps, 128>>>(x, y, nmaps, part, sink); }));
        rep1("F CTA(256)/map read-only no math (argmax)", timeit([&] { k_map_cta<256, false, false><<<(unsigned)nmaps, 256>>>(x, y, nmaps, part, sink); }));
        rep1("G warp/map non-persistent U=8 read-only math", timeit([&] { k_warp_map_np<8, true, false><<<(unsigned)((nmaps + 7) / 8), 256>>>(x, y, nmaps, part, sink); }));
        printf("-- write-only\n");
        rep1("J fill, CTA per 16 KB", timeit([&] { k_fill_chunk<4><<<(unsigned)((nq + 1023) / 1024), 256>>>(y, nq); }));
        rep1("J fill, CTA per 8 KB", timeit([&] { k_fill_chunk<2><<<(unsigned)((nq + 511) / 512), 256>>>(y, nq); }));
        rep1("J fill, CTA per 32 KB", timeit([&] { k_fill_chunk<8><<<(unsigned)((nq + 2047) / 2048), 256>>>(y, nq); }));
        rep1("J fill, persistent U=4 4 CTA/SM", timeit([&] { k_fill_persist<4><<<sms * 4, 256>>>(y, nq); }));
        rep1("J fill, persistent U=8 8 CTA/SM", timeit([&] { k_fill_persist<8><<<sms * 8, 256>>>(y, nq); }));
        rep1("cudaMemsetAsync", timeit([&] { cudaMemsetAsync(y, 0, bytes, 0); }));
    }
    (void)gb2;
    RUN("D one CTA per 16 KB chunk (U=4) math", (k_chunk<4, true>), (nq + 1023) / 1024, nq)
    RUN("D one CTA per 8 KB chunk (U=2) math", (k_chunk<2, true>), (nq + 511) / 512, nq)
    RUN("D one CTA per 4 KB chunk (U=1) math", (k_chunk<1, true>), (nq + 255) / 256, nq)
    RUN("D one CTA per 32 KB chunk (U=8) math", (k_chunk<8, true>), (nq + 2047) / 2048, nq)
    RUN("D one CTA per 16 KB chunk (U=4) copy", (k_chunk<4, false>), (nq + 1023) / 1024, nq)
    RUN("C linear grid-stride U=4 copy, 4 CTA/SM", (k_linear<4, false>), sms * 4, nq)
    return 0;
}
