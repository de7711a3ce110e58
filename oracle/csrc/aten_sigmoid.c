/* TEST INFRASTRUCTURE ONLY -- C restatement of torch.sigmoid on a contiguous fp32 CPU tensor.
 *
 * The reference's decoders take the first row-major maximum of torch.sigmoid(x) (utils/sbp_utils.py:108-109, :73-78;
 * utils/spm_utils.py:230, :112-115).  fp32 sigmoid is many-to-one, so WHICH sigmoid decides near-ties.  The arithmetic
 * lives in third-party code that is not under /root/reference:
 *   - PyTorch (reference pins "PyTorch >= 1.8.1", README.md:12; this image: torch 2.11.0+cu128):
 *     aten/src/ATen/native/cpu/UnaryOpsKernel.cpp sigmoid_kernel, float branch: vectorised body
 *       a = Vectorized<float>(0) - a;  a = a.exp();  a = 1 + a;  a = a.reciprocal()       (reciprocal = IEEE 1/x),
 *     Vectorized<float>::exp() = Sleef_expf8_u10 (AVX2) / Sleef_expf16_u10 (AVX512)  (ATen/cpu/vec/vec256|512_float.h);
 *     elements of a tail shorter than two vectors go through the scalar lambda 1/(1+std::exp(-a)) (glibc expf).
 *   - SLEEF (bundled by PyTorch, 3.6 line): xexpf in src/libm/sleefsimdsp.c -- restated below from its published
 *     algorithm: q = rint(d/ln2); Cody-Waite reduction with L2Uf/L2Lf; degree-6 polynomial; ldexp2.
 * Pinned: oracle/check_aten_sigmoid.py compares this file with torch.sigmoid for ALL 2^32 fp32 inputs (0 mismatches on
 * torch 2.11, AVX512 and AVX2 dispatch); tests/test_oracle_selfcheck.py repeats it on a 2^24 sample on every run.
 * The device code (csrc/common.cuh: sleef_expf_u10, sigmoid_aten_cpu) follows this file operation for operation.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float pow2i(int q) { return as_float((uint32_t)(q + 0x7f) << 23); }

float pose_sleef_expf_u10(float d) {
    const int q = (int)lrintf(d * 1.442695040888963407359924681001892137426645954152985934135449406931f);
    float s, u;
    s = fmaf((float)q, -0.693145751953125f, d);
    s = fmaf((float)q, -1.428606765330187045e-06f, s);
    u = 0.000198527617612853646278381f;
    u = fmaf(u, s, 0.00139304355252534151077271f);
    u = fmaf(u, s, 0.00833336077630519866943359f);
    u = fmaf(u, s, 0.0416664853692054748535156f);
    u = fmaf(u, s, 0.166666671633720397949219f);
    u = fmaf(u, s, 0.5f);
    u = 1.0f + fmaf(s * s, u, s);
    u = (u * pow2i(q >> 1)) * pow2i(q - (q >> 1));
    if (d < -104.0f) u = 0.0f;
    if (100.0f < d) u = INFINITY;
    return u;
}

float pose_aten_sigmoid(float x) { return 1.0f / (1.0f + pose_sleef_expf_u10(-x)); }

void pose_aten_sigmoid_array(const float* x, float* y, long n) {
    for (long i = 0; i < n; ++i) y[i] = pose_aten_sigmoid(x[i]);
}
