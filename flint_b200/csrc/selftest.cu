/* selftest.cu -- device self-test of the branch-free division / square-root helpers (arith.cuh).
 *
 * The helpers replay nvcc's fast-path sequences, so wherever their range test passes the result
 * must equal the plain operator bit for bit.  The test draws operands from three families (raw
 * 64-bit patterns = every exponent, NaN, Inf, subnormals; "physical" magnitudes 2^-40..2^40; pairs
 * with a common denominator) and counts (a) results that differ although the range test passed and
 * (b) how often the range test passed, so a vacuous pass is visible to the caller. */
#include <cuda_runtime.h>
#include "arith.cuh"
#include "../../include/flint_b200.h"

namespace {
__device__ __forceinline__ unsigned long long mix(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double draw(unsigned long long key, int family)
{
    const unsigned long long u = mix(key);
    if (family == 0) return __longlong_as_double((long long)u);                       /* any bit pattern */
    const unsigned long long mant = u & 0x000fffffffffffffULL;
    const int e = 1023 - 40 + (int)((u >> 52) % 81);                                   /* 2^-40 .. 2^40 */
    const unsigned long long sgn = (u >> 63) << 63;
    return __longlong_as_double((long long)(sgn | ((unsigned long long)e << 52) | mant));
}
__device__ __forceinline__ bool same(double a, double b)
{
    return __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b);
}

struct PowC { double c[FLINT_NPOWC]; };
__global__ void selftest_kernel(long n, unsigned long long seed, unsigned long long* out, const __grid_constant__ PowC pc)
{
    unsigned long long bad = 0, pass_div = 0, pass_sqrt = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int family = (int)(i % 3);
        const unsigned long long k = seed ^ ((unsigned long long)i * 4ULL);
        const double x = draw(k, family), d = draw(k + 1, family), x2 = (i % 7 == 0) ? 0.0 : draw(k + 2, family);
        const double r = flint::rcp_newton(d);
        bool ok1 = true, ok2 = true, ok3 = true;
        const double q1 = flint::div_rcp(x, d, r, ok1);
        const double q2 = flint::div_rcp_z(x2, d, r, ok2);
        const double s = flint::sqrt_nb(fabs(x), ok3);
        if (ok1) { pass_div += 1; if (!same(q1, x / d)) bad += 1; }
        if (ok2) { pass_div += 1; if (!same(q2, x2 / d)) bad += 1; }
        if (ok3) { pass_sqrt += 1; if (!same(s, sqrt(fabs(x)))) bad += 1; }
        {   /* the device pow of the step-size controller against det_pow.h: error norms 1e-30..1e4 and any pattern, exponents of the controllers */
            const double px = (i % 5 == 0) ? fabs(x) : exp2(-100.0 + 114.0 * ((double)(mix(k + 7) >> 11) * (1.0 / 9007199254740992.0)));
            const double ex[6] = {0.125, 1.0 / 6.0 - 0.04 * 0.75, 0.04, 0.2, 1.0 / 9.0, -0.3};
            const double py = ex[i % 6];
            if (!same(flint::det_pow_dev(px, py, pc.c), flint_det_pow(px, py))) bad += 1;
        }
    }
    atomicAdd(&out[0], bad); atomicAdd(&out[1], pass_div); atomicAdd(&out[2], pass_sqrt);
}
}  // namespace

extern "C" int flint_b200_selftest_arith(long n, unsigned long long seed, unsigned long long* result3)
{
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, 24) != cudaSuccess) { cudaGetLastError(); return FLINT_ERROR; }
    cudaMemset(d, 0, 24);
    PowC pc;
    flint::det_pow_constants(pc.c);
    selftest_kernel<<<148 * 8, 256>>>(n, seed, d, pc);
    const cudaError_t e = cudaMemcpy(result3, d, 24, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { cudaGetLastError(); return FLINT_ERROR; }
    return FLINT_SUCCESS;
}
