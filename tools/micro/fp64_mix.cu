// Does a non-fp64 instruction issue in the shadow of a DFMA?  8 independent DFMA chains per thread plus K independent
// 32-bit instructions per DFMA (ALU: LOP3, FMA pipe: IMAD), 4 warps per sub-partition.  If the fp64 pipe takes one warp
// instruction per 2 cycles and the scheduler issues 1 instruction per cycle, K <= 1 should be free.
// Build: nvcc -arch=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KNUM, int KDEN, int KIND>   // KNUM/KDEN extra instructions per DFMA
__global__ void k(double* out, double a, double b, int iters, int c1, int c2, long long* cyc)
{
    double x[8];
    int y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a + i + threadIdx.x; y[i] = c1 + i + threadIdx.x; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = __fma_rn(x[i], a, b);
                const int n_before = (u * 8 + i) * KNUM / KDEN, n_after = (u * 8 + i + 1) * KNUM / KDEN;
#pragma unroll
                for (int j = n_before; j < n_after; ++j) {
                    if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j % 8]) : "r"(c1), "r"(c2));
                    else if (KIND == 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(y[j % 8]) : "r"(c1), "r"(c2));
                    else asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; selp.b32 %0, %0, %1, p;}" : "+r"(y[j % 8]) : "r"(c1), "r"(c2));
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
    int z = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; z += y[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + z;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KNUM, int KDEN, int KIND>
void run(int w, double* out, long long* cyc)
{
    int iters = 2000;
    k<KNUM, KDEN, KIND><<<148, w * 128>>>(out, 1.0000001, 1e-9, iters, 12345, 777, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double nd = (double)iters * 32 * w;
    printf("kind %d extra %d/%d per DFMA, warps/smsp %d: %.3f cycles per DFMA per sub-partition (2.0 = pipe peak); issue rate %.3f/clk\n", KIND, KNUM, KDEN, w, c / nd,
           nd * (1.0 + (double)KNUM / KDEN) / c);
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int w : {4, 8}) {
        run<0, 1, 0>(w, out, cyc);
        run<1, 2, 0>(w, out, cyc); run<1, 1, 0>(w, out, cyc); run<3, 2, 0>(w, out, cyc); run<2, 1, 0>(w, out, cyc);
        run<1, 2, 1>(w, out, cyc); run<1, 1, 1>(w, out, cyc); run<3, 2, 1>(w, out, cyc); run<2, 1, 1>(w, out, cyc);
        run<1, 1, 2>(w, out, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
