// fp64 pipe microbenchmark: cycles per instruction for DFMA / DMUL+DADD chains as a function of
// independent chains per thread (ILP) and warps per SM sub-partition.  Build: nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int MODE>
__global__ void k(double* out, double a, double b, int iters, long long* cyc)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) x[i] = __fma_rn(x[i], a, b);
                else if (MODE == 1) x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
                else { x[i] = __fma_rn(x[i], a, b); asm volatile("mov.b32 %0, %0;" : "+r"(iters)); }
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int MODE>
void run(int warps_per_smsp, double* out, long long* cyc)
{
    int iters = 2000;
    int threads = warps_per_smsp * 4 * 32;   // one block per SM
    k<ILP, MODE><<<148, threads>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double ninst = (double)iters * 8 * ILP * (MODE == 1 ? 2 : 1);
    printf("mode %d ILP %d warps/smsp %d: %.2f cyc per fp64 inst per warp, pipe rate %.3f inst/clk/smsp\n", MODE, ILP, warps_per_smsp,
           c / ninst, ninst * warps_per_smsp / c);
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int w : {1, 2, 4, 8}) {
        run<1, 0>(w, out, cyc); run<2, 0>(w, out, cyc); run<4, 0>(w, out, cyc); run<6, 0>(w, out, cyc); run<8, 0>(w, out, cyc);
        run<1, 1>(w, out, cyc); run<3, 1>(w, out, cyc); run<6, 1>(w, out, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
