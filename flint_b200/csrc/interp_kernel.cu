/* interp_kernel.cu -- batched ERK_class%Interpolate (src/ERKInterp.f90:28-140).
 *
 * Evaluates the stored dense output of every trajectory on a grid shared by the ensemble.  The phase is
 * HBM-write bound: algorithmic bytes = N*G*nI*8 written (config 4: 62.9 GB per 2^17 orbits) against
 * ~125 fp64 operations per grid point.
 *
 * Storage written by the step kernel (erk_kernel.cuh, DENSE), one contiguous run per trajectory:
 *   dxint[N][cap+1]                 Xint(1:acc+1)
 *   dbip [N][cap][pstar+1][nI]      Bip(ii, j) of every accepted step, a record of (pstar+1)*nI doubles
 *   dcount[N]                       accepted steps stored
 * so that a thread walking ONE trajectory reads whole 32-byte sectors (the first version kept trajectories
 * innermost and every coefficient read used 8 of the 32 bytes it moved: 110 GB read for 63 GB written).
 *
 * Mapping: one thread per (trajectory, chunk of the grid).  It walks its grid points in order with the
 * reference's forward cursor (:85-103: the smallest step j with Xint(j+1) >= x in the direction of
 * integration; the chunk start is found by the binary search that returns the same j), keeps the current
 * step's coefficients in registers (they serve ~G/acc consecutive points) and evaluates
 * Y = sum_k B(:,k) * theta**k exactly as InterpY (:124-140; powers in __powidf2 order, arith.cuh).
 * Layout 1 ([nI][G][N]) stores are coalesced across the warp's 32 trajectories; layout 0 ([N][G][nI], each
 * trajectory a Fortran Yarr(nI,G)) stores nI contiguous doubles per thread and point, which L2 merges into
 * full sectors as the thread moves along its row.
 */
#include "arith.cuh"
#include "kargs.h"

#define FLINT_INTERP_CHUNK 2048        /* grid points per thread block (16 KB of shared memory) */

namespace flint {

/* per-trajectory validation of the grid (:60-76): strictly monotone in the direction of
 * integration and inside [Xint(1), Xint(acc+1)] */
__global__ void interp_validate_kernel(InterpArgs a, int grid_monotone_up, int grid_monotone_down)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.N) return;
    const int acc = a.dcount[t];
    int st = FLINT_SUCCESS;
    if (acc < 1 || acc > a.cap) st = FLINT_ERROR_INTP_ARRAY;
    else {
        const double* xi = a.dxint + t * (a.cap + 1);
        const double xs = xi[0], xe = xi[acc];
        const double hs = sign1(xe - xs);
        if (a.G > 1 && ((hs > 0 && !grid_monotone_up) || (hs < 0 && !grid_monotone_down))) st = FLINT_ERROR_INTP_ARRAY;
        if (hs * (a.xarr[0] - xs) < 0.0 || hs * (a.xarr[a.G - 1] - xe) > 0.0) st = FLINT_ERROR_INTP_ARRAY;
    }
    a.status[t] = st;
}

/* ---- per-thread prefetch buffer in shared memory, filled with cp.async: the NEXT step's record travels from
 * HBM while the current one is being evaluated (~G/acc points), so a cursor advance costs a shared-memory read
 * instead of a DRAM round trip (ncu: the first use of a freshly loaded record held 1/3 of all stall samples) ---- */
template <int RS> struct RecBuf {
    static constexpr int W = (RS % 2 == 0) ? 2 : 1;            /* doubles per cp.async (16 or 8 bytes) */
    static constexpr int NQ = RS / W;
    static constexpr int BYTES = (RS + 1) * 8 * 128;          /* + one slot per thread for the step's end abscissa */
    __device__ static __forceinline__ void prefetch(double* sm, const double* __restrict__ rec, const double* __restrict__ xend)
    {
        {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + (long)RS * 128 + threadIdx.x);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(xend));
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sm + ((long)q * 128 + threadIdx.x) * W);
            if (W == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rec + q * W));
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(rec + q * W));
        }
        asm volatile("cp.async.commit_group;");
    }
    __device__ static __forceinline__ void take(const double* sm, double* B, double& xend)
    {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        xend = sm[(long)RS * 128 + threadIdx.x];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (W == 2) { const double2 v = reinterpret_cast<const double2*>(sm)[(long)q * 128 + threadIdx.x]; B[2 * q] = v.x; B[2 * q + 1] = v.y; }
            else B[q] = sm[(long)q * 128 + threadIdx.x];
        }
    }
};

template <bool FMA, int PS, int NI>
__global__ void __launch_bounds__(128) interp_kernel(InterpArgs a)
{
    constexpr int RS = (PS + 1) * NI;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long g0 = (long)blockIdx.y * a.chunk;
    long g1 = g0 + a.chunk;
    if (g1 > a.G) g1 = a.G;
    /* the block's piece of the grid in shared memory: every thread reads every point of it, and from L1 the
     * read competes with the coefficient stream (hit rate 24 %, one L2 round trip per point) */
    extern __shared__ double smem[];
    double* xs = smem;                                   /* [FLINT_INTERP_CHUNK] */
    double* rb = smem + FLINT_INTERP_CHUNK;              /* RecBuf: [RS/W][128][W] */
    for (long g = g0 + threadIdx.x; g < g1; g += blockDim.x) xs[g - g0] = a.xarr[g];
    __syncthreads();
    if (t >= a.N || g0 >= g1) return;
    if (a.status[t] != FLINT_SUCCESS) return;
    const int acc = a.dcount[t];
    const double* __restrict__ xi = a.dxint + t * (a.cap + 1);
    const double* __restrict__ rec0 = a.dbip + t * a.cap * (long)RS;
    const double hs = sign1(xi[acc] - xi[0]);
    /* chunk start: smallest j in [1,acc] (1-based) with hs*(Xint(j+1) - x) >= 0 */
    int loc;
    {
        const double x = xs[0];
        int lo = 1, hi = acc;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (hs * (xi[mid] - x) >= 0.0) hi = mid; else lo = mid + 1;
        }
        loc = lo;
    }
    double B[PS + 1][NI];
    double X0 = xi[loc - 1], X1, h;
    RecBuf<RS>::prefetch(rb, rec0 + (long)(loc - 1) * RS, xi + loc);
    RecBuf<RS>::take(rb, &B[0][0], X1);                               /* step loc: coefficients and Xint(loc+1) */
    h = X1 - X0;
    if (loc < acc) RecBuf<RS>::prefetch(rb, rec0 + (long)loc * RS, xi + loc + 1);
    for (long g = g0; g < g1; ++g) {
        const double x = xs[g - g0];
        if (loc < acc && hs * (X1 - x) < 0.0) {                       /* forward cursor (:85-103) */
            loc += 1;
            X0 = X1;
            RecBuf<RS>::take(rb, &B[0][0], X1);                       /* the record that was on its way */
            if (loc < acc && hs * (X1 - x) < 0.0) {                   /* several steps between two grid points: walk Xint */
                do { loc += 1; X0 = X1; X1 = xi[loc]; } while (loc < acc && hs * (X1 - x) < 0.0);
                RecBuf<RS>::prefetch(rb, rec0 + (long)(loc - 1) * RS, xi + loc);
                RecBuf<RS>::take(rb, &B[0][0], X1);
            }
            h = X1 - X0;
            if (loc < acc) RecBuf<RS>::prefetch(rb, rec0 + (long)loc * RS, xi + loc + 1);
        }
        const double theta = (x - X0) / h;
        double tk[PS + 1];
        theta_powers<PS>(theta, tk);
        double y[NI];
#pragma unroll
        for (int ii = 0; ii < NI; ++ii) y[ii] = 0.0;
#pragma unroll
        for (int k = 0; k <= PS; ++k)
#pragma unroll
            for (int ii = 0; ii < NI; ++ii) y[ii] = madd<FMA>(B[k][ii], tk[k], y[ii]);
        if (a.layout == 0) {
            double* __restrict__ o = a.yarr + (t * a.G + g) * NI;
            if (NI % 2 == 0) {
#pragma unroll
                for (int q = 0; q < NI / 2; ++q) reinterpret_cast<double2*>(o)[q] = make_double2(y[2 * q], y[2 * q + 1]);
            } else {
#pragma unroll
                for (int ii = 0; ii < NI; ++ii) o[ii] = y[ii];
            }
        } else {
#pragma unroll
            for (int ii = 0; ii < NI; ++ii) a.yarr[((long)ii * a.G + g) * a.N + t] = y[ii];
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <bool FMA, int PS, int NI>
static cudaError_t launch_one(const InterpArgs& a, dim3 grid, cudaStream_t s)
{
    constexpr int smem = FLINT_INTERP_CHUNK * 8 + RecBuf<(PS + 1) * NI>::BYTES;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(interp_kernel<FMA, PS, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    interp_kernel<FMA, PS, NI><<<grid, 128, smem, s>>>(a);
    return cudaGetLastError();
}
template <bool FMA, int PS>
static cudaError_t launch_ni(const InterpArgs& a, dim3 grid, cudaStream_t s)
{
    switch (a.nI) {
    case 1: return launch_one<FMA, PS, 1>(a, grid, s);
    case 2: return launch_one<FMA, PS, 2>(a, grid, s);
    case 3: return launch_one<FMA, PS, 3>(a, grid, s);
    case 4: return launch_one<FMA, PS, 4>(a, grid, s);
    case 5: return launch_one<FMA, PS, 5>(a, grid, s);
    case 6: return launch_one<FMA, PS, 6>(a, grid, s);
    default: return cudaErrorInvalidValue;
    }
}
template <bool FMA>
static cudaError_t launch_ps(const InterpArgs& a, dim3 grid, cudaStream_t s)
{
    switch (a.pstar) {                         /* DOP54 4, Verner65E 5, DOP853 7, Verner98R 8 */
    case 4: return launch_ni<FMA, 4>(a, grid, s);
    case 5: return launch_ni<FMA, 5>(a, grid, s);
    case 7: return launch_ni<FMA, 7>(a, grid, s);
    case 8: return launch_ni<FMA, 8>(a, grid, s);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_interp(const InterpArgs& a0, int up, int down, int sm_count, cudaStream_t s)
{
    InterpArgs a = a0;
    const int vb = 256;
    interp_validate_kernel<<<(unsigned)((a.N + vb - 1) / vb), vb, 0, s>>>(a, up, down);
    /* enough threads to fill the device (>= 16 warps per SM) when N alone does not: split the grid into chunks */
    const long want = (long)sm_count * 16 * 32;
    long nch = (want + a.N - 1) / a.N;
    const long min_nch = (a.G + FLINT_INTERP_CHUNK - 1) / FLINT_INTERP_CHUNK;      /* a chunk must fit the shared-memory copy of the grid */
    if (nch < min_nch) nch = min_nch;
    if (nch < 1) nch = 1;
    if (nch > a.G) nch = a.G;
    if (nch > 65535) nch = 65535;
    a.chunk = (a.G + nch - 1) / nch;
    if (a.chunk > FLINT_INTERP_CHUNK) return cudaErrorInvalidValue;
    a.nchunks = (int)((a.G + a.chunk - 1) / a.chunk);
    dim3 grid((unsigned)((a.N + 127) / 128), (unsigned)a.nchunks);
    return a.fma ? launch_ps<true>(a, grid, s) : launch_ps<false>(a, grid, s);
}

}  // namespace flint
