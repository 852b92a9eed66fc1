/* interp_eval.cuh -- ERK_class%Interpolate for the steps one CTA holds in shared memory (src/ERKInterp.f90:85-110,124-140).
 *
 * Shared by the two kernels that evaluate dense output on a grid into Yarr(nI, G) per trajectory ([N][G][nI]):
 *   erk_dense_kernel, fused mode (erk_kernel.cuh): every thread has just built the coefficients of one accepted step in
 *       registers and drops them into shared memory -- they are never written to HBM;
 *   interp_steps_kernel (interp_kernel.cu): the coefficient records of 128 consecutive steps arrive from HBM as ONE bulk copy
 *       (cp.async.bulk + mbarrier).
 *
 * STEP-MAJOR evaluation: a warp takes one step at a time and puts its lanes on the grid points INSIDE that step, so every
 * coefficient read is a shared-memory broadcast (one wavefront per 16 bytes).  Round 1's kernel put a warp's lanes on 32
 * consecutive points whatever step they fell in: with ~20 points per step (config 4) a batch straddled 2.7 steps and every
 * LDS.128 cost 2.7 wavefronts -- the L1 data pipe, at 78 %, bound the kernel (profiles/r01_notes.md).  The price is lane
 * utilisation (P points occupy ceil(P / 32) batches); the pipe that then binds is fp64, which is where a polynomial
 * evaluation belongs.
 *
 * Which step a point belongs to is the reference's forward cursor (:85-103): the smallest j with hs * (Xint(j+1) - x) >= 0,
 * never beyond the last step.  For a grid that is monotone in the direction of integration (validated per trajectory, :60-76)
 * that is a count: step j owns the points g with  #{x <= Xint(j)} <= g < #{x <= Xint(j+1)}  ("<=" in the direction of
 * integration), the first step also owns everything before it and the last everything after.
 */
#ifndef FLINT_B200_INTERP_EVAL_CUH
#define FLINT_B200_INTERP_EVAL_CUH

#include "arith.cuh"

namespace flint {

/* X lies before x in the direction of integration: hs * (X - x) < 0 of the reference's cursor test (:88) for hs = +-1.  The sign
 * of an IEEE difference of finite numbers is exact (gradual underflow), so this is one compare; NaN compares false either way */
__device__ __forceinline__ bool before_dir(double hs, double X, double x) { return hs > 0.0 ? (X < x) : (X > x); }

/* number of grid points that do not lie beyond X: the first g with before_dir(hs, X, xarr[g]), G if none */
__device__ __forceinline__ long grid_count_upto(const double* __restrict__ xarr, long G, double hs, double X)
{
    long lo = 0, hi = G;
    while (lo < hi) {
        const long mid = (lo + hi) >> 1;
        if (before_dir(hs, X, xarr[mid])) hi = mid; else lo = mid + 1;
    }
    return lo;
}

/* InterpY (:124-140) for one point of the step [X0, X1] whose coefficients B[k][ii] sit in shared memory; every lane of the warp
 * reads the same addresses */
template <bool FMA, int PS, int NI>
__device__ __forceinline__ void eval_point_bcast(double X0, double X1, double x, const double* __restrict__ B, double (&y)[NI])
{
    const double h = X1 - X0;
    const double theta = (x - X0) / h;
    double tk[PS + 1];
    theta_powers<PS>(theta, tk);
#pragma unroll
    for (int ii = 0; ii < NI; ++ii) y[ii] = 0.0;
    if (NI % 2 == 0) {                                 /* two coefficients per 16-byte broadcast */
        const double2* __restrict__ B2 = reinterpret_cast<const double2*>(B);
#pragma unroll
        for (int k = 0; k <= PS; ++k)
#pragma unroll
            for (int q = 0; q < NI / 2; ++q) {
                const double2 b = B2[k * (NI / 2) + q];
                y[2 * q] = madd<FMA>(b.x, tk[k], y[2 * q]);
                y[2 * q + 1] = madd<FMA>(b.y, tk[k], y[2 * q + 1]);
            }
    } else {
#pragma unroll
        for (int k = 0; k <= PS; ++k)
#pragma unroll
            for (int ii = 0; ii < NI; ++ii) y[ii] = madd<FMA>(B[k * NI + ii], tk[k], y[ii]);
    }
}

/* The CTA evaluates the grid points of its `ns` consecutive steps of ONE trajectory.
 *   sB   [ns][rsp]   coefficients of step i at sB + i * rsp (16-byte aligned when NI is even), B[k][ii] k-major
 *   sX   [ns + 1]    sX[i] = start of step i, sX[ns] = end of the last one
 *   sG   [ns + 1]    scratch: sG[i] .. sG[i+1] = the grid points of step i
 *   is_first / is_last: step 0 is the trajectory's first / step ns-1 its last accepted step
 *   yout = Yarr of this trajectory, [G][NI]
 * All threads of the CTA must call it (it synchronises); sB / sX must be visible (a __syncthreads before). */
template <bool FMA, int PS, int NI>
__device__ __forceinline__ void eval_block_steps(const double* __restrict__ sB, int rsp, const double* __restrict__ sX, long* __restrict__ sG,
                                                 int ns, bool is_first, bool is_last, const double* __restrict__ xarr, long G, double hs,
                                                 double* __restrict__ yout)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int i = tid; i <= ns; i += blockDim.x) {
        long g;
        if (i == 0) g = is_first ? 0 : grid_count_upto(xarr, G, hs, sX[0]);
        else if (i == ns && is_last) g = G;
        else g = grid_count_upto(xarr, G, hs, sX[i]);
        sG[i] = g;
    }
    __syncthreads();
    for (int i = warp; i < ns; i += nwarps) {
        const long g0 = sG[i], g1 = sG[i + 1];
        const double X0 = sX[i], X1 = sX[i + 1];
        const double* __restrict__ B = sB + (long)i * rsp;
        for (long gb = g0; gb < g1; gb += 32) {
            const long g = gb + lane;
            if (g < g1) {
                double y[NI];
                eval_point_bcast<FMA, PS, NI>(X0, X1, xarr[g], B, y);
                double* __restrict__ o = yout + g * NI;
                if (NI % 2 == 0) {
#pragma unroll
                    for (int q = 0; q < NI / 2; ++q) reinterpret_cast<double2*>(o)[q] = make_double2(y[2 * q], y[2 * q + 1]);
                } else {
#pragma unroll
                    for (int ii = 0; ii < NI; ++ii) o[ii] = y[ii];
                }
            }
        }
    }
}

}  // namespace flint
#endif
