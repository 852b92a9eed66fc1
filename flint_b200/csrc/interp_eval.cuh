/* interp_eval.cuh -- ERK_class%Interpolate for the steps one CTA holds in shared memory (src/ERKInterp.f90:85-110,124-140).
 *
 * Shared by the two kernels that evaluate dense output on a grid into Yarr(nI, G) per trajectory ([N][G][nI]):
 *   erk_dense_kernel, fused mode (erk_kernel.cuh): every thread has just built the coefficients of one accepted step in
 *       registers and drops them into shared memory -- they are never written to HBM;
 *   interp_steps_kernel (interp_kernel.cu): the coefficient records of 128 consecutive steps arrive from HBM as ONE bulk copy
 *       (cp.async.bulk + mbarrier).
 *
 * The evaluation is point-major within the block's steps (see eval_block_steps).  Strict arithmetic costs ~113 fp64
 * instructions per 48 bytes written, which puts the fp64 pipe and HBM within 15 % of each other: the phase is co-bound, and what
 * decides the kernel is how few other instructions and shared-memory wavefronts accompany them.
 *
 * Which step a point belongs to is the reference's forward cursor (:85-103): the smallest j with hs * (Xint(j+1) - x) >= 0,
 * never beyond the last step.  For a grid that is monotone in the direction of integration (validated per trajectory, :60-76)
 * that is a count: step j owns the points g with  #{x <= Xint(j)} <= g < #{x <= Xint(j+1)}  ("<=" in the direction of
 * integration), the first step also owns everything before it and the last everything after.
 */
#ifndef FLINT_B200_INTERP_EVAL_CUH
#define FLINT_B200_INTERP_EVAL_CUH

#include "arith.cuh"
#include "kargs.h"

namespace flint {

/* X lies before x in the direction of integration: hs * (X - x) < 0 of the reference's cursor test (:88) for hs = +-1.  The sign
 * of an IEEE difference of finite numbers is exact (gradual underflow), so this is one compare; NaN compares false either way */
__device__ __forceinline__ bool before_dir(double hs, double X, double x) { return hs > 0.0 ? (X < x) : (X > x); }

/* number of grid points that do not lie beyond X: the first g with before_dir(hs, X, xarr[g]), G if none.  The grid is monotone
 * in the direction hs, so that is a bisection -- started from where X would sit in a uniform grid and widened by doubling, so that
 * a (nearly) uniform grid costs three or four dependent loads instead of log2(G) */
__device__ __forceinline__ long grid_count_upto(const double* __restrict__ xarr, long G, double hs, double X)
{
    if (G <= 0) return 0;
    const double xa = xarr[0], xb = xarr[G - 1];
    long lo = 0, hi = G;                                               /* answer in [lo, hi]; points below lo are not beyond X, point hi is */
    if (before_dir(hs, X, xa)) return 0;
    if (!before_dir(hs, X, xb)) return G;
    {
        const double fr = (X - xa) / (xb - xa);                        /* in [0, 1] here; a guess only */
        long c = (long)(fr * (double)(G - 1));
        c = c < 0 ? 0 : (c > G - 1 ? G - 1 : c);
        long w = 1;
        if (before_dir(hs, X, xarr[c])) {                              /* point c lies beyond X: the answer is <= c; walk down */
            hi = c;
            for (;;) {
                const long q = hi - w;
                if (q <= 0) { lo = 0; break; }
                if (before_dir(hs, X, xarr[q])) { hi = q; w <<= 1; } else { lo = q + 1; break; }
            }
        } else {                                                       /* the answer is > c; walk up */
            lo = c + 1;
            for (;;) {
                const long q = lo + w - 1;
                if (q >= G - 1) { hi = G - 1; break; }                 /* point G-1 lies beyond X (checked above) */
                if (before_dir(hs, X, xarr[q])) { hi = q; break; } else { lo = q + 1; w <<= 1; }
            }
        }
    }
    while (lo < hi) {
        const long mid = (lo + hi) >> 1;
        if (before_dir(hs, X, xarr[mid])) hi = mid; else lo = mid + 1;
    }
    return lo;
}

/* the same count when it can be had from ONE round of loads: the three grid points around the uniform-grid guess, fetched
 * together.  Split in two so that a caller can do other work (or start other searches) while the loads travel: grid_probe_issue
 * computes the guess and issues the loads, grid_probe_result decides; ok = false: not decided, call grid_count_upto.
 * xa = xarr[0], xb = xarr[G-1], inv_span = 1 / (xb - xa) (once per caller). */
struct GridProbe { double X, q0, q1, q2; long c; };
__device__ __forceinline__ void grid_probe_issue(GridProbe& p, const double* __restrict__ xarr, long G, double X, double xa, double inv_span)
{
    const double fr = (X - xa) * inv_span;                                    /* inv_span = 1 / (xb - xa): a guess only, verified by the loads */
    long c = (fr >= 0.0 && fr <= 1.0) ? (long)(fr * (double)(G - 1)) : 0;      /* NaN and out-of-range guesses land on 0 */
    c = c > G - 1 ? G - 1 : c;
    const long cm = c > 0 ? c - 1 : 0, cp = c < G - 1 ? c + 1 : G - 1;
    p.X = X; p.c = c;
    p.q0 = xarr[cm]; p.q1 = xarr[c]; p.q2 = xarr[cp];
}
__device__ __forceinline__ long grid_probe_result(const GridProbe& p, long G, double hs, double xa, double xb, bool& ok)
{
    const bool first = before_dir(hs, p.X, xa), none = !before_dir(hs, p.X, xb);
    const bool b0 = before_dir(hs, p.X, p.q0), b1 = before_dir(hs, p.X, p.q1), b2 = before_dir(hs, p.X, p.q2);
    long res = 0;
    bool done = false;
    if (b1) { if (p.c == 0) { res = 0; done = true; } else if (!b0) { res = p.c; done = true; } }
    else if (p.c < G - 1 && b2) { res = p.c + 1; done = true; }
    if (none) { res = G; done = true; }
    if (first) { res = 0; done = true; }
    ok = done;
    return res;
}

/* The CTA evaluates the grid points of its `ns` (<= kEvalSteps) consecutive steps of ONE trajectory.
 *   sB   [ns][rsp]   coefficients of step i at sB + i * rsp (16-byte aligned when NI is even), B[k][ii] k-major
 *   sX   [ns + 1]    sX[i] = start of step i, sX[ns] = end of the last one
 *   sG   [ns + 1]    scratch: sG[i] .. sG[i+1] = the grid points of step i
 *   is_first / is_last: step 0 is the trajectory's first / step ns-1 its last accepted step
 *   yout = Yarr of this trajectory: [G][NI], or with SOA the trajectory's column of [NI][G][N] (element stride soa_n)
 * Every warp takes a contiguous share of the steps and walks their points in order, 32 consecutive points per batch (coalesced
 * loads of x, contiguous stores), but a batch never reaches into a THIRD step: its lanes read the coefficients of at most two
 * records, so a 16-byte read is at most two shared-memory wavefronts (round 1's kernel let a batch straddle 2.7 steps on
 * average and was bound by the L1 data pipe at 78 %), and with ~20 points per step the lanes stay ~90 % full.  Measured
 * alternatives (profiles/r02_notes.md): one step per batch (65 % lanes) and one THREAD per step with its coefficients in
 * registers, steps ordered by point count (uncoalesced stores: 1.5x slower).
 * block_step_ranges (all threads, synchronises) fills sG from sX; eval_block_steps then needs sB, sX and sG visible. */
constexpr int kEvalSteps = 128;
/* sG[i] .. sG[i+1] = the grid points of step i, for the block's ns steps (all threads; ends with a __syncthreads) */
__device__ __forceinline__ void block_step_ranges(const double* __restrict__ sX, long* __restrict__ sG, int ns, bool is_first, bool is_last,
                                                  const double* __restrict__ xarr, long G, double hs)
{
    for (int i = threadIdx.x; i <= ns; i += blockDim.x) {
        long g;
        if (i == 0) g = is_first ? 0 : grid_count_upto(xarr, G, hs, sX[0]);
        else if (i == ns && is_last) g = G;
        else g = grid_count_upto(xarr, G, hs, sX[i]);
        sG[i] = g;
    }
    __syncthreads();
}
/* one warp evaluates the grid points [gA, gB) of the block's steps (gA >= sG[0], gB <= sG[ns]) */
template <bool FMA, int PS, int NI, bool SOA>
__device__ __forceinline__ void eval_point_range(const double* __restrict__ sB, int rsp, const double* __restrict__ sX, const long* __restrict__ sG,
                                                 int ns, const double* __restrict__ xarr, long G, double* __restrict__ yout, long soa_n,
                                                 long gA, long gB, int lane)
{
    if (gA >= gB) return;
    int i = 0;
    {                                                                  /* the step that holds point gA: largest i with sG[i] <= gA */
        int lo = 0, hi = ns - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sG[mid] <= gA) lo = mid; else hi = mid - 1; }
        i = lo;
    }
    const int iend = ns;
    long g = gA;
    const long gend = gB;
    double xn = g + lane < gend ? xarr[g + lane] : 0.0;                /* the next batch's abscissae travel one iteration ahead */
    while (g < gend) {                                                 /* warp-uniform */
        while (sG[i + 1] <= g) ++i;                                    /* the step that holds point g (empty steps are skipped) */
        const long mid = sG[i + 1];                                    /* first point of step i + 1 */
        const long lim = i + 1 < iend ? sG[i + 2] : mid;               /* a batch ends with step i + 1 at the latest */
        long n = lim - g < 32 ? lim - g : 32;
        n = gend - g < n ? gend - g : n;
        const long gl = g + lane;
        const double x = xn;
        { const long gq = g + n + lane; xn = gq < gend ? xarr[gq] : 0.0; }
        if (lane < n) {
            const int ii = i + (gl >= mid ? 1 : 0);
            FLINT_CHK(ii >= 0 && ii < ns && gl >= 0 && gl < G);
            const double X0 = sX[ii], X1 = sX[ii + 1];
            const double* __restrict__ B = sB + (long)ii * rsp;
            const double h = X1 - X0;
            const double theta = (x - X0) / h;                         /* InterpY, :124-140 */
            double tk[PS + 1], y[NI];
            theta_powers<PS>(theta, tk);
#pragma unroll
            for (int q = 0; q < NI; ++q) y[q] = 0.0;
            if (NI % 2 == 0) {                                         /* two coefficients per 16-byte read */
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
                    for (int q = 0; q < NI; ++q) y[q] = madd<FMA>(B[k * NI + q], tk[k], y[q]);
            }
            if (SOA) {
#pragma unroll
                for (int q = 0; q < NI; ++q) yout[((long)q * G + gl) * soa_n] = y[q];
            } else {
                double* __restrict__ o = yout + gl * NI;
                if (NI % 2 == 0) {
#pragma unroll
                    for (int q = 0; q < NI / 2; ++q) reinterpret_cast<double2*>(o)[q] = make_double2(y[2 * q], y[2 * q + 1]);
                } else {
#pragma unroll
                    for (int q = 0; q < NI; ++q) o[q] = y[q];
                }
            }
        }
        g += n;
    }
}

/* warp `warp` of `nwarps` takes an equal share of the block's POINTS (steps hold 0 .. ~60 points each: equal shares of the steps
 * leave the warps of a block unevenly loaded) */
template <bool FMA, int PS, int NI, bool SOA>
__device__ __forceinline__ void eval_steps_share(const double* __restrict__ sB, int rsp, const double* __restrict__ sX, const long* __restrict__ sG,
                                                 int ns, const double* __restrict__ xarr, long G, double* __restrict__ yout, long soa_n,
                                                 int warp, int nwarps, int lane)
{
    const long g0 = sG[0], g1 = sG[ns];
    const long per = (g1 - g0 + nwarps - 1) / nwarps;
    long gA = g0 + warp * per, gB = gA + per;
    if (gB > g1) gB = g1;
    eval_point_range<FMA, PS, NI, SOA>(sB, rsp, sX, sG, ns, xarr, G, yout, soa_n, gA, gB, lane);
}
template <bool FMA, int PS, int NI, bool SOA>
__device__ __forceinline__ void eval_block_steps(const double* __restrict__ sB, int rsp, const double* __restrict__ sX, const long* __restrict__ sG,
                                                 int ns, const double* __restrict__ xarr, long G, double* __restrict__ yout, long soa_n)
{
    eval_steps_share<FMA, PS, NI, SOA>(sB, rsp, sX, sG, ns, xarr, G, yout, soa_n, threadIdx.x >> 5, blockDim.x >> 5, threadIdx.x & 31);
}

}  // namespace flint
#endif
