/* interp_kernel.cu -- batched ERK_class%Interpolate (src/ERKInterp.f90:28-140).
 *
 * Evaluates the stored dense output of every trajectory on a grid shared by the
 * ensemble.  HBM-bound: algorithmic bytes = N*G*nI*8 written (+ the coefficient
 * reads, which mostly hit L1/L2 because ~G/acc consecutive grid points fall in the
 * same step).  The reference's forward cursor search (:85-103) finds, for a
 * monotone grid, the smallest step j with Xint(j+1) >= x (in the direction of
 * integration); a per-point binary search returns the same j, which is what
 * makes the evaluation parallel over grid points.
 *
 * Storage written by the integrate kernel (erk_kernel.cuh, DENSE):
 *   dxint[(cap+1)][N]              Xint(1:acc+1)
 *   dbip [(pstar+1)][nI][cap][N]   Bip(ii, j, step)
 *   dcount[N]                      accepted steps stored
 */
#include "arith.cuh"
#include "../../include/flint_b200.h"
#include <cuda_runtime.h>

namespace flint {

struct InterpArgs {
    long N, G;
    int nI, pstar;
    long cap;
    const double* dxint; const double* dbip; const int* dcount;
    const double* xarr;
    double* yarr; int layout;        /* 0: [N][G][nI], 1: [nI][G][N] */
    int* status;                     /* [N] or NULL */
    int fma;
};

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
        const double xs = a.dxint[t], xe = a.dxint[(long)acc * a.N + t];
        const double hs = sign1(xe - xs);
        if (a.G > 1 && ((hs > 0 && !grid_monotone_up) || (hs < 0 && !grid_monotone_down))) st = FLINT_ERROR_INTP_ARRAY;
        if (hs * (a.xarr[0] - xs) < 0.0 || hs * (a.xarr[a.G - 1] - xe) > 0.0) st = FLINT_ERROR_INTP_ARRAY;
    }
    if (a.status) a.status[t] = st;
}

template <bool FMA>
__global__ void __launch_bounds__(256) interp_kernel(InterpArgs a)
{
    const long tot = a.N * a.G;
    for (long w = (long)blockIdx.x * blockDim.x + threadIdx.x; w < tot; w += (long)gridDim.x * blockDim.x) {
        long t, g;
        if (a.layout == 0) { t = w / a.G; g = w - t * a.G; }      /* lanes sweep the grid of one trajectory */
        else { g = w / a.N; t = w - g * a.N; }                    /* lanes sweep trajectories at one grid point */
        if (a.status && a.status[t] != FLINT_SUCCESS) continue;
        const int acc = a.dcount[t];
        const double x = a.xarr[g];
        const double hs = sign1(a.dxint[(long)acc * a.N + t] - a.dxint[t]);
        /* smallest j in [1,acc] (1-based) with hs*(Xint(j+1) - x) >= 0 */
        int lo = 1, hi = acc;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (hs * (a.dxint[(long)mid * a.N + t] - x) >= 0.0) hi = mid; else lo = mid + 1;
        }
        const int loc = lo;                                       /* X0loc */
        const double X0 = a.dxint[(long)(loc - 1) * a.N + t];
        const double h = a.dxint[(long)loc * a.N + t] - X0;
        const double theta = (x - X0) / h;
        double tk[9];
        {
            const double t2 = theta * theta, t4 = t2 * t2;       /* __powidf2 order, as arith.cuh */
            tk[0] = 1.0; tk[1] = theta; tk[2] = t2; tk[3] = theta * t2; tk[4] = t4; tk[5] = theta * t4;
            tk[6] = t2 * t4; tk[7] = (theta * t2) * t4; tk[8] = t4 * t4;
        }
        for (int ii = 0; ii < a.nI; ++ii) {
            double y = 0.0;
            for (int k = 0; k <= a.pstar; ++k) {
                const double b = a.dbip[(((long)k * a.nI + ii) * a.cap + (loc - 1)) * a.N + t];
                y = madd<FMA>(b, tk[k], y);
            }
            if (a.layout == 0) a.yarr[(t * a.G + g) * a.nI + ii] = y;
            else a.yarr[((long)ii * a.G + g) * a.N + t] = y;
        }
    }
}

cudaError_t launch_interp(const InterpArgs& a, int up, int down, int sm_count, cudaStream_t s)
{
    const int vb = 256;
    interp_validate_kernel<<<(unsigned)((a.N + vb - 1) / vb), vb, 0, s>>>(a, up, down);
    const long tot = a.N * a.G;
    long blocks = (tot + 255) / 256;
    const long maxb = (long)sm_count * 8;      /* grid-stride: a multiple of the SM count */
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    if (a.fma) interp_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(a);
    else interp_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace flint
