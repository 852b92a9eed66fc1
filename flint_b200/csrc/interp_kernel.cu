/* interp_kernel.cu -- batched ERK_class%Interpolate (src/ERKInterp.f90:28-140) from stored coefficient records.
 *
 * Evaluates the dense output of every trajectory on a grid (shared by the ensemble, or one grid per trajectory).  The phase
 * writes N*G*nI*8 bytes (config 4: 62.9 GB per 2^17 orbits) against ~113 (strict) / ~70 (FMA) fp64 instructions per grid
 * point, and reads every coefficient record once.
 *
 * Storage (kargs.h: DenseView): one record per accepted step, X0, X1, Bip(ii, j) as [j][ii]; the records of a trajectory are
 * contiguous within a capacity level.  Every grid point gets the reference's step (forward cursor of :85-103) and
 * Y = sum_k B(:,k) * theta**k exactly as InterpY (:124-140; powers in __powidf2 order, arith.cuh).
 *
 *   interp_steps_kernel   [N][G][nI] (each trajectory a Fortran Yarr(nI,G)); also [nI][G][N] for small ensembles and grown
 *                         stores.  A CTA = 128 consecutive steps of one trajectory: their records arrive as ONE bulk copy
 *                         (cp.async.bulk, completion on an mbarrier) and are evaluated step-major (interp_eval.cuh).
 *   interp_kernel_soa     [nI][G][N] with >= 4096 trajectories: one THREAD per trajectory, lanes on consecutive trajectories
 *                         -> coalesced stores, coefficients in registers.
 * The fused path (Integrate's step logs -> coefficients in registers -> Yarr, no records at all) is erk_dense_kernel
 * (erk_kernel.cuh); it shares the step-major evaluation.
 */
#include <cstdint>
#include "arith.cuh"
#include "kargs.h"
#include "interp_eval.cuh"

namespace flint {

__device__ __forceinline__ bool before(double hs, double X, double x) { return before_dir(hs, X, x); }

/* per-trajectory validation of the grid (:60-76): strictly monotone in the direction of integration and inside
 * [Xint(1), Xint(acc+1)].  Works on any kind of record (logs or coefficients).  A shared grid's monotonicity is checked once on
 * the host; a per-trajectory grid is checked here. */
__global__ void interp_validate_kernel(InterpArgs a, int grid_monotone_up, int grid_monotone_down)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.N) return;
    const long acc = a.dcount[t];
    int st = FLINT_SUCCESS;
    if (acc < 1 || acc > a.dv.capacity()) st = FLINT_ERROR_INTP_ARRAY;       /* nothing stored / steps were dropped */
    else {
        const double xs = a.dv.rec(t, 0)[0], xe = dense_x1(a.dv.rec(t, acc - 1), a.dv.rstride);
        const double hs = sign1(xe - xs);
        const double* __restrict__ xg = a.xarr + (a.per_traj ? t * a.G : 0L);
        int up = grid_monotone_up, down = grid_monotone_down;
        if (a.per_traj) {
            up = 1; down = 1;
            for (long i = 0; i + 1 < a.G; ++i) { if (!(xg[i] < xg[i + 1])) up = 0; if (!(xg[i] > xg[i + 1])) down = 0; }
        }
        if (a.G > 1 && ((hs > 0 && !up) || (hs < 0 && !down))) st = FLINT_ERROR_INTP_ARRAY;
        if (hs * (xg[0] - xs) < 0.0 || hs * (xg[a.G - 1] - xe) > 0.0) st = FLINT_ERROR_INTP_ARRAY;
    }
    a.status[t] = st;
}
cudaError_t launch_interp_validate(const InterpArgs& a, int up, int down, cudaStream_t s)
{
    const int vb = 256;
    interp_validate_kernel<<<(unsigned)((a.N + vb - 1) / vb), vb, 0, s>>>(a, up, down);
    return cudaGetLastError();
}

/* ======================================================================================================
 * Two mappings, one per output layout (both measured on 2^16 orbits x 10^4 points, Verner98R, 6 states):
 *   [N][G][nI]  (each trajectory a Fortran Yarr(nI,G)): one WARP per trajectory, lanes on consecutive grid
 *               points -> contiguous stores, shared coefficient reads            12.4 ms (2.5 TB/s written)
 *   [nI][G][N]  (structure of arrays): one THREAD per trajectory, lanes on consecutive trajectories ->
 *               coalesced stores, coefficients in registers                      14.3 ms (2.2 TB/s written)
 * (the warp-per-trajectory kernel writing [nI][G][N] through a shared-memory transpose: 20 ms).
 * ====================================================================================================== */
#define FLINT_INTERP_CHUNK 2048        /* thread-per-trajectory kernel: grid points per block (16 KB of shared memory) */

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
__global__ void __launch_bounds__(128) interp_kernel_soa(InterpArgs a)
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
    /* record of step j (1-based) = a.dv.rec(t, j - 1): X0, X1, coefficients; Xint(j+1) = X1 of record j-1.  Nearly every
     * trajectory lives in level 0 of the store; the few that grew deeper levels pay a short loop per record */
    auto recp = [&](int j) -> const double* { return a.dv.rec(t, (long)j - 1); };
    auto xi_at = [&](int j) -> double { return j == 0 ? recp(1)[0] : recp(j)[1]; };
    const double hs = sign1(xi_at(acc) - xi_at(0));
    /* chunk start: smallest j in [1,acc] (1-based) with hs*(Xint(j+1) - x) >= 0 */
    int loc;
    {
        const double x = xs[0];
        int lo = 1, hi = acc;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const double xm = xi_at(mid);
            if (!before(hs, xm, x) && xm == xm && x == x) hi = mid; else lo = mid + 1;
        }
        loc = lo;
    }
    double B[PS + 1][NI];
    double X0 = xi_at(loc - 1), X1, h;
    { const double* r = recp(loc); RecBuf<RS>::prefetch(rb, r + 2, r + 1); }
    RecBuf<RS>::take(rb, &B[0][0], X1);                               /* step loc: coefficients and Xint(loc+1) */
    h = X1 - X0;
    if (loc < acc) { const double* r = recp(loc + 1); RecBuf<RS>::prefetch(rb, r + 2, r + 1); }
    for (long g = g0; g < g1; ++g) {
        const double x = xs[g - g0];
        if (loc < acc && before(hs, X1, x)) {                         /* forward cursor (:85-103) */
            loc += 1;
            X0 = X1;
            RecBuf<RS>::take(rb, &B[0][0], X1);                       /* the record that was on its way */
            if (loc < acc && before(hs, X1, x)) {                     /* several steps between two grid points: walk Xint */
                do { loc += 1; X0 = X1; X1 = xi_at(loc); } while (loc < acc && before(hs, X1, x));
                { const double* r = recp(loc); RecBuf<RS>::prefetch(rb, r + 2, r + 1); }
                RecBuf<RS>::take(rb, &B[0][0], X1);
            }
            h = X1 - X0;
            if (loc < acc) { const double* r = recp(loc + 1); RecBuf<RS>::prefetch(rb, r + 2, r + 1); }
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
static cudaError_t launch_soa_one(const InterpArgs& a, dim3 grid, cudaStream_t s)
{
    constexpr int smem = FLINT_INTERP_CHUNK * 8 + RecBuf<(PS + 1) * NI>::BYTES;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(interp_kernel_soa<FMA, PS, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    interp_kernel_soa<FMA, PS, NI><<<grid, 128, smem, s>>>(a);
    return cudaGetLastError();
}
/* ======================================================================================================
 * interp_steps_kernel: a CTA = up to 128 consecutive accepted steps of ONE trajectory (blockIdx.x = trajectory, blockIdx.y =
 * block of steps; capacity levels are multiples of 128 steps, so a block never straddles two levels).  Their coefficient
 * records are contiguous in HBM: one thread arms an mbarrier with the byte count and issues ONE bulk copy (cp.async.bulk,
 * the TMA engine's 1-D form) into shared memory; the CTA then evaluates step-major with broadcast reads (interp_eval.cuh).
 * Three CTAs are resident per SM (3 x 58 KB), so one block's copy travels while the others compute.
 * ====================================================================================================== */
#ifndef FLINT_STEPS_TILE
#define FLINT_STEPS_TILE 64
#endif
#ifndef FLINT_STEPS_THREADS
#define FLINT_STEPS_THREADS 128
#endif
constexpr int kStepsBlock = FLINT_STEPS_TILE;      /* steps per CTA (divides the 128-step granularity of the capacity levels) */
constexpr int kStepsThreads = FLINT_STEPS_THREADS;
constexpr int kStepsResident = 65536 / (80 * kStepsThreads) < 16 ? 65536 / (80 * kStepsThreads) : 16;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <bool FMA, int PS, int NI, bool SOA>
__global__ void __launch_bounds__(kStepsThreads, kStepsResident) interp_steps_kernel(InterpArgs a)
{
    const long t = blockIdx.x;
    if (a.status[t] != FLINT_SUCCESS) return;                          /* validated before (launch_interp) */
    const long acc = a.dcount[t];
    const long s0 = (long)blockIdx.y * kStepsBlock;
    if (s0 >= acc) return;
    const int ns = (int)(acc - s0 < kStepsBlock ? acc - s0 : kStepsBlock);
    const int rstride = a.dv.rstride;
    extern __shared__ __align__(128) double ssm[];
    double* sR = ssm;                                                  /* [128][rstride]: the records as they lie in HBM */
    double* sX = ssm + kStepsBlock * rstride;                          /* [129] (+1 pad) */
    long* sG = reinterpret_cast<long*>(sX + kStepsBlock + 2);          /* [129] */
    uint64_t* bar = reinterpret_cast<uint64_t*>(sG + kStepsBlock + 2);
    const double* __restrict__ src = a.dv.rec(t, s0);
    const unsigned bytes = (unsigned)ns * (unsigned)rstride * 8u;      /* a multiple of 16: rstride is even */
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sR)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    }
    /* while the records travel: the step boundaries straight from HBM (8 bytes per step) and, from them, every step's range of
     * grid points -- the bisection's dependent loads overlap the bulk copy instead of following it */
    for (int i = threadIdx.x; i <= ns; i += blockDim.x) sX[i] = i < ns ? src[(long)i * rstride] : src[(long)(ns - 1) * rstride + 1];
    __syncthreads();                                                   /* also: the barrier is initialised before anyone polls it */
    const double hs = sign1(sX[1] - sX[0]);
    const double* __restrict__ xg = a.xarr + (a.per_traj ? t * a.G : 0L);
    block_step_ranges(sX, sG, ns, s0 == 0, s0 + ns == acc, xg, a.G, hs);
    {
        unsigned ok = 0;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
        } while (!ok);
    }
    double* yo = SOA ? a.yarr + t : a.yarr + t * a.G * NI;            /* [nI][G][N]: this trajectory's column */
    eval_block_steps<FMA, PS, NI, SOA>(sR + 2, rstride, sX, sG, ns, xg, a.G, yo, a.N);
}

template <bool FMA, int PS, int NI>
static cudaError_t launch_steps_one(const InterpArgs& a, cudaStream_t s)
{
    const long cap = a.dv.capacity();
    const long gy = (cap + kStepsBlock - 1) / kStepsBlock;
    if (gy < 1 || gy > 65535) return cudaErrorInvalidConfiguration;
    const int smem = (kStepsBlock * a.dv.rstride + 2 * (kStepsBlock + 2) + 2) * 8;
    const dim3 grid((unsigned)a.N, (unsigned)gy);
    if (a.layout == 0) {
        if (smem > 48 * 1024) {
            const cudaError_t e = cudaFuncSetAttribute(interp_steps_kernel<FMA, PS, NI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
        }
        interp_steps_kernel<FMA, PS, NI, false><<<grid, kStepsThreads, smem, s>>>(a);
    } else {
        if (smem > 48 * 1024) {
            const cudaError_t e = cudaFuncSetAttribute(interp_steps_kernel<FMA, PS, NI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
        }
        interp_steps_kernel<FMA, PS, NI, true><<<grid, kStepsThreads, smem, s>>>(a);
    }
    return cudaGetLastError();
}
template <bool FMA, int PS>
static cudaError_t launch_steps_ni(const InterpArgs& a, cudaStream_t s)
{
    switch (a.nI) {
    case 1: return launch_steps_one<FMA, PS, 1>(a, s);
    case 2: return launch_steps_one<FMA, PS, 2>(a, s);
    case 3: return launch_steps_one<FMA, PS, 3>(a, s);
    case 4: return launch_steps_one<FMA, PS, 4>(a, s);
    case 5: return launch_steps_one<FMA, PS, 5>(a, s);
    case 6: return launch_steps_one<FMA, PS, 6>(a, s);
    default: return cudaErrorInvalidValue;
    }
}
template <bool FMA>
static cudaError_t launch_steps_ps(const InterpArgs& a, cudaStream_t s)
{
    switch (a.pstar) {                         /* DOP54 4, Verner65E 5, DOP853 7, Verner98R 8 */
    case 4: return launch_steps_ni<FMA, 4>(a, s);
    case 5: return launch_steps_ni<FMA, 5>(a, s);
    case 7: return launch_steps_ni<FMA, 7>(a, s);
    case 8: return launch_steps_ni<FMA, 8>(a, s);
    default: return cudaErrorInvalidValue;
    }
}

template <bool FMA, int PS>
static cudaError_t launch_soa_ni(const InterpArgs& a, dim3 grid, cudaStream_t s)
{
    switch (a.nI) {
    case 1: return launch_soa_one<FMA, PS, 1>(a, grid, s);
    case 2: return launch_soa_one<FMA, PS, 2>(a, grid, s);
    case 3: return launch_soa_one<FMA, PS, 3>(a, grid, s);
    case 4: return launch_soa_one<FMA, PS, 4>(a, grid, s);
    case 5: return launch_soa_one<FMA, PS, 5>(a, grid, s);
    case 6: return launch_soa_one<FMA, PS, 6>(a, grid, s);
    default: return cudaErrorInvalidValue;
    }
}
template <bool FMA>
static cudaError_t launch_soa_ps(const InterpArgs& a, dim3 grid, cudaStream_t s)
{
    switch (a.pstar) {
    case 4: return launch_soa_ni<FMA, 4>(a, grid, s);
    case 5: return launch_soa_ni<FMA, 5>(a, grid, s);
    case 7: return launch_soa_ni<FMA, 7>(a, grid, s);
    case 8: return launch_soa_ni<FMA, 8>(a, grid, s);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_interp(const InterpArgs& a0, int up, int down, int sm_count, cudaStream_t s)
{
    InterpArgs a = a0;
    const cudaError_t ev = launch_interp_validate(a, up, down, s);
    if (ev != cudaSuccess) return ev;
    /* structure-of-arrays output, enough trajectories to fill the device with threads, one shared grid */
    if (a.layout != 0 && a.N >= 4096 && !a.per_traj) {
        const long want_t = (long)sm_count * 16 * 32;
        long nc = (want_t + a.N - 1) / a.N;
        const long min_nc = (a.G + FLINT_INTERP_CHUNK - 1) / FLINT_INTERP_CHUNK;
        if (nc < min_nc) nc = min_nc;
        if (nc > a.G) nc = a.G;
        if (nc <= 65535) {
            a.chunk = (a.G + nc - 1) / nc;
            a.nchunks = (int)((a.G + a.chunk - 1) / a.chunk);
            dim3 grid1((unsigned)((a.N + 127) / 128), (unsigned)a.nchunks);
            return a.fma ? launch_soa_ps<true>(a, grid1, s) : launch_soa_ps<false>(a, grid1, s);
        }
    }
    return a.fma ? launch_steps_ps<true>(a, s) : launch_steps_ps<false>(a, s);
}

}  // namespace flint
