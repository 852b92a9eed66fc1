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
 * so that a warp working on ONE trajectory reads whole records with coalesced 16-byte loads (the first version kept
 * trajectories innermost and every coefficient read used 8 of the 32 bytes it moved: 110 GB read for 63 GB written).
 *
 * Every grid point gets the reference's step: the smallest j with Xint(j+1) >= x in the direction of integration
 * (the forward cursor of :85-103; a chunk starts with the binary search that returns the same j), then
 * Y = sum_k B(:,k) * theta**k exactly as InterpY (:124-140; powers in __powidf2 order, arith.cuh).
 */
#include "arith.cuh"
#include "kargs.h"

#ifndef FLINT_INTERP_LDS128
#define FLINT_INTERP_LDS128 0
#endif

namespace flint {

/* hs * (X - x) < 0 of the reference's cursor test (:88) for hs = +-1: the sign of an IEEE difference of finite numbers is
 * exact (gradual underflow), so this is X < x for a forward and X > x for a backward integration -- one compare
 * instead of a subtraction, a product and a compare; NaN compares false either way */
#ifndef FLINT_INTERP_OLDCMP
#define FLINT_INTERP_OLDCMP 0
#endif
__device__ __forceinline__ bool before(double hs, double X, double x)
{
    if (FLINT_INTERP_OLDCMP) return hs * (X - x) < 0.0;
    return hs > 0.0 ? (X < x) : (X > x);
}

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
    const double* __restrict__ xi = a.dxint + t * (a.cap + 1);
    const double* __restrict__ rec0 = a.dbip + t * a.cap * (long)a.rstride;
    const double hs = sign1(xi[acc] - xi[0]);
    /* chunk start: smallest j in [1,acc] (1-based) with hs*(Xint(j+1) - x) >= 0 */
    int loc;
    {
        const double x = xs[0];
        int lo = 1, hi = acc;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (!before(hs, xi[mid], x) && xi[mid] == xi[mid] && x == x) hi = mid; else lo = mid + 1;
        }
        loc = lo;
    }
    double B[PS + 1][NI];
    double X0 = xi[loc - 1], X1, h;
    RecBuf<RS>::prefetch(rb, rec0 + (long)(loc - 1) * a.rstride, xi + loc);
    RecBuf<RS>::take(rb, &B[0][0], X1);                               /* step loc: coefficients and Xint(loc+1) */
    h = X1 - X0;
    if (loc < acc) RecBuf<RS>::prefetch(rb, rec0 + (long)loc * a.rstride, xi + loc + 1);
    for (long g = g0; g < g1; ++g) {
        const double x = xs[g - g0];
        if (loc < acc && before(hs, X1, x)) {                         /* forward cursor (:85-103) */
            loc += 1;
            X0 = X1;
            RecBuf<RS>::take(rb, &B[0][0], X1);                       /* the record that was on its way */
            if (loc < acc && before(hs, X1, x)) {                     /* several steps between two grid points: walk Xint */
                do { loc += 1; X0 = X1; X1 = xi[loc]; } while (loc < acc && before(hs, X1, x));
                RecBuf<RS>::prefetch(rb, rec0 + (long)(loc - 1) * a.rstride, xi + loc);
                RecBuf<RS>::take(rb, &B[0][0], X1);
            }
            h = X1 - X0;
            if (loc < acc) RecBuf<RS>::prefetch(rb, rec0 + (long)loc * a.rstride, xi + loc + 1);
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
/* ---- the evaluation kernel: one WARP per trajectory, lanes on 32 consecutive grid points ----
 * A thread-per-trajectory walk (the previous version) diverges: each lane changes step every ~G/acc points, so
 * the warp ran the ~100-instruction "load the next step's coefficients" path at almost every point with one
 * or two lanes active (ncu: 24.5 of 32 lanes, +40 % instructions).  With lanes on consecutive points of ONE
 * trajectory the points of a batch span only 1 + 32*acc/G steps; their records are copied once per warp into
 * a small shared-memory ring with coalesced 16-byte loads and read back with LDS, where lanes on the same step
 * broadcast.  A block is 8 warps = 8 trajectories on the same grid batch, so the [nI][G][N] layout can be written
 * as 64-byte rows (8 trajectories x 8 bytes: whole sectors) through a shared-memory transpose. */
/* one step record, RS doubles, from HBM into a shared-memory ring slot: coalesced 16-byte cp.async by the warp's lanes */
template <int RS>
__device__ __forceinline__ void fetch_record(const double* __restrict__ r, double* d, int lane)
{
    if (RS % 2 == 0) {
        if (lane < RS / 2) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(d + 2 * lane);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(r + 2 * lane));
        }
    } else {
        for (int q = lane; q < RS; q += 32) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(d + q);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(r + q));
        }
    }
}
/* InterpY (src/ERKInterp.f90:124-140) for one point of the step [X0, X1] whose coefficients B[k][ii] sit in shared memory */
template <bool FMA, int PS, int NI>
__device__ __forceinline__ void eval_point(double X0, double X1, double x, const double* __restrict__ B, double (&y)[NI])
{
    const double h = X1 - X0;
    const double theta = (x - X0) / h;
    double tk[PS + 1];
    theta_powers<PS>(theta, tk);
    if (FLINT_INTERP_LDS128 && NI % 2 == 0) {          /* two coefficients per 16-byte shared-memory read */
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

constexpr int kTW = 8;          /* trajectories (warps) per block */
constexpr int kRing = 8;        /* step records resident (or on their way) per warp */
constexpr int kAhead = 3;       /* records requested beyond the ones a batch needs: they arrive while it is evaluated */
constexpr int kWin = 128;       /* Xint entries staged per warp: win[i] = xi[win_lo + i]; refilled when < kWinLow lie ahead */
constexpr int kWinLow = 48;

template <bool FMA, int PS, int NI>
__global__ void __launch_bounds__(kTW * 32) interp_kernel(InterpArgs a)
{
    constexpr int RS = (PS + 1) * NI;
    constexpr int RSP = RS + (RS & 1);                  /* ring stride: keeps 16-byte alignment for odd records */
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ __align__(16) double s_ring[kTW][kRing][RSP];
    __shared__ double s_win[kTW][kWin + 1];
    __shared__ double s_tile[NI][32][kTW];
    __shared__ int s_active[kTW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long t0 = (long)blockIdx.x * kTW, t = t0 + warp;
    const long g0 = (long)blockIdx.y * a.chunk;
    long g1 = g0 + a.chunk;
    if (g1 > a.G) g1 = a.G;
    const bool active = t < a.N && a.status[t] == FLINT_SUCCESS;       /* warp-uniform */
    if (lane == 0) s_active[warp] = active ? 1 : 0;
    int acc = 1;
    const double* __restrict__ xi = a.dxint;
    const double* __restrict__ rec0 = a.dbip;
    double hs = 1.0;
    int loc_base = 1;
    if (active) {
        acc = a.dcount[t];
        xi = a.dxint + t * (a.cap + 1);
        rec0 = a.dbip + t * a.cap * (long)a.rstride;
        hs = sign1(xi[acc] - xi[0]);
        const double x = a.xarr[g0];                                   /* chunk start: smallest j in [1,acc] with hs*(Xint(j+1) - x) >= 0 */
        int lo = 1, hi = acc;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (!before(hs, xi[mid], x) && xi[mid] == xi[mid] && x == x) hi = mid; else lo = mid + 1;
        }
        loc_base = lo;
    }
    double* win = s_win[warp];
    int win_lo = -1;                                                   /* win[i] = xi[win_lo + i] */
    int ring_lo = 1, ring_hi = 1;                                      /* steps [ring_lo, ring_hi) are resident, step j in slot j % kRing */
    double xnext = 0.0;
    if (active) xnext = a.xarr[g0 + lane < g1 ? g0 + lane : g1 - 1];
    for (long gb = g0; gb < g1; gb += 32) {
        const long g = gb + lane;
        const bool valid = g < g1;
        double y[NI];
#pragma unroll
        for (int ii = 0; ii < NI; ++ii) y[ii] = 0.0;
        if (active) {
            const double x = xnext;
            { const long gn = g + 32; xnext = a.xarr[gn < g1 ? gn : g1 - 1]; }     /* next batch's abscissae, one iteration early */
            bool pending = valid;
            while (__any_sync(FULL, pending)) {
                if (win_lo < 0 || loc_base - 1 - win_lo > kWin - kWinLow) {       /* stage Xint(loc_base ..) */
                    win_lo = loc_base - 1;
                    __syncwarp();
                    for (int i = lane; i <= kWin; i += 32) { const int k = win_lo + i; win[i] = xi[k <= acc ? k : acc]; }
                    __syncwarp();
                }
                int j = loc_base;                                      /* forward cursor (:85-103), per lane */
                bool more = false;
                if (pending) {
                    while (j < acc && j - win_lo < kWin && before(hs, win[j - win_lo], x)) j += 1;
                    more = j < acc && j - win_lo >= kWin && before(hs, win[kWin], x);     /* the window ended before the step was found */
                }
                const bool ok = pending && !more && (j - loc_base) < kRing - kAhead;
                const unsigned okm = __ballot_sync(FULL, ok);
                if (okm == 0u) {                                       /* many small steps between two points: move on */
                    const int first = __ffs(__ballot_sync(FULL, pending)) - 1;
                    loc_base = __shfl_sync(FULL, j, first);            /* > loc_base: the window's end, or a step far ahead */
                    win_lo = -1;
                    continue;
                }
                const int last = 31 - __clz(okm);
                const int jmax = __shfl_sync(FULL, j, last);
                const int jmin = __shfl_sync(FULL, j, __ffs(okm) - 1);
                /* records of steps jmin..jmax (+ kAhead more) into the ring, asynchronously */
                if (ring_hi <= jmin || ring_lo > jmin) { ring_lo = jmin; ring_hi = jmin; }
                const bool had_all = jmax < ring_hi;                   /* everything this pass reads was requested by an earlier pass */
                int want_hi = jmax + 1 + kAhead;
                if (want_hi > acc + 1) want_hi = acc + 1;
                __syncwarp();                                          /* nobody still reads the slots about to be overwritten */
                for (int sidx = ring_hi; sidx < want_hi; ++sidx) fetch_record<RS>(rec0 + (long)(sidx - 1) * a.rstride, s_ring[warp][sidx % kRing], lane);
                asm volatile("cp.async.commit_group;");
                if (want_hi > ring_hi) ring_hi = want_hi;
                if (ring_hi - ring_lo > kRing) ring_lo = ring_hi - kRing;
                if (had_all) asm volatile("cp.async.wait_group 1;" ::: "memory");   /* only this pass's look-ahead may still fly */
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                if (ok) {
                    eval_point<FMA, PS, NI>(win[j - 1 - win_lo], win[j - win_lo], x, s_ring[warp][j % kRing], y);
                    pending = false;
                }
                loc_base = jmax;
            }
            if (a.layout == 0 && valid) {
                double* __restrict__ o = a.yarr + (t * a.G + g) * NI;
                if (NI % 2 == 0) {
#pragma unroll
                    for (int q = 0; q < NI / 2; ++q) reinterpret_cast<double2*>(o)[q] = make_double2(y[2 * q], y[2 * q + 1]);
                } else {
#pragma unroll
                    for (int ii = 0; ii < NI; ++ii) o[ii] = y[ii];
                }
            }
        }
        if (a.layout != 0) {                                           /* [nI][G][N]: transpose 8 trajectories x 32 points through shared memory */
#pragma unroll
            for (int ii = 0; ii < NI; ++ii) s_tile[ii][lane][warp] = y[ii];
            __syncthreads();
            const int p = threadIdx.x >> 3, w = threadIdx.x & 7;       /* 8 consecutive threads = 8 consecutive trajectories = 64 bytes */
            if (gb + p < g1 && t0 + w < a.N && s_active[w]) {
#pragma unroll
                for (int ii = 0; ii < NI; ++ii) a.yarr[((long)ii * a.G + gb + p) * a.N + t0 + w] = s_tile[ii][p][w];
            }
            __syncthreads();
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <bool FMA, int PS>
static cudaError_t launch_ni(const InterpArgs& a, dim3 grid, cudaStream_t s)
{
    switch (a.nI) {
    case 1: interp_kernel<FMA, PS, 1><<<grid, kTW * 32, 0, s>>>(a); break;
    case 2: interp_kernel<FMA, PS, 2><<<grid, kTW * 32, 0, s>>>(a); break;
    case 3: interp_kernel<FMA, PS, 3><<<grid, kTW * 32, 0, s>>>(a); break;
    case 4: interp_kernel<FMA, PS, 4><<<grid, kTW * 32, 0, s>>>(a); break;
    case 5: interp_kernel<FMA, PS, 5><<<grid, kTW * 32, 0, s>>>(a); break;
    case 6: interp_kernel<FMA, PS, 6><<<grid, kTW * 32, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
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
    const int vb = 256;
    interp_validate_kernel<<<(unsigned)((a.N + vb - 1) / vb), vb, 0, s>>>(a, up, down);
    if (a.layout != 0 && a.N >= 4096) {       /* structure-of-arrays output and enough trajectories to fill the device with threads */
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
    /* one warp per (trajectory, chunk of the grid): at least ~32 warps per SM in flight when N alone does not give them */
    const long want = (long)sm_count * 32;
    long nch = (want + a.N - 1) / a.N;
    const long max_nch = (a.G + 31) / 32;
    if (nch > max_nch) nch = max_nch;
    if (nch > 65535) nch = 65535;
    if (nch < 1) nch = 1;
    a.chunk = ((a.G + nch - 1) / nch + 31) / 32 * 32;
    a.nchunks = (int)((a.G + a.chunk - 1) / a.chunk);
    dim3 grid((unsigned)((a.N + kTW - 1) / kTW), (unsigned)a.nchunks);
    return a.fma ? launch_ps<true>(a, grid, s) : launch_ps<false>(a, grid, s);
}

}  // namespace flint
