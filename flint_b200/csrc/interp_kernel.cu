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
 *   interp_items_kernel   [N][G][nI] with an even nI (each trajectory a Fortran Yarr(nI,G)): persistent, warp-specialised; a
 *                         thread keeps one step's coefficients in registers for <= 3 grid points, results leave as bulk stores.
 *   interp_steps_kernel   [N][G][nI] with an odd nI; also [nI][G][N] for small ensembles and grown stores.  A CTA = 128 consecutive steps of one trajectory: their records arrive as ONE bulk copy
 *                         (cp.async.bulk, completion on an mbarrier) and are evaluated step-major (interp_eval.cuh).
 *   interp_kernel_soa     [nI][G][N] with >= 4096 trajectories: one THREAD per trajectory, lanes on consecutive trajectories
 *                         -> coalesced stores, coefficients in registers.
 * The fused path (Integrate's step logs -> coefficients in registers -> Yarr, no records at all) is erk_dense_kernel
 * (erk_kernel.cuh); it shares the step-major evaluation.
 */
#include <cstdint>
#include <cstdlib>
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
    const int rstride = a.dv.rstride;
    extern __shared__ __align__(128) double ssm[];
    double* sR = ssm;                                                  /* [128][rstride]: the records as they lie in HBM */
    double* sX = ssm + kStepsBlock * rstride;                          /* [129] (+1 pad) */
    long* sG = reinterpret_cast<long*>(sX + kStepsBlock + 2);          /* [129] */
    uint64_t* bar = reinterpret_cast<uint64_t*>(sG + kStepsBlock + 2);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    /* gridDim.y covers level 0 of the store (launch_steps_one); a trajectory that grew deeper levels has its further blocks dealt
     * round-robin over its CTAs */
    unsigned phase = 0;
    for (long s0 = (long)blockIdx.y * kStepsBlock; s0 < acc; s0 += (long)gridDim.y * kStepsBlock, phase ^= 1u) {
    const int ns = (int)(acc - s0 < kStepsBlock ? acc - s0 : kStepsBlock);
    const double* __restrict__ src = a.dv.rec(t, s0);
    const unsigned bytes = (unsigned)ns * (unsigned)rstride * 8u;      /* a multiple of 16: rstride is even */
    __syncthreads();                                                   /* the barrier is initialised; the previous block's buffers are free */
    if (threadIdx.x == 0) {
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
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        } while (!ok);
    }
    double* yo = SOA ? a.yarr + t : a.yarr + t * a.G * NI;            /* [nI][G][N]: this trajectory's column */
    eval_block_steps<FMA, PS, NI, SOA>(sR + 2, rstride, sX, sG, ns, xg, a.G, yo, a.N);
    }
}

/* ======================================================================================================
 * interp_tiles_kernel: the same evaluation as a PERSISTENT, warp-specialised pipeline.  interp_steps_kernel above is bound by
 * the latency chain of each short-lived CTA (bulk copy -> step boundaries -> bisection -> evaluate: ncu long_scoreboard 2.6 and
 * barrier 0.9 of ~10 stall cycles per issue, a third of the resident warps idle in a prologue at any time).  Here a CTA lives for
 * the whole launch and walks tiles (trajectory, 64 consecutive steps) dealt round-robin; its warp 0 is the PRODUCER: for the tile
 * kTileStages-1 ahead it arms an mbarrier and issues the bulk copy of the tile's records (cp.async.bulk, TMA engine), and for the
 * tile whose copy has landed it reads the step boundaries from shared memory, counts each step's grid points (bisection with a
 * uniform-grid guess) and publishes the tile on a second mbarrier.  The CONSUMER warps only ever wait on that barrier, evaluate
 * their equal share of the tile's points (interp_eval.cuh) and hand the buffer back on a third.  Copies, range searches and
 * evaluation of different tiles overlap; no consumer ever waits for HBM.
 * ====================================================================================================== */
#ifndef FLINT_TILE_STEPS
#define FLINT_TILE_STEPS 64
#endif
#ifndef FLINT_TILE_CONSUMERS
#define FLINT_TILE_CONSUMERS 8
#endif
#ifndef FLINT_TILE_STAGES
#define FLINT_TILE_STAGES 2
#endif
#ifndef FLINT_TILE_CTAS
#define FLINT_TILE_CTAS 3
#endif
constexpr int kTileSteps = FLINT_TILE_STEPS, kTileConsumers = FLINT_TILE_CONSUMERS, kTileStages = FLINT_TILE_STAGES;
constexpr int kTileThreads = 32 * (kTileConsumers + 1);

__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    unsigned ok = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity)       /* has the phase of that parity completed?  Does not block */
{
    unsigned ok = 0;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
/* doubles per pipeline stage: records, step boundaries, point ranges, tile header (16-byte multiples) */
__host__ __device__ constexpr int tile_stage_doubles(int rstride) { return kTileSteps * rstride + 2 * (kTileSteps + 2) + 4; }

template <bool FMA, int PS, int NI, bool SOA>
__global__ void __launch_bounds__(kTileThreads, FLINT_TILE_CTAS) interp_tiles_kernel(InterpArgs a, long gy)
{
    const int rstride = a.dv.rstride;
    const int stage_d = tile_stage_doubles(rstride);
    extern __shared__ __align__(128) double tsm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + (long)kTileStages * stage_d);   /* tx[S], ready[S], empty[S] */
    uint64_t* bar_tx = bars; uint64_t* bar_ready = bars + kTileStages; uint64_t* bar_empty = bars + 2 * kTileStages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTileStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_tx + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_ready + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar_empty + s)), "r"(kTileConsumers));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long n_items = a.N * gy;
    auto stage_rec = [&](int s) -> double* { return tsm + (long)s * stage_d; };
    auto stage_x = [&](int s) -> double* { return stage_rec(s) + kTileSteps * rstride; };
    auto stage_g = [&](int s) -> long* { return reinterpret_cast<long*>(stage_x(s) + kTileSteps + 2); };
    auto stage_hdr = [&](int s) -> long* { return stage_g(s) + kTileSteps + 2; };          /* [0] trajectory, [1] ns (0: the end), [2] first step */

    if (warp == 0) {
        /* ------------------------------ producer ------------------------------ */
        long item = blockIdx.x;                       /* next candidate tile */
        long k_issue = 0, k_done = 0;                 /* tiles whose copy was issued / that were published */
        long q_t[kTileStages], q_s0[kTileStages]; int q_ns[kTileStages]; long q_acc[kTileStages];
        bool more = true;
        while (more || k_done < k_issue) {
            /* (1) publish the oldest tile in flight -- BEFORE the producer may block on a buffer below, so that the consumers
             * always find their next tile ready: boundaries from the records in shared memory, ranges of grid points */
            if (k_done < k_issue && (k_issue - k_done >= kTileStages - 1 || !more)) {
                const int s = (int)(k_done % kTileStages);
                mbar_wait(bar_tx + s, (unsigned)((k_done / kTileStages) & 1));
                const long t = q_t[s], s0 = q_s0[s], acc = q_acc[s];
                const int ns = q_ns[s];
                const double* sR = stage_rec(s);
                double* sX = stage_x(s);
                long* sG = stage_g(s);
                const double hs = sign1(sR[1] - sR[0]);            /* records are (X0, X1, ...): every step points in the direction of integration */
                const double* __restrict__ xg = a.xarr + (a.per_traj ? t * a.G : 0L);
                for (int i = lane; i <= ns; i += 32) {
                    const double X = i < ns ? sR[(long)i * rstride] : sR[(long)(ns - 1) * rstride + 1];
                    sX[i] = X;
                    long g;
                    if (i == 0) g = s0 == 0 ? 0 : grid_count_upto(xg, a.G, hs, X);
                    else if (i == ns && s0 + ns == acc) g = a.G;
                    else g = grid_count_upto(xg, a.G, hs, X);
                    sG[i] = g;
                }
                if (lane == 0) { long* h = stage_hdr(s); h[0] = t; h[1] = ns; h[2] = s0; }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ready + s);          /* release: the consumers acquire with their wait */
                ++k_done;
            }
            /* (2) the next tile's copy, as soon as its buffer is free (the tile kTileStages before it consumed) */
            if (more && k_issue - k_done < kTileStages) {
                long t = -1, s0 = 0, acc = 0;
                for (; item < n_items; item += gridDim.x) {            /* skip tiles beyond a trajectory's last step / failed grids */
                    t = item / gy; s0 = (item - t * gy) * kTileSteps;
                    acc = a.dcount[t];
                    if (s0 < acc && a.status[t] == FLINT_SUCCESS) break;
                    t = -1;
                }
                if (t < 0) { more = false; continue; }
                item += gridDim.x;
                const int s = (int)(k_issue % kTileStages);
                const int ns = (int)(acc - s0 < kTileSteps ? acc - s0 : kTileSteps);
                if (k_issue >= kTileStages) mbar_wait(bar_empty + s, (unsigned)((k_issue / kTileStages - 1) & 1));
                if (lane == 0) {
                    const unsigned bytes = (unsigned)ns * (unsigned)rstride * 8u;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar_tx + s)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(stage_rec(s))), "l"(a.dv.rec(t, s0)), "r"(bytes), "r"(smem_u32(bar_tx + s)) : "memory");
                }
                q_t[s] = t; q_s0[s] = s0; q_ns[s] = ns; q_acc[s] = acc;
                ++k_issue;
            }
        }
        /* the end: one more "tile" with ns = 0 */
        {
            const int s = (int)(k_done % kTileStages);
            if (k_done >= kTileStages) mbar_wait(bar_empty + s, (unsigned)((k_done / kTileStages - 1) & 1));
            if (lane == 0) { stage_hdr(s)[1] = 0; mbar_arrive(bar_ready + s); }
        }
    } else {
        /* ------------------------------ consumers ------------------------------ */
        const int cw = warp - 1;
        for (long k = 0;; ++k) {
            const int s = (int)(k % kTileStages);
            mbar_wait(bar_ready + s, (unsigned)((k / kTileStages) & 1));
            const long* h = stage_hdr(s);
            const int ns = (int)h[1];
            if (ns == 0) break;
            mbar_wait(bar_tx + s, (unsigned)((k / kTileStages) & 1));     /* complete long ago (the producer read the records): this thread's own view of the copy */
            const long t = h[0];
            const double* __restrict__ xg = a.xarr + (a.per_traj ? t * a.G : 0L);
            double* yo = SOA ? a.yarr + t : a.yarr + t * a.G * NI;
            eval_steps_share<FMA, PS, NI, SOA>(stage_rec(s) + 2, rstride, stage_x(s), stage_g(s), ns, xg, a.G, yo, a.N, cw, kTileConsumers, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + s);
        }
    }
}

/* ======================================================================================================
 * interp_items_kernel ([N][G][nI], nI even): the coefficients live in REGISTERS and the results leave as bulk stores.
 *
 * The kernels above put a warp's lanes on 32 consecutive grid points, so every lane re-reads its step's 54 coefficients from
 * shared memory for every point (27 16-byte reads, two wavefronts each when the batch straddles two steps) and the warp's 48-byte
 * results leave as three strided 16-byte stores (12 cache lines each): ~92 L1 wavefronts per 32 points beside ~57 cycles of fp64
 * issue and ~100 cycles of HBM time per SM -- three co-bound pipes -- and each short-lived CTA starts with a chain of dependent
 * loads (status, step count, boundaries, range search) that a third of the resident warps sit in at any time.
 *
 * Here the work item is (step, chunk of <= P consecutive grid points of that step) and a THREAD owns an item: it reads the
 * step's coefficients from shared memory ONCE, keeps them in registers for its <= P points (the point-independent first term
 * of the sum, 0 + B(:,0) * theta**0, is formed once per item), and drops each result into its warp's staging buffer in shared
 * memory at the point's place in Yarr.  The 32 items of a warp are consecutive in point order, so the buffer is one
 * contiguous piece of Yarr: one lane sends it with a bulk store (cp.async.bulk shared -> global, TMA engine) while the warp
 * already works on its next 32 items.  P is odd, so lanes on full chunks of one step hit distinct banks with their 16-byte
 * stores (P * 12 words = 4 mod 32 for P = 3), and the store's record stride is an odd number of 16-byte units (capi.cu:
 * dense_rstride; 29 for Verner98R with 6 states) so that lanes on different steps read their coefficients from different
 * banks although a tile of records arrives as ONE bulk copy.
 *
 * The kernel is persistent (one CTA per SM, trajectories dealt round-robin, a trajectory's tiles of 64 consecutive steps in
 * sequence) and warp-specialised.  Warp 0 is the HELPER: up to kItemsStages tiles ahead it issues the bulk copy of a tile's
 * records (completion on an mbarrier); for a tile that has landed it reads the step boundaries from shared memory and starts
 * the search for the first grid point of every step -- 64 boundaries on 32 lanes (the tile's first one is carried from the tile
 * before it), one round of loads each for a near-uniform grid (grid_probe_issue); the searches of two tiles are in flight, and the older one is finished
 * (grid_probe_result, warp scan for the first item of every step) and published on a second mbarrier.  The helper never blocks
 * on a barrier.  The eleven CONSUMER warps deal the 32-item rounds of the tile stream among themselves round-robin and wait
 * for nothing but the "tile planned" barrier.
 *
 * Measured (2^15 orbits x 10^4 points x 6 states, Verner98R, profiles/r02_notes.md): 5.78 -> 4.64 ms strict, 5.68 -> 4.15 ms
 * FMA against interp_steps_kernel.  Measured and dropped: a separate planning kernel (0.73 ms: 8 useful bytes per 64 fetched),
 * one bulk copy per record into padded slots (TMA issue-bound), the next round's search and x loads one round ahead (spills,
 * 6.3 ms), theta from a reciprocal of h kept per item (5.2 ms: the range check and fall-back per point cost more than the five
 * DFMA they save), 32-step tiles (helper-bound), P = 5 / 7 and 10 consumers.
 * ====================================================================================================== */
#ifndef FLINT_ITEMS_P
#define FLINT_ITEMS_P 3
#endif
#ifndef FLINT_ITEMS_CONSUMERS
#define FLINT_ITEMS_CONSUMERS 11
#endif
#ifndef FLINT_ITEMS_STAGES
#define FLINT_ITEMS_STAGES 5
#endif
constexpr int kItemsP = FLINT_ITEMS_P, kItemsConsumers = FLINT_ITEMS_CONSUMERS, kItemsStages = FLINT_ITEMS_STAGES;
constexpr int kItemsThreads = 32 * (kItemsConsumers + 1);
#ifndef FLINT_ITEMS_STEPS
#define FLINT_ITEMS_STEPS 64
#endif
constexpr int kItemsSteps = FLINT_ITEMS_STEPS;      /* steps per tile: 32 or 64 (divides the 128-step granularity of the capacity levels) */
constexpr int kPlanRow = kItemsSteps + 4;           /* int2 per tile: [i] = (first grid point, first item) of step i, [ns] = (end, items);
                                                       [65] = (trajectory, ns), [66] = (first step, steps of the trajectory) */
/* record stride in shared memory: >= rstride and = 2 mod 4 doubles, i.e. an odd number of 16-byte units (the store's own stride
 * is that already -- capi.cu: dense_rstride -- and a tile is ONE bulk copy; any other store is copied record by record) */
__host__ __device__ constexpr int items_rsp(int rstride) { return rstride + ((2 - rstride % 4) + 4) % 4; }
__host__ __device__ constexpr int items_stage_doubles(int rstride) { return kItemsSteps * items_rsp(rstride) + kPlanRow; }
__host__ __device__ constexpr int items_smem_doubles(int rstride, int ni)
{
    return kItemsStages * items_stage_doubles(rstride) + 3 * kItemsStages /* barriers */ + (kItemsStages & 1) + kItemsConsumers * 32 * kItemsP * ni;
}

template <bool FMA, int PS, int NI>
__global__ void __launch_bounds__(kItemsThreads, 1) interp_items_kernel(InterpArgs a)
{
    static_assert(NI % 2 == 0, "16-byte results");
    constexpr int P = kItemsP, S = kItemsStages;
    const int rstride = a.dv.rstride, rsp = items_rsp(rstride);
    const int stage_d = items_stage_doubles(rstride);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    extern __shared__ __align__(128) double ism[];
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(ism + (long)S * stage_d);      /* records landed (transaction bytes) */
    uint64_t* bar_ready = bar_full + S;                                             /* tile planned (helper) */
    uint64_t* bar_empty = bar_ready + S;                                            /* tile consumed (every consumer warp) */
    double* sYall = reinterpret_cast<double*>(bar_empty + S + (S & 1));
    auto stage_rec = [&](int s) -> double* { return ism + (long)s * stage_d; };
    auto stage_row = [&](int s) -> int2* { return reinterpret_cast<int2*>(stage_rec(s) + kItemsSteps * rsp); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_full + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_ready + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar_empty + s)), "r"(kItemsConsumers));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        /* ------------------------------ helper: bulk copies ahead, ranges and items of what has landed ------------------------------ */
        constexpr int NB = kItemsSteps / 32;                            /* boundaries per lane: lane + 1 + 32 u, u < NB (1 .. kItemsSteps) */
        long k_issue = 0, k_probe = 0, k_done = 0;                      /* tiles whose copy was issued / whose searches are in flight / published */
        /* the trajectories of this CTA, 32 at a time: step counts (0: none / failed grid) one batch ahead of their use */
        long tb = blockIdx.x;                                           /* first trajectory of the current batch */
        int q = 32;                                                     /* next lane of the batch; 32: load a batch */
        long my_acc = 0;
        long cur_t = -1, cur_acc = 0, cur_s0 = 0;                       /* trajectory being dealt, next step to issue */
        bool more = true;
        double xa = 0.0, xb = 0.0, inv_span = 0.0; long x_t = -2;       /* grid ends of the trajectory whose grid is loaded */
        /* the searches of up to TWO tiles are in flight (prA: the older), so a tile's loads have a whole turn of the loop to arrive.
         * A tile's first boundary is the last one of the tile before it (same trajectory, published just before) or the start of
         * the trajectory: it is carried, not searched -- kItemsSteps searches per tile on 32 lanes, NB rounds */
        GridProbe prA[NB], prB[NB];
        int n_fly = 0;
        int g_carry = 0;                                                /* first grid point of the next tile's first step */
        /* The helper never blocks on a barrier: a tile that has not landed, or a buffer that is not free yet, is looked at again in
         * the next turn of the loop, so that publishing what is ready never waits for either */
        while (more || k_done < k_issue) {
            bool did = false, probed = false;
            /* (1) start the searches of the next tile that has landed */
            if (n_fly < 2 && k_probe < k_issue && mbar_test(bar_full + (int)(k_probe % S), (unsigned)((k_probe / S) & 1))) {
                did = true; probed = true;
                const int s = (int)(k_probe % S);
                const int2* row = stage_row(s);
                const long t = row[kItemsSteps + 1].x;
                const int ns = row[kItemsSteps + 1].y;
                const double* sR = stage_rec(s);
                const double* __restrict__ xg = a.xarr + (a.per_traj ? t * a.G : 0L);
                const long xt = a.per_traj ? t : -1;
                if (xt != x_t) { xa = xg[0]; xb = xg[a.G - 1]; inv_span = 1.0 / (xb - xa); x_t = xt; }
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    const int i = lane + 1 + 32 * u;                       /* boundary i = end of step i - 1 */
                    const double X = i < ns ? sR[(long)i * rsp] : (i == ns ? sR[(long)(ns - 1) * rsp + 1] : xa);
                    if (n_fly == 0) grid_probe_issue(prA[u], xg, a.G, X, xa, inv_span);
                    else grid_probe_issue(prB[u], xg, a.G, X, xa, inv_span);
                }
                ++n_fly;
                ++k_probe;
            }
            /* (2) publish the oldest tile in flight, once a second one is behind it (or nothing could be started this turn) */
            if (n_fly == 2 || (n_fly == 1 && !probed)) {
                did = true;
                const int s = (int)(k_done % S);
                int2* row = stage_row(s);
                const long t = row[kItemsSteps + 1].x;
                const int ns = row[kItemsSteps + 1].y;
                const long s0 = row[kItemsSteps + 2].x, acc = row[kItemsSteps + 2].y;
                const double* sR = stage_rec(s);
                const double* __restrict__ xg = a.xarr + (a.per_traj ? t * a.G : 0L);
                const double hs = sign1(sR[1] - sR[0]);                  /* records are (X0, X1, ...): every step points in the direction of integration */
                double ta = xa, tb2 = xb;                                 /* a per-trajectory grid: the ends of THIS tile's grid (the newer tile may have another) */
                if (a.per_traj && x_t != t) { ta = xg[0]; tb2 = xg[a.G - 1]; }
                int ge[NB];                                               /* first grid point beyond step lane + 32 u (= boundary lane + 1 + 32 u) */
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    const int i = lane + 1 + 32 * u;
                    bool ok;
                    const long r = grid_probe_result(prA[u], a.G, hs, ta, tb2, ok);
                    if (i > ns) ge[u] = 0;
                    else if (i == ns && s0 + ns == acc) ge[u] = (int)a.G;
                    else ge[u] = (int)(ok ? r : grid_count_upto(xg, a.G, hs, prA[u].X));
                }
                const int g0 = s0 == 0 ? 0 : g_carry;                     /* boundary 0 */
                int base = 0;                                             /* items of the steps before this group of 32 */
                int prev_last = g0;                                       /* boundary 32 u: the end of the group before */
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    int gs = __shfl_up_sync(0xffffffffu, ge[u], 1);       /* start of step lane + 32 u */
                    if (lane == 0) gs = prev_last;
                    const int i = lane + 32 * u;                          /* this lane's step */
                    const int c = i < ns ? (ge[u] - gs + P - 1) / P : 0;  /* its items */
                    int inc = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
                    if (i < ns) row[i] = make_int2(gs, base + inc - c);
                    if (i == ns - 1) row[ns] = make_int2(ge[u], base + inc);      /* the end of the tile: last boundary, number of items */
                    base += __shfl_sync(0xffffffffu, inc, 31);
                    prev_last = __shfl_sync(0xffffffffu, ge[u], 31);
                }
                g_carry = prev_last;                                      /* boundary kItemsSteps (used only when the trajectory goes on: ns == kItemsSteps) */
                FLINT_CHK(ns >= 1 && ns <= kItemsSteps && t >= 0 && t < a.N && s0 >= 0 && s0 + ns <= acc && base >= 0 && g0 >= 0 && g0 <= a.G);
#pragma unroll
                for (int u = 0; u < NB; ++u) FLINT_CHK(ge[u] >= 0 && ge[u] <= a.G);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ready + s);                 /* release: the consumers acquire with their wait */
                ++k_done;
                --n_fly;
#pragma unroll
                for (int u = 0; u < NB; ++u) prA[u] = prB[u];
            }
            /* (3) the next tile's copy, as soon as its buffer is free (the tile kItemsStages before it consumed) */
            if (more && k_issue - k_done < S && (k_issue < S || mbar_test(bar_empty + (int)(k_issue % S), (unsigned)((k_issue / S - 1) & 1)))) {
                did = true;
                while (cur_t < 0 || cur_s0 >= cur_acc) {                   /* next trajectory with steps left */
                    if (q == 32) {
                        if (tb >= a.N) { more = false; break; }
                        const long tq = tb + (long)lane * gridDim.x;
                        my_acc = (tq < a.N && a.status[tq] == FLINT_SUCCESS) ? a.dcount[tq] : 0;
                        q = 0;
                    }
                    cur_t = tb + (long)q * gridDim.x;
                    cur_acc = __shfl_sync(0xffffffffu, my_acc, q);
                    cur_s0 = 0;
                    if (++q == 32) tb += 32L * gridDim.x;
                    if (cur_t >= a.N) { more = false; cur_t = -1; break; }
                }
                if (!more) continue;
                const long t = cur_t, s0 = cur_s0, acc = cur_acc;
                const int ns = (int)(acc - s0 < kItemsSteps ? acc - s0 : kItemsSteps);
                cur_s0 += kItemsSteps;
                const int s = (int)(k_issue % S);
                const double* __restrict__ src = a.dv.rec(t, s0);
                double* dst = stage_rec(s);
                if (lane == 0) {
                    int2* row = stage_row(s);
                    row[kItemsSteps + 1] = make_int2((int)t, ns);
                    row[kItemsSteps + 2] = make_int2((int)s0, (int)acc);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                 ::"r"(smem_u32(bar_full + s)), "r"((unsigned)ns * (unsigned)rstride * 8u) : "memory");
                }
                __syncwarp();
                if (rsp == rstride) {                                      /* ONE bulk copy */
                    if (lane == 0)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(smem_u32(dst)), "l"(src), "r"((unsigned)ns * (unsigned)rstride * 8u), "r"(smem_u32(bar_full + s)) : "memory");
                } else {
                    for (int i = lane; i < ns; i += 32)                    /* one bulk copy per record, into its padded slot */
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(smem_u32(dst + (long)i * rsp)), "l"(src + (long)i * rstride), "r"((unsigned)rstride * 8u), "r"(smem_u32(bar_full + s)) : "memory");
                }
                ++k_issue;
            }
            if (!did) __nanosleep(100);
        }
        /* the end: one more "tile" without steps */
        {
            const int s = (int)(k_done % S);
            if (k_done >= S) mbar_wait(bar_empty + s, (unsigned)((k_done / S - 1) & 1));
            if (lane == 0) { stage_row(s)[kItemsSteps + 1] = make_int2(-1, 0); mbar_arrive(bar_ready + s); }
        }
        return;
    }
    /* ------------------------------ consumers ------------------------------ */
    const int cw = warp - 1;
    double* sY = sYall + (long)cw * 32 * P * NI;                       /* this warp's piece of Yarr */
    bool pending = false;                                              /* a bulk store of sY may still be reading it */
    int roff = 0;                                                      /* rounds dealt so far, mod the number of consumers */
    for (long k = 0;; ++k) {
        const int s = (int)(k % S);
        mbar_wait(bar_ready + s, (unsigned)((k / S) & 1));
        const int2* row = stage_row(s);
        const int ns = row[kItemsSteps + 1].y;
        if (ns == 0) break;
        mbar_wait(bar_full + s, (unsigned)((k / S) & 1));              /* complete long ago (the helper read the records): this thread's own view of the copy */
        const double* sR = stage_rec(s);
        const long t = row[kItemsSteps + 1].x;
        const int T = row[ns].y;
        const int R = (T + 31) >> 5;
        const double* __restrict__ xg = a.xarr + (a.per_traj ? t * a.G : 0L);
        double* __restrict__ yo = a.yarr + t * a.G * NI;
        int r = cw - roff; if (r < 0) r += kItemsConsumers;
        for (; r < R; r += kItemsConsumers) {
            const int kk = r * 32 + lane < T ? r * 32 + lane : T - 1;
            const bool valid = r * 32 + lane < T;
            int i;
            {                                                          /* the step of item kk: largest i with first_item[i] <= kk */
                int lo = 0, hi = ns - 1;
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (row[mid].y <= kk) lo = mid; else hi = mid - 1; }
                i = lo;
            }
            const int2 e0 = row[i];
            const int ge = row[i + 1].x;
            int ga = e0.x + (kk - e0.y) * P;
            int len = ge - ga < P ? ge - ga : P;
            if (!valid) { ga = ge; len = 0; }
            const int wbase = __shfl_sync(0xffffffffu, ga, 0);
            const int wend = __shfl_sync(0xffffffffu, ga + len, 31);   /* items are consecutive in point order; invalid lanes sit at the end */
            FLINT_CHK(t >= 0 && t < a.N && i >= 0 && i < ns && len >= 0 && ga >= wbase && ga + len <= wend && wend <= a.G &&
                      wend - wbase <= 32 * P && e0.y <= kk && kk < row[i + 1].y + (valid ? 0 : 1));
            double xs[P];
#pragma unroll
            for (int j = 0; j < P; ++j) xs[j] = j < len ? xg[ga + j] : 0.0;
            double2 b[(PS + 1) * (NI / 2)];
            const double* rec = sR + (long)i * rsp;
            {
                const double2* __restrict__ B2 = reinterpret_cast<const double2*>(rec + 2);
#pragma unroll
                for (int q = 0; q < (PS + 1) * (NI / 2); ++q) b[q] = B2[q];
            }
            const double X0 = rec[0], h = rec[1] - X0;
            /* the sum's first term, 0 + B(:,0) * theta**0, does not depend on the point: once per item (same operations, same bits) */
#pragma unroll
            for (int q = 0; q < NI / 2; ++q) { b[q].x = madd<FMA>(b[q].x, 1.0, 0.0); b[q].y = madd<FMA>(b[q].y, 1.0, 0.0); }
            if (pending) {                                             /* the previous piece has left shared memory */
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            double2* __restrict__ o = reinterpret_cast<double2*>(sY + (long)(ga - wbase) * NI);
            /* FMA arithmetic (not bound by the fp64 pipe): every lane evaluates P points, those beyond its chunk on x = 0 and
             * unused -- no branch separates the points and the division chain of the next one overlaps the sums of this one.
             * Strict arithmetic (bound by the fp64 pipe): a point is evaluated only by the lanes that have it */
#pragma unroll
            for (int j = 0; j < P; ++j) {
                if (FMA || j < len) {
                    const double theta = (xs[j] - X0) / h;             /* InterpY, :124-140 */
                    double tk[PS + 1], y[NI];
                    theta_powers<PS>(theta, tk);
#pragma unroll
                    for (int q = 0; q < NI / 2; ++q) { y[2 * q] = b[q].x; y[2 * q + 1] = b[q].y; }
#pragma unroll
                    for (int kq = 1; kq <= PS; ++kq)
#pragma unroll
                        for (int q = 0; q < NI / 2; ++q) {
                            y[2 * q] = madd<FMA>(b[kq * (NI / 2) + q].x, tk[kq], y[2 * q]);
                            y[2 * q + 1] = madd<FMA>(b[kq * (NI / 2) + q].y, tk[kq], y[2 * q + 1]);
                        }
                    if (j < len) {
#pragma unroll
                        for (int q = 0; q < NI / 2; ++q) o[j * (NI / 2) + q] = make_double2(y[2 * q], y[2 * q + 1]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the generic-proxy writes above, visible to the bulk store */
            __syncwarp();
            if (lane == 0 && wend > wbase) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(yo + (long)wbase * NI), "r"(smem_u32(sY)), "r"((unsigned)((wend - wbase) * NI * 8)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            pending = true;
        }
        roff = (roff + R) % kItemsConsumers;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + s);
    }
    if (pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   /* shared memory outlives its readers */
}

template <bool FMA, int PS, int NI>
static cudaError_t launch_items_one(const InterpArgs& a, int sm_count, cudaStream_t s, bool* served)
{
    if constexpr (NI % 2 != 0) return cudaSuccess;
    else {
        if (a.G < 1 || a.N > 0x7fffffffL || a.G > 0x7fffffffL / 8 || a.dv.capacity() > 0x7fffffffL) return cudaSuccess;
        const int smem = items_smem_doubles(a.dv.rstride, NI) * 8;
        if (smem > 227 * 1024) return cudaSuccess;                     /* records too large for the pipeline: the step-major kernel serves */
        cudaError_t e = cudaFuncSetAttribute(interp_items_kernel<FMA, PS, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        const long grid = a.N < sm_count ? a.N : sm_count;
        interp_items_kernel<FMA, PS, NI><<<(unsigned)grid, kItemsThreads, smem, s>>>(a);
        *served = true;
        return cudaGetLastError();
    }
}
template <bool FMA, int PS>
static cudaError_t launch_items_ni(const InterpArgs& a, int sm_count, cudaStream_t s, bool* served)
{
    switch (a.nI) {
    case 2: return launch_items_one<FMA, PS, 2>(a, sm_count, s, served);
    case 4: return launch_items_one<FMA, PS, 4>(a, sm_count, s, served);
    case 6: return launch_items_one<FMA, PS, 6>(a, sm_count, s, served);
    default: return cudaSuccess;
    }
}
template <bool FMA>
static cudaError_t launch_items_ps(const InterpArgs& a, int sm_count, cudaStream_t s, bool* served)
{
    switch (a.pstar) {
    case 4: return launch_items_ni<FMA, 4>(a, sm_count, s, served);
    case 5: return launch_items_ni<FMA, 5>(a, sm_count, s, served);
    case 7: return launch_items_ni<FMA, 7>(a, sm_count, s, served);
    case 8: return launch_items_ni<FMA, 8>(a, sm_count, s, served);
    default: return cudaSuccess;
    }
}

template <bool FMA, int PS, int NI>
static cudaError_t launch_tiles_one(const InterpArgs& a, int sm_count, cudaStream_t s)
{
    const long cap = a.dv.capacity();
    const long gy = (cap + kTileSteps - 1) / kTileSteps;
    if (gy < 1) return cudaErrorInvalidConfiguration;
    const int smem = (kTileStages * tile_stage_doubles(a.dv.rstride) + 3 * kTileStages) * 8;
    const void* k = a.layout == 0 ? (const void*)interp_tiles_kernel<FMA, PS, NI, false> : (const void*)interp_tiles_kernel<FMA, PS, NI, true>;
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    int bps = 0;
    { const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, kTileThreads, (size_t)smem); if (e != cudaSuccess) return e; }
    if (bps < 1) return cudaErrorInvalidConfiguration;
    long grid = (long)sm_count * bps;
    if (grid > a.N * gy) grid = a.N * gy;
    InterpArgs b = a;
    long gyv = gy;
    void* args[2] = {&b, &gyv};
    return cudaLaunchKernel(k, dim3((unsigned)grid), dim3(kTileThreads), args, (size_t)smem, s);
}
template <bool FMA, int PS>
static cudaError_t launch_tiles_ni(const InterpArgs& a, int sm_count, cudaStream_t s)
{
    switch (a.nI) {
    case 1: return launch_tiles_one<FMA, PS, 1>(a, sm_count, s);
    case 2: return launch_tiles_one<FMA, PS, 2>(a, sm_count, s);
    case 3: return launch_tiles_one<FMA, PS, 3>(a, sm_count, s);
    case 4: return launch_tiles_one<FMA, PS, 4>(a, sm_count, s);
    case 5: return launch_tiles_one<FMA, PS, 5>(a, sm_count, s);
    case 6: return launch_tiles_one<FMA, PS, 6>(a, sm_count, s);
    default: return cudaErrorInvalidValue;
    }
}
template <bool FMA>
static cudaError_t launch_tiles_ps(const InterpArgs& a, int sm_count, cudaStream_t s)
{
    switch (a.pstar) {
    case 4: return launch_tiles_ni<FMA, 4>(a, sm_count, s);
    case 5: return launch_tiles_ni<FMA, 5>(a, sm_count, s);
    case 7: return launch_tiles_ni<FMA, 7>(a, sm_count, s);
    case 8: return launch_tiles_ni<FMA, 8>(a, sm_count, s);
    default: return cudaErrorInvalidValue;
    }
}

template <bool FMA, int PS, int NI>
static cudaError_t launch_steps_one(const InterpArgs& a, cudaStream_t s)
{
    const long cap = a.dv.capacity();
    long gy = (cap + kStepsBlock - 1) / kStepsBlock;
    if (gy < 1) return cudaErrorInvalidConfiguration;
    { const long gy0 = a.dv.cum[1] / kStepsBlock; if (gy0 >= 1 && gy > gy0) gy = gy0; }   /* level 0; deeper blocks: the kernel's loop */
    if (gy > 65535) return cudaErrorInvalidConfiguration;
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

cudaError_t launch_interp(const InterpArgs& a0, int up, int down, int sm_count, cudaStream_t s, int* launches)
{
    InterpArgs a = a0;
    if (launches) *launches = 2;                                       /* validation + one evaluation kernel */
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
    /* FLINT_B200_INTERP_KERNEL = items | steps | tiles picks one kernel for comparison (profiles/r02_notes.md); by default the
     * thread-per-item kernel takes every [N][G][nI] call it can serve (nI even, 16-byte aligned Yarr), the step-major one the rest */
    const int which = [] { const char* e = getenv("FLINT_B200_INTERP_KERNEL"); return !e ? 0 : e[0] == 'i' ? 1 : e[0] == 's' ? 2 : e[0] == 't' ? 3 : 0; }();
    const bool items_ok = a.layout == 0 && a.nI % 2 == 0 && (reinterpret_cast<uintptr_t>(a.yarr) & 15) == 0;
    if (items_ok && which <= 1) {
        bool served = false;
        const cudaError_t e = a.fma ? launch_items_ps<true>(a, sm_count, s, &served) : launch_items_ps<false>(a, sm_count, s, &served);
        if (e != cudaSuccess) return e;
        if (served) return cudaSuccess;
    }
    const bool use_steps = which != 3;
    if (use_steps) return a.fma ? launch_steps_ps<true>(a, s) : launch_steps_ps<false>(a, s);
    return a.fma ? launch_tiles_ps<true>(a, sm_count, s) : launch_tiles_ps<false>(a, sm_count, s);
}

}  // namespace flint
