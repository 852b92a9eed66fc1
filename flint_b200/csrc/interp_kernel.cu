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
    /* FLINT_B200_INTERP_KERNEL=steps: the one-tile-per-CTA kernel (kept for comparison; profiles/r02_notes.md) */
    static const bool use_steps = [] { const char* e = getenv("FLINT_B200_INTERP_KERNEL"); return !(e && e[0] == 't'); }();
    if (use_steps) return a.fma ? launch_steps_ps<true>(a, s) : launch_steps_ps<false>(a, s);
    return a.fma ? launch_tiles_ps<true>(a, sm_count, s) : launch_tiles_ps<false>(a, sm_count, s);
}

}  // namespace flint
