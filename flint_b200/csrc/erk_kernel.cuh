/* erk_kernel.cuh -- the ensemble integrator kernels (one trajectory per thread).
 *
 * Device restatement of ERK_class%Integrate (src/ERKIntegrate.f90:29-802) for N
 * independent trajectories.  B200-first structure, not the reference's loop nest.  The driver is
 * FOUR kernels around a per-trajectory state block in HBM (SoA over trajectories):
 *
 *    erk_init_kernel    one thread per trajectory: argument checks, F0, StepSz0Hairer, first event
 *                       values (:204-259)  ->  state block
 *    erk_step_kernel    the hot loop: persistent lanes load a trajectory from the state block and
 *                       step it until it finishes (results stored) or -- EVMODE 2 -- until a step
 *                       with an event to locate, which is PARKED: step data to the state block,
 *                       index appended to a list, the lane takes the next trajectory
 *    erk_event_kernel   one thread per parked trajectory: sub-step scan, Brent, actions, commit of
 *                       the step  ->  state block; the next erk_step_kernel pass resumes them
 *    erk_dense_kernel   InterpOn: one thread per (trajectory, accepted step): dense-output
 *                       coefficients from the step the hot loop logged
 *
 * Why: everything that one lane does alone costs 32 lanes.  Measured on the single-kernel version
 * (profiles/r01_notes.md): 7.5 % of the warp instructions ran with < 8 active lanes (event location
 * and per-trajectory initialisation inside the lane loop) and took 14 % of the time, and every such
 * excursion pushed ~100 KB of rare code through the 32 KB instruction cache of an SM whose 16 warps
 * share it (no_instruction stalls 3.1 vs 0.8 cycles per issue without events).  Batched by kind, the
 * rare work runs converged and the hot kernel holds only the step loop.  EVMODE 1 keeps the
 * in-loop event handling for configurations that locate events on (nearly) every step (unmasked
 * events with an EventStepSz), where parking would advance one step per pass.
 *
 *  * persistent lanes: a thread owns one trajectory at a time; when it finishes it
 *    pulls the next index from a global counter, so lanes whose trajectories need
 *    fewer steps do not idle until the slowest lane of the warp is done;
 *  * the adaptive loop is flattened (one step ATTEMPT per warp iteration for every
 *    lane that owns a trajectory), accept/reject is predicated data flow; one CTA-wide barrier per
 *    attempt (which is also the exit vote) keeps the CTA's warps in the same part of the loop;
 *  * state and stage vectors live in registers (generated straight-line stage code,
 *    erk_methods_gen.cuh; KStore below); HBM traffic is the state block and the SoA results,
 *    a few hundred bytes per trajectory and pass;
 *  * everything rare (sub-step event scanning, Brent root finding, event actions)
 *    is one __noinline__ function working on a context in local memory, so its
 *    register needs do not tax the hot loop.  It REMATERIALISES the stage vectors
 *    of the current step (recomputes the step) instead of keeping 13-21 stage
 *    vectors alive across the call;
 *  * dense coefficients are evaluated lazily: when the reference computes them only
 *    because a (possibly masked) event has an EventStepSz (quirk Q1,
 *    src/ERKIntegrate.f90:332-344) the kernel accounts the FCalls the reference
 *    would report but evaluates nothing unless an event actually needs them.
 *
 * Line citations `:NNN` below are into src/ERKIntegrate.f90 unless noted.
 */
#ifndef FLINT_B200_ERK_KERNEL_CUH
#define FLINT_B200_ERK_KERNEL_CUH

#ifdef FLINT_GANG_TRACE
#include <cstdio>
#endif
#ifndef __CUDACC_RTC__
#include <type_traits>
#endif
#include "erk_methods_gen.cuh"
#include "systems.cuh"
#include "interp_eval.cuh"

namespace flint {

constexpr double kEps = 2.220446049250313e-16;   /* FLINT_EPS = epsilon(1.0_WP), src/FLINTUtils.f90:49 */
constexpr int kRootMaxItr = 50;                  /* FLINT_ROOTFIND_MAXITR, src/FLINTUtils.f90:52 */
constexpr int kStiffTestSteps = 1000;            /* STIFFTEST_STEPS, src/ERK.f90:53 */
constexpr double kLastStepExpFac = 1.01;         /* LASTSTEP_EXPFAC, src/ERK.f90:43 */
constexpr int kMaxSS = FLINT_MAXNUMSTEPSZEVENTS; /* MAXNUMSTEPSZEVENTS, src/ERK.f90:48 */

/* system parameters (flint_sys_create) and, for user functors that declare load_params(const double*, int), the
 * `params` optional of Integrate (src/FLINT_base.f90:400-468; the reference hands it to F on every call) */
template <class SYS, class = void> struct has_load_params : meta::false_type {};
template <class SYS>
struct has_load_params<SYS, meta::void_t<decltype(meta::declval<SYS&>().load_params((const double*)nullptr, 0))>> : meta::true_type {};
template <class SYS> __device__ __forceinline__ void sys_load(SYS& sys, const KArgs& p)
{
    sys.load(p.sp);
    if constexpr (has_load_params<SYS>::value) sys.load_params(p.params, p.n_params);
}

/* ---- DetectEvent (:715-741) ---- */
__device__ __forceinline__ bool detect_event(bool masked_in, int dir, double v0, double v1)
{
    if (!masked_in || is_nan(v0) || is_nan(v1)) return false;
    if (dir == 0) return (v0 <= 0.0 && v1 >= 0.0) || (v0 >= 0.0 && v1 <= 0.0);
    if (dir == 1) return (v0 <= 0.0 && v1 >= 0.0);
    if (dir == -1) return (v0 >= 0.0 && v1 <= 0.0);
    return false;
}

/* ---- Root: Brent (src/FLINTUtils.f90:69-207) ---- */
template <class Func>
__device__ __forceinline__ void brent_root(double a, double b, double ya, double yb, double tol, Func f,
                                           double& x, double& fx, int& niter, int& error)
{
    double aa = a, bb = b, cc = 0.0, fa = ya, fb = yb, fc, xm, ss, dd = 0.0, ee = 0.0, pp, qq, rr, tolx;
    const double NEARZERO = 1.0e-20;
    error = 0; niter = 0;
    x = bb;                                                            /* defined stand-in for the reference's undefined x */
    if (fabs(fa) <= NEARZERO) { x = aa; fx = fa; return; }
    else if (fabs(fb) <= NEARZERO) { x = bb; fx = fb; return; }
    if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) { error = 1; return; }
    fc = fb;
    bool done = false;
    int itr = 0;
    while (!done && itr < kRootMaxItr) {
        if ((fc > 0.0 && fb > 0.0) || (fc < 0.0 && fb < 0.0)) { cc = aa; fc = fa; dd = bb - aa; ee = dd; }
        if (fabs(fc) < fabs(fb)) { aa = bb; bb = cc; cc = aa; fa = fb; fb = fc; fc = fa; }
        tolx = __dadd_rn(__dmul_rn(2.0 * kEps, fabs(bb)), 0.5 * tol);
        xm = 0.5 * (cc - bb);
        if (fabs(xm) <= tolx || fabs(fa) < NEARZERO) { x = bb; done = true; fx = f(x); }
        else {
            if (fabs(ee) >= tolx && fabs(fa) > fabs(fb)) {
                ss = fb / fa;
                if (fabs(aa - cc) < NEARZERO) { pp = (2.0 * xm) * ss; qq = 1.0 - ss; }
                else {
                    qq = fa / fc; rr = fb / fc;
                    pp = __dmul_rn(ss, __dsub_rn(__dmul_rn(__dmul_rn(2.0 * xm, qq), qq - rr), __dmul_rn(bb - aa, rr - 1.0)));
                    qq = __dmul_rn(__dmul_rn(qq - 1.0, rr - 1.0), ss - 1.0);
                }
                if (pp > NEARZERO) qq = -qq;
                pp = fabs(pp);
                const double lim = dmin(__dsub_rn(__dmul_rn(3.0 * xm, qq), fabs(__dmul_rn(tolx, qq))), fabs(__dmul_rn(ee, qq)));
                if ((2.0 * pp) < lim) { ee = dd; dd = pp / qq; }
                else { dd = xm; ee = dd; }
            } else { dd = xm; ee = dd; }
            aa = bb; fa = fb;
            if (fabs(dd) > tolx) bb = bb + dd;
            else { if (xm > 0.0) bb = bb + fabs(tolx); else bb = bb - fabs(tolx); }
            fb = f(bb);
            itr = itr + 1;
        }
    }
    if (itr >= kRootMaxItr) error = 2;
    niter = itr;
}

/* ---- launch configuration of erk_step_kernel per system.  A system may declare
 *   static constexpr int STEP_BLOCK, STEP_MIN_BLOCKS;   CTA size and resident CTAs per SM (__launch_bounds__)
 *   static constexpr int STEP_BARRIER;                  1: one __syncthreads per step attempt, 0: none
 * Defaults: 128 x 4 with the barrier.  What is being traded (profiles/r01_notes.md): the hot loop is about the size of the
 * 32 KB instruction cache of an SM.  Warps that run it in step share instruction fetches; warps in different phases keep the
 * fp64 pipe busier.  A CTA's warps start together and stay roughly in step on their own; the barrier makes that exact. ---- */
template <class SYS, class = void> struct has_step_block : meta::false_type {};
template <class SYS> struct has_step_block<SYS, meta::void_t<decltype(SYS::STEP_BLOCK)>> : meta::true_type {};
template <class SYS, class = void> struct has_step_barrier : meta::false_type {};
template <class SYS> struct has_step_barrier<SYS, meta::void_t<decltype(SYS::STEP_BARRIER)>> : meta::true_type {};
template <class SYS> __host__ __device__ constexpr int step_block()
{
    if constexpr (has_step_block<SYS>::value) return SYS::STEP_BLOCK; else return FLINT_BLOCK_THREADS;
}
template <class SYS> __host__ __device__ constexpr int step_min_blocks()
{
    if constexpr (has_step_block<SYS>::value) return SYS::STEP_MIN_BLOCKS; else return FLINT_MIN_BLOCKS;
}
template <class SYS> __host__ __device__ constexpr int step_barrier_value()
{
    if constexpr (has_step_barrier<SYS>::value) return SYS::STEP_BARRIER; else return 1;
}

/* ---- stage-vector store.  Generated code reads k_j[c] as ks.get<slot>(c) and writes a whole vector with
 * ks.put<slot>(f).  Slots in SMEM_MASK live in shared memory ([slot][component][thread]: conflict-free, the
 * offset is an immediate), the others in registers.  Why: at a right-hand-side call the caller holds ~10 live
 * stage vectors, which leaves the callee too few registers to run its two square-root and reciprocal chains
 * in lock step -- ptxas then serialises ~68 dependent fp64 operations of 8 cycles each
 * (profiles/r01_notes.md).  Components that are structurally +0 are never stored. ---- */
template <class SYS> __host__ __device__ constexpr bool comp_stored(int c) { return !SYS::zc(c) && SYS::fsrc(c) != -1; }
template <class SYS> __host__ __device__ constexpr int comp_rank(int c)
{
    int r = 0;
    for (int i = 0; i < c; ++i) if (comp_stored<SYS>(i)) ++r;
    return r;
}
template <class SYS> __host__ __device__ constexpr int comps_stored() { return comp_rank<SYS>(SYS::N); }
__host__ __device__ constexpr int slot_rank(unsigned mask, int slot)
{
    int r = 0;
    for (int i = 0; i < slot; ++i) if ((mask >> i) & 1u) ++r;
    return r;
}
template <class METHOD, class SYS, unsigned SMEM_MASK>
struct KStore {
    static constexpr int N = SYS::N, NC = comps_stored<SYS>();
    double r[METHOD::NSLOT][N];
    volatile double* s;                          /* this thread's column of the dynamic shared memory; volatile: every get is a load
                                                    (otherwise the compiler forwards the stored registers and nothing is saved) */
    template <int SLOT> __device__ __forceinline__ double get(int c) const
    {
        if (!comp_stored<SYS>(c)) return 0.0;
        if constexpr ((SMEM_MASK >> SLOT) & 1u) return s[(slot_rank(SMEM_MASK, SLOT) * NC + comp_rank<SYS>(c)) * step_block<SYS>()];
        else return r[SLOT][c];
    }
    template <int SLOT> __device__ __forceinline__ void put(const double (&v)[N])
    {
#pragma unroll
        for (int c = 0; c < N; ++c) {
            if (!comp_stored<SYS>(c)) continue;
            if constexpr ((SMEM_MASK >> SLOT) & 1u) s[(slot_rank(SMEM_MASK, SLOT) * NC + comp_rank<SYS>(c)) * step_block<SYS>()] = v[c];
            else r[SLOT][c] = v[c];
        }
    }
    template <int SLOT> __device__ __forceinline__ void load(double (&v)[N]) const
    {
#pragma unroll
        for (int c = 0; c < N; ++c) v[c] = get<SLOT>(c);
    }
};
template <class METHOD, class SYS, unsigned SMEM_MASK>
__host__ __device__ constexpr int kstore_smem_bytes() { return slot_rank(SMEM_MASK, 32) * comps_stored<SYS>() * step_block<SYS>() * 8; }

/* ---- context of the rare event path (lives in local memory) ---- */
template <class METHOD, class SYS>
struct EvCtx {
    static constexpr int N = SYS::N, M = SYS::M_MAX, PS = METHOD::PSTAR;
    /* the accepted step whose events must be located */
    double X, h, h_stage;
    double Y[N], Y1[N], F0old[N], ENY[N];
    double EV0[M], EV1[M];
    int EvDir[M];
    unsigned evmask, det, ss;
    int action, st, fcalls, evcount;
    bool LAST, bip_counted, haveB;
    double B[PS + 1][N];
    /* what the deferred commit needs besides (and returns through X, h, Y1 = new Y, F0new = new F0) */
    double F0new[N], Err, hbyhopt, hbyhoptOld, hmax;
    int acc;
    bool lastRej;
};

/* Rematerialise the stage vectors of the current step (recompute it from X, Y, F0) and build the
 * dense coefficients with (h, Y1) -- which an event action may have changed: quirk Q6 rebuilds the
 * dense stages with the new h and the old k1..k_sint. */
template <class METHOD, class SYS, bool FMA>
__device__ __forceinline__ void remat_bip(const SYS& sys, const KArgs& p, EvCtx<METHOD, SYS>& c)
{
    if (c.haveB) return;
    constexpr int N = SYS::N;
    KStore<METHOD, SYS, 0u> ks;
    ks.s = nullptr;
    double Yn[N], Yp[N], e;
    METHOD::template step<SYS, FMA, true, false>(sys, c.X, c.Y, c.F0old, c.h_stage, c.h, ks, Yn, Yp, p.tol, false, e, p.coef);
    METHOD::template dense_coeffs<SYS, FMA>(ks, c.Y, c.h, c.Y1, c.B, p.coef);
    c.haveB = true;
}
template <class METHOD, class SYS, bool FMA>
__device__ __forceinline__ void ensure_bip(const SYS& sys, const KArgs& p, EvCtx<METHOD, SYS>& c)
{
    if (!c.bip_counted) { c.fcalls += METHOD::FC_DENSE; c.bip_counted = true; }   /* BipComputed / FCalls as the reference */
    remat_bip<METHOD, SYS, FMA>(sys, p, c);
}
/* Sub-step scanning, event location and action handling (:332-547). */
template <class METHOD, class SYS, bool FMA>
__device__ __noinline__ void event_rare_path(const SYS& sys, const KArgs& p, EvCtx<METHOD, SYS>& c, long idx, double hSign)
{
    constexpr int N = SYS::N, M = SYS::M_MAX, PS = METHOD::PSTAR;
    const int m = p.m;
    const double X = c.X;
    double EX0[M][kMaxSS], EX1[M][kMaxSS], EV0[M][kMaxSS], EV1[M][kMaxSS], EXm[M][kMaxSS];
    int nSS[M];
    double EventX0 = X, EventX1;
    int dummy_action = 0;
    for (int i = 0; i < M; ++i) {
        nSS[i] = 1;
        for (int j = 0; j < kMaxSS; ++j) {
            EX0[i][j] = X; EX1[i][j] = X + c.h; EV0[i][j] = 0.0; EV1[i][j] = 0.0; EXm[i][j] = 0.0;
        }
        EV0[i][0] = c.EV0[i]; EV1[i][0] = c.EV1[i];
    }

    /* ---- sub-step scanning (:337-412) ---- */
    if (c.ss) {
        for (int ev = 0; ev < m; ++ev) {
            if (!((c.evmask >> ev) & 1u) || !((c.ss >> ev) & 1u)) continue;
            ensure_bip<METHOD, SYS, FMA>(sys, p, c);
            /* only event `ev` is evaluated inside the step (EvalEvents(ev) = 1), so its two bracket values are scalars */
            double v0 = EV0[ev][0], v1 = EV1[ev][0];
            const double Eventh = hSign * p.ev_stepsz[ev];
            EventX1 = X + Eventh;
            const double Vh = v1;
            c.det &= ~(1u << ev);
            bool last = false;
            while (EventX1 <= (X + c.h)) {                               /* :372 (quirk Q7) */
                if (EventX1 != (X + c.h)) {
                    double Vt[M];
#pragma unroll
                    for (int i = 0; i < M; ++i) Vt[i] = 0.0;
                    interp_y<FMA, PS, N>(EventX1, X, c.h, c.B, c.ENY);
                    sys.events(p.evset, EventX1, c.ENY, 1u << ev, Vt, c.EvDir, 0, dummy_action);
#pragma unroll
                    for (int i = 0; i < M; ++i) if (i == ev) v1 = Vt[i];
                } else v1 = Vh;
                if (detect_event((c.evmask >> ev) & 1u, c.EvDir[ev], v0, v1)) {
                    c.det |= (1u << ev);
                    const int slot = nSS[ev] - 1;
                    EX0[ev][slot] = EventX0; EX1[ev][slot] = EventX1; EV0[ev][slot] = v0; EV1[ev][slot] = v1;
                    nSS[ev] = nSS[ev] + 1;
                    if (nSS[ev] > kMaxSS) break;
                }
                if (last) break;
                EventX0 = EventX1;
                EventX1 = EventX1 + Eventh;
                if (EventX1 > X + c.h) { EventX1 = X + c.h; last = true; }
                v0 = v1;
            }
            if (nSS[ev] > 1) nSS[ev] = nSS[ev] - 1;
        }
    }

    /* ---- locate (:416-460) ---- */
    if (c.det) {
        for (int ev = 0; ev < m; ++ev) {
            for (int se = 0; se < nSS[ev]; ++se) {
                if (!((c.det >> ev) & 1u)) continue;
                const int opt = p.ev_opt[ev];
                if (opt == FLINT_EVENTOPTION_STEPBEGIN) EXm[ev][se] = EX0[ev][se];
                else if (opt == FLINT_EVENTOPTION_STEPEND) EXm[ev][se] = EX1[ev][se];
                else {
                    ensure_bip<METHOD, SYS, FMA>(sys, p, c);
                    auto single_event = [&](double xx) -> double {       /* SingleEvent :746-763 */
                        double ff[N], val[M]; int edir[M]; int da = 0;
                        for (int i = 0; i < M; ++i) { val[i] = 0.0; edir[i] = 0; }
                        interp_y<FMA, PS, N>(xx, X, c.h, c.B, ff);
                        sys.events(p.evset, xx, ff, 1u << ev, val, edir, 0, da);
                        return val[ev];
                    };
                    int rootiter = 0, rooterror = 0;
                    double fxr = EV0[ev][se];
                    brent_root(EX0[ev][se], EX1[ev][se], EV0[ev][se], EV1[ev][se], p.etol[ev], single_event,
                               EXm[ev][se], fxr, rootiter, rooterror);
                    EV0[ev][se] = fxr;
                    if (rooterror != 0 && rooterror != 2) c.st = FLINT_ERROR_EVENTROOT;   /* :452-455 */
                }
            }
        }
        /* ---- handle earliest-first (:465-546) ---- */
        while (c.det) {
            int mi = -1;
            for (int i = 0; i < m; ++i)
                if (((c.det >> i) & 1u) && (mi < 0 || EXm[i][0] < EXm[mi][0])) mi = i;
            double EYm[N];
            const int opt = p.ev_opt[mi];
            if (opt == FLINT_EVENTOPTION_STEPBEGIN) { for (int i = 0; i < N; ++i) EYm[i] = c.Y[i]; }
            else if (opt == FLINT_EVENTOPTION_STEPEND) { for (int i = 0; i < N; ++i) EYm[i] = c.Y1[i]; }
            else interp_y<FMA, PS, N>(EXm[mi][0], X, c.h, c.B, EYm);
            for (int i = 0; i < N; ++i) c.ENY[i] = EYm[i];
            {
                double val[M];
                for (int i = 0; i < M; ++i) val[i] = EV0[i][0];
                sys.events(p.evset, EXm[mi][0], c.ENY, 1u << mi, val, c.EvDir, mi + 1, c.action);   /* :487-488 */
                for (int i = 0; i < M; ++i) EV0[i][0] = val[i];
            }
            if (p.ev_rec && c.evcount < p.ev_cap) {                      /* EventData :493 */
                const long e = c.evcount;
                FLINT_CHK(e >= 0 && e < p.ev_cap && idx >= 0 && idx < p.N);
                p.ev_rec[(0 * p.ev_cap + e) * p.N + idx] = EXm[mi][0];
                for (int i = 0; i < N; ++i) p.ev_rec[((long)(1 + i) * p.ev_cap + e) * p.N + idx] = EYm[i];
                p.ev_rec[((long)(N + 1) * p.ev_cap + e) * p.N + idx] = (double)(mi + 1);
            }
            c.evcount = c.evcount + 1;
            if (c.action & FLINT_EVENTACTION_MASK) c.evmask &= ~(1u << mi);   /* :499-501 */
            const int act = c.action & 0x7f;
            if (act == FLINT_EVENTACTION_TERMINATE) {                    /* :505-512 */
                c.h = EXm[mi][0] - X;
                for (int i = 0; i < N; ++i) c.Y1[i] = EYm[i];
                c.LAST = true; c.bip_counted = false; c.haveB = false; c.st = FLINT_EVENT_TERM;
                break;
            } else if (act == FLINT_EVENTACTION_RESTARTINT || act == FLINT_EVENTACTION_CHANGESOLY) {   /* :513-528 */
                c.h = EXm[mi][0] - X;
                for (int i = 0; i < N; ++i) c.Y1[i] = EYm[i];
                c.LAST = false; c.bip_counted = false; c.haveB = false;
                break;
            }
            if ((c.ss >> mi) & 1u) {                                     /* :534-544 */
                nSS[mi] = nSS[mi] - 1;
                if (nSS[mi] == 0) c.det &= ~(1u << mi);
                else for (int j = 0; j < nSS[mi]; ++j) EXm[mi][j] = EXm[mi][j + 1 < kMaxSS ? j + 1 : j];
            } else c.det &= ~(1u << mi);
        }
    }
    for (int i = 0; i < M; ++i) c.EV1[i] = EV1[i][0];                   /* what `EV0 = EV1` (:550) carries to the next step */
}

/* ---- dense-output storage (InterpOn): records of a growing store, kargs.h: DenseView.
 * The hot loop only LOGS an accepted step (2n+2 doubles); the extra dense stages and the coefficients are computed afterwards by
 * erk_dense_kernel, one thread per (trajectory, step), by recomputing the step from its log -- the same bits, because F is
 * pure and the arithmetic is the same straight-line code.  Measured before the split (dense stages, coefficients and the
 * 54-double store inside the lane loop): Verner98R 5.2x, DOP853 2.7x slower than without InterpOn. ---- */
template <class METHOD, class SYS>
__device__ __forceinline__ void store_coef_record(double* __restrict__ rec, int rstride, int nI, const int* istates, double X0, double X1,
                                                  const double (&B)[METHOD::PSTAR + 1][SYS::N])
{
    constexpr int N = SYS::N, PS = METHOD::PSTAR;
    reinterpret_cast<double2*>(rec)[0] = make_double2(X0, X1);        /* records are 16-byte aligned (rstride is even) */
    double* __restrict__ c = rec + 2;
    if (nI == N) {                     /* all states, in order (the default): the record is B itself */
        if (((PS + 1) * N) % 2 == 0) {
#pragma unroll
            for (int q = 0; q < (PS + 1) * N / 2; ++q)
                reinterpret_cast<double2*>(c)[q] = make_double2(B[(2 * q) / N][(2 * q) % N], B[(2 * q + 1) / N][(2 * q + 1) % N]);
        } else {
#pragma unroll
            for (int j = 0; j <= PS; ++j)
#pragma unroll
                for (int cc = 0; cc < N; ++cc) c[j * N + cc] = B[j][cc];
        }
    } else {                           /* a subset (InterpStates): B is only ever indexed with constants, so it stays in registers */
        for (int ii = 0; ii < nI; ++ii) {
            const int want = istates[ii] - 1;
#pragma unroll
            for (int j = 0; j <= PS; ++j) {
                double v = 0.0;
#pragma unroll
                for (int cc = 0; cc < N; ++cc) if (want == cc) v = B[j][cc];
                c[j * nI + ii] = v;
            }
        }
    }
    rec[rstride - 1] = kDenseCoef;
}
/* accepted step `acc` (1-based) of the hot loop: X, h, Y, F0.  A call that cannot grow the store drops what does not fit
 * (Interpolate then reports FLINT_ERROR_INTP_ARRAY for that trajectory); every other call suspends the trajectory first. */
template <class SYS>
__device__ __forceinline__ void log_dense_step(const KArgs& p, long idx, int acc, double X, double h, const double (&Y)[SYS::N], const double (&F0)[SYS::N])
{
    constexpr int N = SYS::N;
    if (acc > p.dv.capacity()) return;
    double* __restrict__ rec = p.dv.rec(idx, acc - 1);
    reinterpret_cast<double2*>(rec)[0] = make_double2(X, h);
#pragma unroll
    for (int q = 0; q < N / 2; ++q) reinterpret_cast<double2*>(rec)[1 + q] = make_double2(Y[2 * q], Y[2 * q + 1]);
    if (N % 2) rec[2 + N - 1] = Y[N - 1];
#pragma unroll
    for (int c = 0; c < N; ++c) rec[2 + N + c] = F0[c];
    rec[p.dv.rstride - 1] = kDenseLog;
}
/* a step committed by the event code: its end may have been moved to the event (h != h_stage) and Y1 replaced by the event
 * state (:505-528); the dense stages are then rebuilt with the new h and the old stages (quirk Q6, :507-510,558-562) */
template <class SYS>
__device__ __forceinline__ void log_dense_event_step(const KArgs& p, long idx, int acc, double X, double h, const double (&Y)[SYS::N],
                                                     const double (&F0)[SYS::N], double h_stage, const double (&Y1)[SYS::N])
{
    constexpr int N = SYS::N;
    if (acc > p.dv.capacity()) return;
    double* __restrict__ rec = p.dv.rec(idx, acc - 1);
    rec[0] = X; rec[1] = h;
#pragma unroll
    for (int c = 0; c < N; ++c) { rec[2 + c] = Y[c]; rec[2 + N + c] = F0[c]; rec[3 + 2 * N + c] = Y1[c]; }
    rec[2 + 2 * N] = h_stage;
    rec[p.dv.rstride - 1] = kDenseEvLog;
}

/* ---- the tail of an accepted step (:555-635): dense store, stale-action quirk, natural-step output,
 * advance, new step size, too-small test.  Shared by the hot path (state in registers) and by the
 * deferred event commit (state in the local-memory context). ---- */
/* dense_mode (InterpOn only): 1 = log the step here as an event-committed step (h_stage = the step size its stages were
 * computed with), 2 = the hot path has logged it already, 0 = InterpOn is off. */
template <class METHOD, class SYS, bool FMA, bool EVENTS>
__device__ __forceinline__ void commit_tail(const KArgs& p, long idx, int dense_mode, bool eventsOn, int evAction, double hSign, double hmax,
        double& X, double (&Y)[SYS::N], double (&F0)[SYS::N], double& h_io, double& hbyhoptOld, bool& lastRej, bool LAST, int& st,
        int& fcalls, int acc, double h, double (&Y1)[SYS::N], const double (&F0new)[SYS::N], double Err, double hbyhopt,
        bool bip_counted, const double (&ENY)[SYS::N], double h_stage, double& powKey, double& powVal)
{
    constexpr int N = SYS::N;
    const long NN = p.N;
    const bool const_step = p.const_step != 0;
    if (dense_mode != 0) {                                              /* :555-566 */
        if (!bip_counted) { fcalls += METHOD::FC_DENSE; bip_counted = true; }
        if (dense_mode == 1) log_dense_event_step<SYS>(p, idx, acc, X, h, Y, F0, h_stage, Y1);
    }
    if (EVENTS && eventsOn) {                                           /* :572-582 (quirks Q4, Q5) */
        const int act = evAction & 0x7f;
        if (act == FLINT_EVENTACTION_RESTARTINT) fcalls += 1;
        else if (act == FLINT_EVENTACTION_CHANGESOLY) {
#pragma unroll
            for (int c = 0; c < N; ++c) Y1[c] = ENY[c];
            fcalls += 1;
        }
    }
    if (p.steps_on && acc < p.steps_cap) {                              /* :585-588 */
        p.xint[(long)acc * NN + idx] = X + h;
#pragma unroll
        for (int c = 0; c < N; ++c) p.yint[((long)c * p.steps_cap + acc) * NN + idx] = Y1[c];
    }
    X = X + h;                                                          /* :591-592 */
#pragma unroll
    for (int c = 0; c < N; ++c) { Y[c] = Y1[c]; F0[c] = F0new[c]; }
    double hnew;
    if (!const_step) {                                                  /* :595-605, src/StepSz.f90:156-168 */
        const double SFl = p.szp[0], SFmin = p.szp[1], SFmax = p.szp[2], beta = p.szp[3];
        if (beta != 0.0) {
            /* hbyhoptOld is a running maximum of the error norm: it changes a few times per trajectory, so its power
             * is kept from step to step where the method's default controller uses it (DOP54: one pow per step saved) */
            constexpr bool CACHE = METHOD::ID == ERK_DOP54;
            double pw;
            if (CACHE && hbyhoptOld == powKey) pw = powVal;
            else {
                pw = det_pow_dev(hbyhoptOld, beta, p.powc);
                if (CACHE) { powKey = hbyhoptOld; powVal = pw; }
            }
            hbyhopt = hbyhopt / pw;
        }
        hnew = h * dmin(SFmax, dmax(SFmin, SFl / hbyhopt));
        hbyhoptOld = dmax(Err, hbyhoptOld);
        if (fabs(hnew) > hmax) hnew = hSign * hmax;
        if (lastRej) hnew = hSign * dmin(fabs(hnew), fabs(h));
        lastRej = false;
    } else hnew = h;
    h_io = hnew;                                                        /* :627 */
    if (!const_step && !LAST && fabs(hnew) * 0.01 <= fabs(X) * kEps) st = FLINT_ERROR_STEPSZ_TOOSMALL;   /* :632-635 */
}

/* ---- deferred commit of a step with events to locate.  Called at the top of the lane loop, where only
 * the loop-carried state is live: the step's data was stashed in the context when the sign change was
 * seen, so nothing of the hot step has to survive the call.  Runs the rare path, then the commit tail. ---- */
template <class METHOD, class SYS, bool FMA>
__device__ __noinline__ void deferred_event_commit(const SYS& sys, const KArgs& p, EvCtx<METHOD, SYS>& c, long idx, double hSign)
{
    constexpr int M = SYS::M_MAX;
    event_rare_path<METHOD, SYS, FMA>(sys, p, c, idx, hSign);
#pragma unroll
    for (int i = 0; i < M; ++i) c.EV0[i] = c.EV1[i];                  /* :550 */
    const bool dense = p.dv.nlev > 0;    /* the step is logged with (h, h_stage, Y1): erk_dense_kernel rebuilds its dense stages (quirk Q6) */
    double Xn = c.X, hio = c.h, pk = qnan(), pv = 0.0;
    commit_tail<METHOD, SYS, FMA, true>(p, idx, dense ? 1 : 0, true, c.action, hSign, c.hmax, Xn, c.Y, c.F0old, hio, c.hbyhoptOld, c.lastRej,
                                               c.LAST, c.st, c.fcalls, c.acc, c.h, c.Y1, c.F0new, c.Err, c.hbyhopt, c.bip_counted, c.ENY, c.h_stage, pk, pv);
    c.X = Xn; c.h = hio;       /* new X, next step size; new Y in c.Y, new F0 in c.F0old */
}

/* ===================================================================== trajectory state block */
/* Everything a trajectory carries from one step to the next, plus (parked trajectories) the data of
 * the accepted step whose events are to be located.  SoA: field f of trajectory i is at [f * cap + i]. */
template <int N, int M> struct SD {            /* double fields */
    static constexpr int X = 0, H = 1, HMAX = 2, HOLD = 3, Y = 4, F0 = 4 + N, EV0 = 4 + 2 * N,
                         HSTEP = 4 + 2 * N + M, Y1 = HSTEP + 1, F0N = Y1 + N, EV1 = F0N + N, ERR = EV1 + M, HBH = ERR + 1,
                         X0 = HBH + 1, COUNT = X0 + 1;
};
struct SI {                                    /* int fields */
    static constexpr int TOTAL = 0, ACC = 1, REJ = 2, FC = 3, ST = 4, FLAGS = 5, STIFF = 6, EVP = 7, EVC = 8, DET = 9, SS = 10, COUNT = 11;
};
constexpr int kFlagLast = 1, kFlagLastRej = 2, kFlagEventsOn = 4, kFlagDone = 8, kFlagBipCounted = 16;

/* the loop-carried state of one trajectory, in registers */
template <class SYS> struct Lane {
    static constexpr int N = SYS::N, M = SYS::M_MAX;
    double X, h, Xf, hSign, hmax, hbyhoptOld;
    double Y[N], F0[N], EV0[M];
    int total, acc, rej, fcalls, st;
    bool LAST, lastRej, eventsOn;
    int stiffMode, stiffThr, nonStiffThr, stiffOut;
    int EvDir[M];
    unsigned evmask;
    int evAction, evCount;
};

template <class SYS>
__device__ __forceinline__ void state_store(const KArgs& p, long idx, const Lane<SYS>& L, int extra_flags)
{
    FLINT_CHK(idx >= 0 && idx < p.scap);
    constexpr int N = SYS::N, M = SYS::M_MAX;
    using D = SD<N, M>;
    const long c = p.scap;
    double* d = p.sd + idx;
    int* q = p.si + idx;
    d[D::X * c] = L.X; d[D::H * c] = L.h; d[D::HMAX * c] = L.hmax; d[D::HOLD * c] = L.hbyhoptOld;
#pragma unroll
    for (int i = 0; i < N; ++i) { d[(D::Y + i) * c] = L.Y[i]; d[(D::F0 + i) * c] = L.F0[i]; }
#pragma unroll
    for (int i = 0; i < M; ++i) d[(D::EV0 + i) * c] = L.EV0[i];
    q[SI::TOTAL * c] = L.total; q[SI::ACC * c] = L.acc; q[SI::REJ * c] = L.rej; q[SI::FC * c] = L.fcalls; q[SI::ST * c] = L.st;
    q[SI::FLAGS * c] = (L.LAST ? kFlagLast : 0) | (L.lastRej ? kFlagLastRej : 0) | (L.eventsOn ? kFlagEventsOn : 0) | extra_flags;
    /* only `== 15` (stiffThr) and `== 6` (nonStiffThr) are ever tested (:786-795): counts past them saturate instead of wrapping */
    q[SI::STIFF * c] = (L.stiffMode & 3) | ((L.stiffOut + 1) << 2) | ((L.stiffThr > 15 ? 15 : L.stiffThr) << 4) | ((L.nonStiffThr > 15 ? 15 : L.nonStiffThr) << 8);
    int evp = (int)(L.evmask & 0xffu) | ((L.evAction & 0xff) << 8);
#pragma unroll
    for (int i = 0; i < M; ++i) evp |= ((L.EvDir[i] + 1) & 3) << (16 + 2 * i);
    q[SI::EVP * c] = evp; q[SI::EVC * c] = L.evCount;
}
/* returns the flags word */
template <class SYS>
__device__ __forceinline__ int state_load(const KArgs& p, long idx, Lane<SYS>& L)
{
    FLINT_CHK(idx >= 0 && idx < p.scap);
    constexpr int N = SYS::N, M = SYS::M_MAX;
    using D = SD<N, M>;
    const long c = p.scap;
    const double* d = p.sd + idx;
    const int* q = p.si + idx;
    const int flags = q[SI::FLAGS * c];
    L.X = d[D::X * c]; L.h = d[D::H * c]; L.hmax = d[D::HMAX * c]; L.hbyhoptOld = d[D::HOLD * c];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (SYS::zc(i)) { L.Y[i] = 0.0; L.F0[i] = 0.0; }                 /* identically +0 (plane variant) */
        else { L.Y[i] = d[(D::Y + i) * c]; L.F0[i] = d[(D::F0 + i) * c]; }
    }
#pragma unroll
    for (int i = 0; i < M; ++i) L.EV0[i] = d[(D::EV0 + i) * c];
    L.Xf = p.xf[idx];
    L.hSign = sign1(L.Xf - d[D::X0 * c]);                                /* :94 */
    L.total = q[SI::TOTAL * c]; L.acc = q[SI::ACC * c]; L.rej = q[SI::REJ * c]; L.fcalls = q[SI::FC * c]; L.st = q[SI::ST * c];
    L.LAST = (flags & kFlagLast) != 0; L.lastRej = (flags & kFlagLastRej) != 0; L.eventsOn = (flags & kFlagEventsOn) != 0;
    const int sf = q[SI::STIFF * c];
    L.stiffMode = sf & 3; L.stiffOut = ((sf >> 2) & 3) - 1; L.stiffThr = (sf >> 4) & 15; L.nonStiffThr = (sf >> 8) & 15;
    const int evp = q[SI::EVP * c];
    L.evmask = (unsigned)(evp & 0xff); L.evAction = (evp >> 8) & 0xff;
#pragma unroll
    for (int i = 0; i < M; ++i) L.EvDir[i] = ((evp >> (16 + 2 * i)) & 3) - 1;
    L.evCount = q[SI::EVC * c];
    return flags;
}

/* final status resolution and result store (:642-708) */
template <class SYS>
__device__ __forceinline__ void store_results(const KArgs& p, long idx, const Lane<SYS>& L)
{
    constexpr int N = SYS::N;
    const long NN = p.N;
    int fs = L.st;
    if (L.LAST && L.st == FLINT_SUCCESS) {}
    else if (L.eventsOn && L.LAST && L.st == FLINT_EVENT_TERM) {}
    else if (L.total >= p.max_steps && L.st == FLINT_SUCCESS) fs = FLINT_ERROR_MAXSTEPS;
    else if (L.st == FLINT_ERROR_STEPSZ_TOOSMALL) {}
    else if (L.eventsOn && L.st == FLINT_ERROR_EVENTROOT) {}
    else if (L.st == FLINT_STIFF_PROBLEM) {}
    else fs = FLINT_ERROR;
    p.xf[idx] = L.X;
#pragma unroll
    for (int i = 0; i < N; ++i) p.yf[(long)i * NN + idx] = L.Y[i];
    if (!p.const_step && p.stepsz) p.stepsz[idx] = L.h;
    p.status[idx] = fs;
    if (p.counters) {
        p.counters[0 * NN + idx] = L.total; p.counters[1 * NN + idx] = L.acc;
        p.counters[2 * NN + idx] = L.rej; p.counters[3 * NN + idx] = L.fcalls;
    }
    if (p.stifftest) p.stifftest[idx] = L.stiffOut;
    if (p.ev_count) p.ev_count[idx] = L.evCount;
    if (p.steps_count) p.steps_count[idx] = L.acc + 1;
    if (p.dense_count) p.dense_count[idx] = L.acc;       /* allocated(me%Xint): acc + 1 abscissae, acc coefficient sets */
    p.si[SI::FLAGS * p.scap + idx] = kFlagDone;
}
template <class SYS> __host__ __device__ constexpr bool is_plane_variant()
{
    for (int c = 0; c < SYS::N; ++c) if (SYS::zc(c)) return true;
    return false;
}

template <class SYS>
__device__ __forceinline__ bool keeps_going(const KArgs& p, const Lane<SYS>& L)
{
    return L.st == FLINT_SUCCESS && !L.LAST && L.total < p.max_steps;      /* :263 */
}

/* ===================================================================== init: one thread per trajectory (:204-259) */
template <class METHOD, class SYS, bool FMA>
__global__ void __launch_bounds__(128) erk_init_kernel(const __grid_constant__ KArgs p)
{
    constexpr int N = SYS::N, M = SYS::M_MAX;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long NN = p.N;
    if (idx >= NN) return;
    SYS sys;
    sys_load(sys, p);
    const int m = p.m;
    const bool const_step = p.const_step != 0;
    Lane<SYS> L;
    L.X = p.x0 ? p.x0[idx] : p.x0_scalar;
    p.sd[SD<N, M>::X0 * p.scap + idx] = L.X;                              /* for the direction of integration */
#pragma unroll
    for (int c = 0; c < N; ++c) { L.Y[c] = p.y0[(long)c * NN + idx]; L.F0[c] = 0.0; }
#pragma unroll
    for (int i = 0; i < M; ++i) { L.EV0[i] = 0.0; L.EvDir[i] = 0; }
    if (SYS::ZPLANE != 0u && p.zflag) {                                     /* every trajectory in the plane? (bitwise +0.0) */
        bool in_plane = true;
#pragma unroll
        for (int c = 0; c < N; ++c) if ((SYS::ZPLANE >> c) & 1u) in_plane = in_plane && (__double_as_longlong(L.Y[c]) == 0LL);
        if (!in_plane) *p.zflag = 0;
    }
    L.Xf = p.xf[idx];
    L.total = 0; L.acc = 0; L.rej = 0; L.fcalls = 0; L.st = FLINT_SUCCESS;
    L.LAST = false; L.lastRej = false; L.stiffThr = 0; L.nonStiffThr = 0;
    L.hbyhoptOld = p.szp[5];
    L.evAction = FLINT_EVENTACTION_CONTINUE; L.evCount = 0;
    L.hSign = sign1(L.Xf - L.X);                                            /* :94 */
    L.h = p.stepsz ? p.stepsz[idx] : 0.0;
    L.hmax = 0.0;
    bool early = false;   /* argument errors: the reference returns before touching Xf/Yf (:196-199) */
    if (const_step && (L.hSign * L.h) <= 0.0) { L.st = FLINT_ERROR_CONSTSTEPSZ; early = true; }   /* :99-103 */
    L.stiffMode = p.stifftest ? p.stifftest[idx] : p.stiff_mode;
    L.stiffOut = L.stiffMode;
    if (p.stiff_present && (L.stiffMode < 0 || L.stiffMode > 2)) { L.st = FLINT_ERROR_PARAMS; early = true; }   /* :143-154 */
    if (!p.stiff_present) L.stiffMode = 0;
    L.evmask = p.evmask_bits;
    L.eventsOn = p.events_on && L.evmask != 0u;                             /* :175-193 */
    if (early) {
        /* the reference returns before it touches Xf / Yf / StepSz.  Xf, StepSz and StiffTest keep the caller's values (they are
         * in/out); every pure output gets a defined value, because on the host-buffer path it is a staging buffer that is
         * copied back whole: Yf = Y0, counters and counts 0 */
        p.status[idx] = L.st;
        if (p.counters) { p.counters[0 * NN + idx] = 0; p.counters[1 * NN + idx] = 0; p.counters[2 * NN + idx] = 0; p.counters[3 * NN + idx] = 0; }
#pragma unroll
        for (int c = 0; c < N; ++c) p.yf[(long)c * NN + idx] = L.Y[c];
        if (p.ev_count) p.ev_count[idx] = 0;
        if (p.steps_count) p.steps_count[idx] = 0;
        if (p.dense_count) p.dense_count[idx] = 0;
        p.si[SI::FLAGS * p.scap + idx] = kFlagDone;
        return;
    }
    if (L.eventsOn) {                                                       /* :216-221 */
        double yy[N]; int da = 0;
#pragma unroll
        for (int c = 0; c < N; ++c) yy[c] = L.Y[c];
        sys.events(p.evset, L.X, yy, (1u << m) - 1u, L.EV0, L.EvDir, 0, da);
    }
    rhs_eval<SYS, FMA>(sys, L.Y, L.F0);                                     /* :224-225 */
    L.fcalls = 1;
    L.hmax = p.hmax_opt;
    if (L.hmax == 0.0) L.hmax = fabs(L.Xf - L.X);                           /* :237 */
    if (L.h == 0.0) {                                                       /* StepSz0Hairer, src/StepSz.f90:53-117 */
        double t0[N], t1[N], sc0[N];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            sc0[c] = madd<FMA>(p.tol.rt(c), fabs(L.Y[c]), p.tol.at(c));     /* :230-234 */
            t0[c] = L.Y[c] / sc0[c]; t1[c] = L.F0[c] / sc0[c];
        }
        const double d0 = norm2<FMA, N>(t0), d1 = norm2<FMA, N>(t1);
        double h0 = (d0 <= 1.0e-5 || d1 <= 1.0e-5) ? 1.0e-6 : (0.01 * d0) / d1;
        h0 = L.hSign * dmin(h0, L.hmax);
        double yh[N], F1[N];
#pragma unroll
        for (int c = 0; c < N; ++c) yh[c] = madd<FMA>(h0, L.F0[c], L.Y[c]);
        rhs_eval<SYS, FMA>(sys, yh, F1);
#pragma unroll
        for (int c = 0; c < N; ++c) t0[c] = (F1[c] - L.F0[c]) / sc0[c];
        const double d2 = norm2<FMA, N>(t0) / h0;
        const double dMax = dmax(d1, fabs(d2));
        double h1;
        if (dMax <= 1.0e-15) h1 = dmax(1.0e-6, fabs(h0) * 1.0e-3);
        else h1 = det_pow_dev(0.01 / dMax, 1.0 / (double)METHOD::P, p.powc);
        L.h = L.hSign * dmin(dmin(100.0 * fabs(h0), h1), L.hmax);
        L.fcalls += 1;
    }
    if (p.steps_on && p.steps_cap > 0) {                                    /* :254-257 */
        p.xint[0 * NN + idx] = L.X;
#pragma unroll
        for (int c = 0; c < N; ++c) p.yint[((long)c * p.steps_cap + 0) * NN + idx] = L.Y[c];
    }
    if (!keeps_going<SYS>(p, L)) store_results<SYS>(p, idx, L);
    else state_store<SYS>(p, idx, L, 0);
}

/* ===================================================================== the hot loop */
/* EVMODE 0: no events.  1: events located inside the lane loop.  2: steps with events to locate are parked.
 *
 * GANG (small shards: a few trajectories per lane, e.g. the 2^20-orbit ensemble split over 8 GPUs = 1.73 per lane).  Handing
 * whole trajectories to lanes then costs max-over-lanes: with 1.73 per lane every warp has a lane that integrates two, the
 * others idle inside live warps (ncu, 2^17 orbits: 23.7 of 32 lanes per instruction, 20.3 ms against 15.2 at the 2^20 rate,
 * profiles/r02_notes.md).  The gang-scheduled kernel time-shares the lanes instead: a CTA owns a contiguous n_items / grid of the
 * pass's items and works in ROUNDS of p.round_len step attempts.  Before a round the CTA selects, among its unfinished
 * trajectories, the blockDim.x that are furthest BEHIND in the independent variable (progress (X - X0) / (Xf - X0); histogram
 * threshold in shared memory); every lane loads one, advances it for the round and writes it back.  Trajectories that need more
 * attempts per unit of x are therefore scheduled more often and all of them reach Xf together: the lanes stay full until the
 * last rounds, and no lane ever parks or refills on its own (a first version that let lanes swap individually at x milestones
 * ran 7.9 % of its instructions with < 8 active lanes and pushed the CTA's warps out of step, no_instruction 0.3 -> 2.0).
 * Arithmetic per trajectory is untouched: results are bit-identical. */
template <class METHOD, class SYS, bool FMA, int EVMODE, bool GANG = false>
__global__ void __launch_bounds__(step_block<SYS>(), step_min_blocks<SYS>())
erk_step_kernel(const __grid_constant__ KArgs p)
{
    constexpr int N = SYS::N, M = SYS::M_MAX, PS = METHOD::PSTAR;
    constexpr bool EVENTS = EVMODE != 0;
    static_assert(!(GANG && EVMODE == 1), "the gang-scheduled kernel has no in-loop event handling");
    /* the plane variant and the general kernel are launched back to back; the flag erk_init_kernel left says which runs */
    if (p.zflag && ((*p.zflag != 0) != is_plane_variant<SYS>())) return;
    SYS sys;
    sys_load(sys, p);
    const int m = p.m;
    const bool const_step = p.const_step != 0;
    const double fac = const_step ? 1.0 : kLastStepExpFac;              /* :106-110 */
    const double SFl = p.szp[0], SFmin = p.szp[1], beta = p.szp[3], beta_mult = p.szp[4];
    const double expo = madd<FMA>(-beta, beta_mult, 1.0 / ((double)METHOD::Q + 1.0));   /* src/StepSz.f90:153 */
    const long n_items = p.list_in ? (long)*p.n_in : p.N;

    long idx = -1;
    bool have = false, exhausted = false, pending = false;
    Lane<SYS> L;
    L.X = 0; L.h = 0; L.Xf = 0; L.hSign = 1; L.hmax = 0; L.hbyhoptOld = 0;
    L.total = 0; L.acc = 0; L.rej = 0; L.fcalls = 0; L.st = FLINT_SUCCESS;
    L.LAST = false; L.lastRej = false; L.eventsOn = false;
    L.stiffMode = 0; L.stiffThr = 0; L.nonStiffThr = 0; L.stiffOut = 0;
    L.evmask = 0; L.evAction = FLINT_EVENTACTION_CONTINUE; L.evCount = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) { L.Y[i] = 0; L.F0[i] = 0; }
#pragma unroll
    for (int i = 0; i < M; ++i) { L.EV0[i] = 0; L.EvDir[i] = 0; }
    EvCtx<METHOD, SYS> ctx;                 /* EVMODE 1 only: local memory, touched when an event has to be located */
    double powKey = qnan(), powVal = 0.0;   /* pow(hbyhoptOld, beta) of this lane's trajectory (commit_tail) */
    extern __shared__ double flint_smem[];
    KStore<METHOD, SYS, METHOD::SMEM_SLOTS> ks;
    ks.s = flint_smem + threadIdx.x;

    /* ---- GANG: this CTA's items, their progress keys and the selection scratch (dynamic shared memory behind the KStore) ---- */
    constexpr unsigned FULLMASK = 0xffffffffu;
    __shared__ unsigned s_kmin, s_kmax, s_hist[256];
    __shared__ int s_nact, s_nsel, s_thr_bin, s_thr_take, s_taken;
    float* key = reinterpret_cast<float*>(flint_smem + kstore_smem_bytes<METHOD, SYS, METHOD::SMEM_SLOTS>() / 8);   /* [n_local]: progress in [0, 1]; 2 = retired */
    long lo_item = 0;
    int n_local = 0, jloc = -1, sweeps = 0;
    int* sel = nullptr;                                                /* [blockDim.x]: the round's selection */
    if (GANG) {
        const long per = (n_items + gridDim.x - 1) / gridDim.x;
        lo_item = (long)blockIdx.x * per;
        const long nl = n_items - lo_item;
        n_local = (int)(nl < 0 ? 0 : (nl > per ? per : nl));
        sel = reinterpret_cast<int*>(key + ((per + 1) & ~1L));
        sweeps = (n_local + (int)blockDim.x - 1) / (int)blockDim.x;
        for (int j = threadIdx.x; j < n_local; j += blockDim.x) {
            const long i = p.list_in ? (long)p.list_in[lo_item + j] : lo_item + j;
            FLINT_CHK(lo_item + j < n_items && i < p.N);
            float kf = 2.0f;
            if (i >= 0 && !(p.si[SI::FLAGS * p.scap + i] & kFlagDone)) {
                const double X0 = p.sd[SD<N, M>::X0 * p.scap + i];
                kf = (float)((p.sd[SD<N, M>::X * p.scap + i] - X0) / (p.xf[i] - X0));
                kf = kf >= 0.0f ? (kf <= 1.0f ? kf : 1.0f) : 0.0f;      /* NaN -> 0 */
            }
            key[j] = kf;
        }
    }

    unsigned iter = 0;
    bool finished = false;
    while (!finished) {                    /* GANG: one trip = one round; otherwise a single trip */
    int round_left = 0;
    if (GANG) {
        /* ---------------- select the blockDim.x unfinished trajectories with the least progress ---------------- */
        if (threadIdx.x == 0) { s_kmin = 0x7f800000u; s_kmax = 0u; s_nact = 0; s_nsel = 0; s_taken = 0; }
        if (threadIdx.x < 256) s_hist[threadIdx.x] = 0u;
        __syncthreads();
        {
            unsigned kmin = 0x7f800000u, kmax = 0u; int nact = 0;      /* non-negative floats order like their bit patterns */
            for (int j = threadIdx.x; j < n_local; j += blockDim.x) {
                const float kf = key[j];
                if (kf < 1.5f) { const unsigned u = __float_as_uint(kf); kmin = u < kmin ? u : kmin; kmax = u > kmax ? u : kmax; ++nact; }
            }
            kmin = __reduce_min_sync(FULLMASK, kmin); kmax = __reduce_max_sync(FULLMASK, kmax); nact = __reduce_add_sync(FULLMASK, nact);
            if ((threadIdx.x & 31) == 0 && nact) { atomicMin(&s_kmin, kmin); atomicMax(&s_kmax, kmax); atomicAdd(&s_nact, nact); }
        }
        __syncthreads();
        const int nact = s_nact;
        if (nact == 0) break;
        const float fmin = __uint_as_float(s_kmin), frange = __uint_as_float(s_kmax) - fmin;
        const float fscale = frange > 0.0f ? 255.0f / frange : 0.0f;
        const bool all = nact <= (int)blockDim.x;
        if (!all) {
            for (int j = threadIdx.x; j < n_local; j += blockDim.x) {
                const float kf = key[j];
                if (kf < 1.5f) atomicAdd(&s_hist[(int)((kf - fmin) * fscale)], 1u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {        /* first bin at which the running count reaches blockDim.x: one warp, 8 bins per lane */
                unsigned c[8], tot = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { c[q] = s_hist[threadIdx.x * 8 + q]; tot += c[q]; }
                unsigned incl = tot;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(FULLMASK, incl, d); if ((int)threadIdx.x >= d) incl += v; }
                unsigned run = incl - tot;                                /* trajectories in the bins before this lane's */
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (run < blockDim.x && run + c[q] >= blockDim.x) { s_thr_bin = threadIdx.x * 8 + q; s_thr_take = (int)(blockDim.x - run); }
                    run += c[q];
                }
            }
            __syncthreads();
        }
        {
            const int tb = all ? 256 : s_thr_bin, take = all ? 0 : s_thr_take;
            for (int w = 0; w < sweeps; ++w) {                         /* warp-aggregated compaction into sel[] */
                const int j = w * (int)blockDim.x + (int)threadIdx.x;
                bool pick = false;
                if (j < n_local) {
                    const float kf = key[j];
                    if (kf < 1.5f) {
                        const int bin = all ? 0 : (int)((kf - fmin) * fscale);
                        pick = bin < tb || (bin == tb && atomicAdd(&s_taken, 1) < take);
                    }
                }
                const unsigned mk = __ballot_sync(FULLMASK, pick);
                int base = 0;
                if ((threadIdx.x & 31) == 0 && mk) base = atomicAdd(&s_nsel, __popc(mk));
                base = __shfl_sync(FULLMASK, base, 0);
                if (pick) sel[base + __popc(mk & ((1u << (threadIdx.x & 31)) - 1u))] = j;
            }
        }
        __syncthreads();
#ifdef FLINT_GANG_TRACE
        if (threadIdx.x == 0 && blockIdx.x == 0) printf("round nact %d nsel %d fmin %.4f fmax %.4f clk %lld\n", nact, s_nsel, fmin, fmin + frange, clock64());
#endif
        have = false; jloc = -1;
        if ((int)threadIdx.x < s_nsel) {
            jloc = sel[threadIdx.x];
            const long i = p.list_in ? (long)p.list_in[lo_item + jloc] : lo_item + jloc;
            FLINT_CHK(jloc >= 0 && jloc < n_local && i >= 0 && i < p.N);
            state_load<SYS>(p, i, L);
            idx = i; have = true; powKey = qnan();
        }
        /* near the end a round outlasts most of its trajectories (lanes idle until the CTA meets again): shorter rounds there */
        round_left = fmin > 0.96f ? (p.round_len >> 2) + 1 : (fmin > 0.88f ? (p.round_len >> 1) + 1 : p.round_len);
    }
    for (;;) {
        /* ---------------- refill: take the next trajectory from the state block ---------------- */
        if (GANG) {
            if (round_left-- == 0 || !__any_sync(FULLMASK, have)) break;
        } else {
        if (!have && !exhausted) {
            const long k = (long)atomicAdd(p.work_counter, 1ULL);     /* every item through the counter (nvcc aggregates it per warp):
                                                                         the kernel does not depend on how many lanes were launched */
            if (k < n_items) {
                const long i = p.list_in ? (long)p.list_in[k] : k;
                FLINT_CHK(i < p.N);
                if (i >= 0) {
                    const int flags = state_load<SYS>(p, i, L);
                    if (!(flags & kFlagDone)) { idx = i; have = true; powKey = qnan(); }
                }
            } else exhausted = true;
        }
        /* one barrier per step attempt keeps the warps of a CTA in the same part of the loop, so they
         * share instruction-cache fills instead of evicting each other's lines (the hot loop is about the
         * size of the 32 KB L1.5 instruction cache).  The smallest loop -- plane variant without events, 27 KB --
         * fits on its own and runs 2 % faster without the barrier (measured), and so do systems whose right-hand side is a
         * handful of inlined operations (Lorenz DOP54: 83.8 -> 82.4 ms per 2^18 trajectories; with the barrier off two-body
         * DOP853 and the halo orbits lose 3-6 %). */
        /* STEP_BARRIER = K: the CTA meets (and votes on its exit) every K-th iteration; 0: never, each warp votes for itself */
        constexpr int kEvery = has_step_barrier<SYS>::value ? step_barrier_value<SYS>()
                                                            : ((FLINT_SYNC_BLOCK && !(EVMODE == 0 && SYS::INLINE_ACCEL)) ? 1 : 0);
        if (kEvery == 1) { if (__syncthreads_and(!have && exhausted)) { finished = true; break; } }
        else if (kEvery > 1) { if ((iter++ % (unsigned)(kEvery > 1 ? kEvery : 2)) == 0 && __syncthreads_and(!have && exhausted)) { finished = true; break; } }
        else if (__all_sync(0xffffffffu, !have && exhausted)) { finished = true; break; }
        }

        /* ---------------- EVMODE 1: finish an accepted step whose events have to be located ---------------- */
        if (EVMODE == 1 && pending) {
            deferred_event_commit<METHOD, SYS, FMA>(sys, p, ctx, idx, L.hSign);
            L.X = ctx.X; L.h = ctx.h; L.hbyhoptOld = ctx.hbyhoptOld; L.lastRej = ctx.lastRej; L.LAST = ctx.LAST; L.st = ctx.st;
            L.fcalls = ctx.fcalls; L.evmask = ctx.evmask; L.evAction = ctx.action; L.evCount = ctx.evcount;
#pragma unroll
            for (int i = 0; i < N; ++i) { L.Y[i] = ctx.Y[i]; L.F0[i] = ctx.F0old[i]; }
#pragma unroll
            for (int i = 0; i < M; ++i) { L.EV0[i] = ctx.EV0[i]; L.EvDir[i] = ctx.EvDir[i]; }
            pending = false;
            if (!keeps_going<SYS>(p, L)) { store_results<SYS>(p, idx, L); have = false; }
        }

        /* ---------------- InterpOn: every level of the dense store full -> suspend; the host adds a level and resumes ---------------- */
        if (have && p.dv.nlev > 0 && L.acc >= p.dense_avail && !(EVMODE == 1 && pending)) {
            state_store<SYS>(p, idx, L, 0);
            { const unsigned long long w = atomicAdd(p.n_ovf, 1ULL); FLINT_CHK(w < (unsigned long long)p.N && idx >= 0 && idx < p.N); p.ovf_list[w] = (int)idx; }
            have = false;
            if (GANG) key[jloc] = 2.0f;
        }
        /* ---------------- one step attempt for every lane that owns a trajectory (:263-637) ---------------- */
        if (have) {
            if (L.hSign * (madd<FMA>(fac, L.h, L.X) - L.Xf) > 0.0) { L.h = L.Xf - L.X; L.LAST = true; }   /* :268-271 */
            const double h = L.h;
            double Ynew[N], Ypre[N], Err;
            METHOD::template step<SYS, FMA, false, true>(sys, L.X, L.Y, L.F0, h, h, ks, Ynew, Ypre, p.tol, !const_step, Err, p.coef);   /* :274 */
            if constexpr (lazy_range<SYS>::value) {
                if (sys.bad) {      /* an operand left the range of the branch-free operators: the attempt again, stage-by-stage fallback */
                    sys.bad = false;
                    METHOD::template step<SYS, FMA, false, false>(sys, L.X, L.Y, L.F0, h, h, ks, Ynew, Ypre, p.tol, !const_step, Err, p.coef);
                }
            }
            L.total += 1;
            double hbyhopt = 1.0;
            if (!const_step) hbyhopt = det_pow_dev(Err, expo, p.powc);            /* src/StepSz.f90:153 */
            bool parked = false;
            if (Err <= 1.0) {                                               /* :283 */
                L.fcalls += METHOD::FC_ACC;
                L.acc += 1;
                if (p.step_once) L.LAST = true;                             /* :289 */
                bool bip_counted = false;                                   /* BipComputed :292 */

                if (L.stiffMode == 1 || L.stiffMode == 2) {                 /* CheckStiffness :299-308, :767-798 */
                    bool stiff = false;
                    if ((L.acc % kStiffTestSteps) == 0 || L.stiffThr > 0) {
                        double dk[N], dy[N];
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            dk[c] = ks.template get<METHOD::SLOT_KSINT>(c) - ks.template get<METHOD::SLOT_KSINT_M1>(c);
                            dy[c] = (METHOD::YSINT_ASSIGNED ? Ynew[c] : 0.0) - Ypre[c];
                        }
                        const double sN = norm2<FMA, N>(dk), sD = norm2<FMA, N>(dy);
                        double hLamb = 0.0;
                        if (sD > 0.0) hLamb = (fabs(h) * sN) / sD;
                        if (hLamb > 6.1) { L.nonStiffThr = 0; L.stiffThr += 1; if (L.stiffThr == 15) stiff = true; }
                        else { L.nonStiffThr += 1; if (L.nonStiffThr == 6) L.stiffThr = 0; }
                    }
                    if (stiff) {
                        L.stiffOut = -1;
                        if (L.stiffMode == 1) { L.st = FLINT_STIFF_PROBLEM; L.LAST = true; }
                    }
                }

                bool rare = false;
                if (EVENTS && L.eventsOn) {                                 /* :313-552 */
                    double EV1[M], yy[N];
#pragma unroll
                    for (int i = 0; i < M; ++i) EV1[i] = 0.0;
#pragma unroll
                    for (int c = 0; c < N; ++c) yy[c] = Ynew[c];
                    int da = 0;
                    sys.events(p.evset, L.X + h, yy, (1u << m) - 1u, EV1, L.EvDir, 0, da);   /* :324-325 */
                    unsigned det = 0, ss = 0;
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        if (i < m) {
                            if (detect_event((L.evmask >> i) & 1u, L.EvDir[i], L.EV0[i], EV1[i])) det |= (1u << i);   /* :326-328 */
                            if (!(p.ev_stepsz[i] == 0.0 || p.ev_stepsz[i] > L.hSign * h)) ss |= (1u << i);            /* :332-335 */
                        }
                    }
                    if (ss && !bip_counted) { L.fcalls += METHOD::FC_DENSE; bip_counted = true; }   /* :337-344, quirk Q1 (lazy) */
                    rare = det || (ss & L.evmask);
                    if (rare) {
                        if (EVMODE == 1) {     /* stash the step; it is committed at the top of the next lane iteration */
                            ctx.X = L.X; ctx.h = h; ctx.h_stage = h;
#pragma unroll
                            for (int i = 0; i < N; ++i) {
                                ctx.Y[i] = L.Y[i]; ctx.Y1[i] = Ynew[i]; ctx.F0old[i] = L.F0[i]; ctx.ENY[i] = yy[i];
                                ctx.F0new[i] = ks.template get<METHOD::SLOT_F0NEW>(i);
                            }
#pragma unroll
                            for (int i = 0; i < M; ++i) { ctx.EV0[i] = L.EV0[i]; ctx.EV1[i] = EV1[i]; ctx.EvDir[i] = L.EvDir[i]; }
                            ctx.evmask = L.evmask; ctx.det = det; ctx.ss = ss; ctx.action = L.evAction; ctx.st = L.st; ctx.fcalls = L.fcalls;
                            ctx.evcount = L.evCount; ctx.LAST = L.LAST; ctx.bip_counted = bip_counted; ctx.haveB = false;
                            ctx.Err = Err; ctx.hbyhopt = hbyhopt; ctx.hbyhoptOld = L.hbyhoptOld; ctx.hmax = L.hmax; ctx.acc = L.acc;
                            ctx.lastRej = L.lastRej;
                            pending = true;
                        } else {               /* EVMODE 2: park the trajectory with the data of this step */
                            using D = SD<N, M>;
                            const long c = p.scap;
                            state_store<SYS>(p, idx, L, bip_counted ? kFlagBipCounted : 0);
                            double* d = p.sd + idx;
                            d[D::HSTEP * c] = h; d[D::ERR * c] = Err; d[D::HBH * c] = hbyhopt;
#pragma unroll
                            for (int i = 0; i < N; ++i) { d[(D::Y1 + i) * c] = Ynew[i]; d[(D::F0N + i) * c] = ks.template get<METHOD::SLOT_F0NEW>(i); }
#pragma unroll
                            for (int i = 0; i < M; ++i) d[(D::EV1 + i) * c] = EV1[i];
                            p.si[SI::DET * c + idx] = (int)det; p.si[SI::SS * c + idx] = (int)ss;
                            { const unsigned long long w = atomicAdd(p.n_out, 1ULL); FLINT_CHK(w < (unsigned long long)p.N && idx >= 0 && idx < p.N); p.list_out[w] = (int)idx; }
                            parked = true; have = false;
                            if (GANG) key[jloc] = 2.0f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < M; ++i) L.EV0[i] = EV1[i];      /* :550 */
                    }
                }
                if (!rare) {
                    const bool dense = p.dv.nlev > 0;
                    if (dense) log_dense_step<SYS>(p, idx, L.acc, L.X, h, L.Y, L.F0);     /* :555-566; coefficients: erk_dense_kernel */
                    const int dense_mode = dense ? 2 : 0;
                    double hio = h, F0new[N];
                    ks.template load<METHOD::SLOT_F0NEW>(F0new);
                    commit_tail<METHOD, SYS, FMA, EVENTS>(p, idx, dense_mode, L.eventsOn, L.evAction, L.hSign, L.hmax, L.X, L.Y, L.F0, hio, L.hbyhoptOld,
                                                          L.lastRej, L.LAST, L.st, L.fcalls, L.acc, h, Ynew, F0new, Err, hbyhopt,
                                                          bip_counted, Ynew, h, powKey, powVal);
                    L.h = hio;
                }
            } else {                                                        /* :607-624 */
                L.fcalls += METHOD::FC_REJ;
                if (L.acc >= 1) L.rej += 1;
                L.h = h * dmax(SFmin, SFl / hbyhopt);                       /* src/StepSz.f90:176, :627 */
                L.lastRej = true;
                L.LAST = false;
                if (!const_step && fabs(L.h) * 0.01 <= fabs(L.X) * kEps) L.st = FLINT_ERROR_STEPSZ_TOOSMALL;   /* :632-635 */
            }
            if (!pending && !parked && !keeps_going<SYS>(p, L)) {
                store_results<SYS>(p, idx, L); have = false;
                if (GANG) key[jloc] = 2.0f;
            }
        }
    }
    if (GANG) {                            /* end of the round: back to the state block, with the new progress as the key */
        if (have) {
            state_store<SYS>(p, idx, L, 0);
            const double X0 = p.sd[SD<N, M>::X0 * p.scap + idx];
            float kf = (float)((L.X - X0) / (L.Xf - X0));
            key[jloc] = kf >= 0.0f ? (kf <= 1.0f ? kf : 1.0f) : 0.0f;
        }
        __syncthreads();                   /* state blocks and keys are read by other lanes of this CTA in the next round */
    }
    }
}

/* ===================================================================== parked trajectories: locate + commit */
template <class METHOD, class SYS, bool FMA>
__global__ void __launch_bounds__(128) erk_event_kernel(const __grid_constant__ KArgs p)
{
    constexpr int N = SYS::N, M = SYS::M_MAX;
    using D = SD<N, M>;
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long n_parked = (long)*p.n_in;
    if (j == 0) {                                   /* stream-ordered between two step passes: recycle their counters */
        *p.work_counter = 0ULL;
        *p.n_out = 0ULL;
    }
    if (j >= n_parked) return;
    const long idx = p.list_in[j];
    FLINT_CHK(idx >= 0 && idx < p.N);
    SYS sys;
    sys_load(sys, p);
    Lane<SYS> L;
    const int flags = state_load<SYS>(p, idx, L);
    const long c = p.scap;
    const double* d = p.sd + idx;
    EvCtx<METHOD, SYS> ctx;
    ctx.X = L.X; ctx.h = d[D::HSTEP * c]; ctx.h_stage = ctx.h;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ctx.Y[i] = L.Y[i]; ctx.Y1[i] = d[(D::Y1 + i) * c]; ctx.F0old[i] = L.F0[i]; ctx.ENY[i] = ctx.Y1[i];
        ctx.F0new[i] = d[(D::F0N + i) * c];
    }
#pragma unroll
    for (int i = 0; i < M; ++i) { ctx.EV0[i] = L.EV0[i]; ctx.EV1[i] = d[(D::EV1 + i) * c]; ctx.EvDir[i] = L.EvDir[i]; }
    ctx.evmask = L.evmask; ctx.det = (unsigned)p.si[SI::DET * c + idx]; ctx.ss = (unsigned)p.si[SI::SS * c + idx];
    ctx.action = L.evAction; ctx.st = L.st; ctx.fcalls = L.fcalls;
    ctx.evcount = L.evCount; ctx.LAST = L.LAST; ctx.bip_counted = (flags & kFlagBipCounted) != 0; ctx.haveB = false;
    ctx.Err = d[D::ERR * c]; ctx.hbyhopt = d[D::HBH * c]; ctx.hbyhoptOld = L.hbyhoptOld; ctx.hmax = L.hmax; ctx.acc = L.acc;
    ctx.lastRej = L.lastRej;
    deferred_event_commit<METHOD, SYS, FMA>(sys, p, ctx, idx, L.hSign);
    L.X = ctx.X; L.h = ctx.h; L.hbyhoptOld = ctx.hbyhoptOld; L.lastRej = ctx.lastRej; L.LAST = ctx.LAST; L.st = ctx.st;
    L.fcalls = ctx.fcalls; L.evmask = ctx.evmask; L.evAction = ctx.action; L.evCount = ctx.evcount;
#pragma unroll
    for (int i = 0; i < N; ++i) { L.Y[i] = ctx.Y[i]; L.F0[i] = ctx.F0old[i]; }
#pragma unroll
    for (int i = 0; i < M; ++i) { L.EV0[i] = ctx.EV0[i]; L.EvDir[i] = ctx.EvDir[i]; }
    if (!keeps_going<SYS>(p, L)) store_results<SYS>(p, idx, L);     /* the next pass skips it (done flag) */
    else state_store<SYS>(p, idx, L, 0);
}

/* ===================================================================== dense output: coefficients from the step logs */
/* One thread per (trajectory, accepted step); a CTA = 128 consecutive steps of one trajectory (blockIdx.x = trajectory,
 * blockIdx.y = block of steps), so its logs and its output are contiguous.  Every thread recomputes its step from the log, runs
 * the extra dense stages and builds the power-basis coefficients in registers (src/ERKStepInt.f90:250-323,473-563,659-700).  Then
 *   p.dense_eval == 0: the coefficient record is written to p.dvo (in place when dvo is dv);
 *   p.dense_eval != 0: FUSED Interpolate -- the coefficients go to shared memory and the CTA evaluates the grid points that lie
 *       in its 128 steps (interp_eval.cuh) straight into Yarr: Bip never exists in HBM (config 4: 58 GB less written and read
 *       per 2^18 orbits, and a store of 2n+2 instead of (pstar+1)*nI doubles per step).
 * resident CTAs per SM: 3 (168 registers) where the stage set and the prefetched log fit them (DOP54, Verner65E), else 2 --
 * measured per 2^16 planar CR3BP orbits, kernel alone (ncu): Verner65E 22.9 ms (3) / 25.8 (2); DOP853 14.0 (3, 184 bytes of
 * spills since the log prefetch) / 9.7 (2); Verner98R spills at 3 */
#ifndef FLINT_DENSE_MIN_BLOCKS
#define FLINT_DENSE_MIN_BLOCKS(METHOD) ((METHOD::NSLOT) <= 8 ? 3 : 2)
#endif
template <class METHOD> __host__ __device__ constexpr int dense_smem_doubles(int nI) { return dense_smem_doubles_ps(METHOD::PSTAR, nI); }
template <class METHOD, class SYS, bool FMA>
__global__ void __launch_bounds__(kDenseBlock, FLINT_DENSE_MIN_BLOCKS(METHOD)) erk_dense_kernel(const __grid_constant__ KArgs p)
{
    constexpr int N = SYS::N, PS = METHOD::PSTAR;
    /* plane variant and general kernel launched back to back, like the step kernels: the flag says which one runs.  The
     * plane variant skips the stage sums of the identically +0 components and forms their coefficients from +0 stage
     * values with the general expressions, so the zeros carry the same signs */
    if (p.zflag && ((*p.zflag != 0) != is_plane_variant<SYS>())) return;
    const int tid = threadIdx.x;
    extern __shared__ __align__(16) double dsm[];
    /* A CTA serves the trajectories blockIdx.x, blockIdx.x + gridDim.x, ...: a thread lives for one step (~10 us of arithmetic)
     * and used to spend a fifth of that waiting for its log (step count -> record, two DRAM round trips; ncu: 22 % of the stall
     * samples).  Now the NEXT trajectory's step count and this thread's log record of it travel (cp.async, into a slot of shared
     * memory behind the kernel's other buffers) while the current step is computed. */
    double* pf = dsm + ((p.dense_eval ? dense_smem_doubles_ps(PS, p.nI) : (p.dense_bulk ? kDenseBlock * p.dvo.rstride : 0)) + 1) / 2 * 2;
    double* myslot = pf + tid * kDensePfSlot;
    int* pfcount = reinterpret_cast<int*>(pf + kDenseBlock * kDensePfSlot);       /* [2]: step counts, alternating */
    const long cap0 = p.dv.cum[1];
    constexpr int nlog = 2 * N + 2;                                   /* X, h, Y, F0 (an event log's further fields: from the store, rare) */
    static_assert(nlog % 2 == 0 && nlog <= kDensePfSlot - 2, "a log fits its prefetch slot in 16-byte units");
    auto prefetch = [&](long tn, int par) {
        if (tn < p.N) {
            const double* src = p.dv.base[0] + (tn * cap0 + (long)blockIdx.y * kDenseBlock + tid) * (long)p.dv.rstride;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(myslot);
#pragma unroll
            for (int q = 0; q < nlog; q += 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + q * 8), "l"(src + q));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (kDensePfSlot - 2) * 8), "l"(src + p.dv.rstride - 1));
            if (tid == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(pfcount + par)), "l"(p.dense_count + tn));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int par = 0;
    prefetch((long)blockIdx.x, par);
    for (long t = blockIdx.x; t < p.N; t += gridDim.x, par ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                                  /* the step count thread 0 fetched; everybody is done with the previous trajectory */
    const long ntot = pfcount[par];
    const long nacc = ntot < p.dv.capacity() ? ntot : p.dv.capacity();
    /* this thread's log of the first block, out of its slot, before the next prefetch overwrites it */
    double lg[nlog], lkind;
#pragma unroll
    for (int q = 0; q < nlog; ++q) lg[q] = myslot[q];
    lkind = myslot[kDensePfSlot - 2];
    __syncwarp();
    prefetch(t + gridDim.x, par ^ 1);
    if (p.dense_eval && p.istatus[t] != FLINT_SUCCESS) continue;      /* this trajectory's grid failed the validation (:60-76) */
    /* gridDim.y covers level 0 of the store (launch_dense); the few trajectories that grew deeper levels have their further blocks
     * dealt round-robin over the same CTAs, instead of every trajectory paying empty CTAs for the deepest level anyone reached */
    for (long s0 = (long)blockIdx.y * kDenseBlock; s0 < nacc; s0 += (long)gridDim.y * kDenseBlock) {
    const bool first_blk = s0 == (long)blockIdx.y * kDenseBlock;     /* its logs are in lg[]; deeper blocks read the store */
    const int ns = (int)(nacc - s0 < kDenseBlock ? nacc - s0 : kDenseBlock);
    const int rs = (PS + 1) * p.nI, rsp = rs + (rs & 1);
    double* sB = dsm;                                       /* [128][rsp] */
    double* sX = dsm + kDenseBlock * rsp;                   /* [129] (+1 pad) */
    long* sG = reinterpret_cast<long*>(sX + kDenseBlock + 2);   /* [129] */
    static_assert(kDenseBlock == kEvalSteps, "the block's steps are what the evaluation walks");
    SYS sys;
    sys_load(sys, p);
    if (tid < ns) {
        const double* __restrict__ rec = p.dv.rec(t, s0 + tid);
        const double kind = first_blk ? lkind : rec[p.dv.rstride - 1];
        const double X = first_blk ? lg[0] : rec[0], h = first_blk ? lg[1] : rec[1];                /* a coefficient record: X0, X1 */
        if (kind == kDenseCoef) {                           /* built by an earlier pass */
            if (p.dense_eval) {
                sX[tid] = X; if (tid == ns - 1) sX[ns] = h;
                for (int q = 0; q < rs; ++q) sB[tid * rsp + q] = rec[2 + q];
            } else if (p.dense_bulk || p.dvo.base[0] != p.dv.base[0]) {
                double* __restrict__ o = p.dense_bulk ? dsm + (long)tid * p.dvo.rstride : p.dvo.rec(t, s0 + tid);
                o[0] = X; o[1] = h;
                for (int q = 0; q < rs; ++q) o[2 + q] = rec[2 + q];
                o[p.dvo.rstride - 1] = kDenseCoef;
            }
        } else {
            double Y[N], F0[N], Y1[N], Yn[N], Yp[N], e, B[PS + 1][N];
#pragma unroll
            for (int c = 0; c < N; ++c) { Y[c] = first_blk ? lg[2 + c] : rec[2 + c]; F0[c] = first_blk ? lg[2 + N + c] : rec[2 + N + c]; }
            double h_stage = h;
            const bool ev = kind == kDenseEvLog;
            if (ev) {                                       /* rare: straight from the store */
                h_stage = rec[2 + 2 * N];
#pragma unroll
                for (int c = 0; c < N; ++c) Y1[c] = rec[3 + 2 * N + c];
            }
            KStore<METHOD, SYS, 0u> ks;
            ks.s = nullptr;
            METHOD::template step<SYS, FMA, true, false>(sys, X, Y, F0, h_stage, h, ks, Yn, Yp, p.tol, false, e, p.coef);
            if (!ev) {
#pragma unroll
                for (int c = 0; c < N; ++c) Y1[c] = Yn[c];
            }
            METHOD::template dense_coeffs<SYS, FMA>(ks, Y, h, Y1, B, p.coef);
            if (p.dense_eval) {
                sX[tid] = X; if (tid == ns - 1) sX[ns] = X + h;              /* Xint(j+1) = X + h (:555-566) */
                if (p.nI == N) {
#pragma unroll
                    for (int j = 0; j <= PS; ++j)
#pragma unroll
                        for (int c = 0; c < N; ++c) sB[tid * rsp + j * N + c] = B[j][c];
                } else {
                    for (int ii = 0; ii < p.nI; ++ii) {
                        const int want = p.istates[ii] - 1;
#pragma unroll
                        for (int j = 0; j <= PS; ++j) {
                            double v = 0.0;
#pragma unroll
                            for (int c = 0; c < N; ++c) if (want == c) v = B[j][c];
                            sB[tid * rsp + j * p.nI + ii] = v;
                        }
                    }
                }
            } else store_coef_record<METHOD, SYS>(p.dense_bulk ? dsm + (long)tid * p.dvo.rstride : p.dvo.rec(t, s0 + tid), p.dvo.rstride,
                                                  p.nI, p.istates, X, X + h, B);
        }
    }
    if (!p.dense_eval) {
        if (p.dense_bulk) {
        /* the block's records are contiguous in the store and were assembled in shared memory in the store's own layout: ONE bulk
         * copy (TMA engine) writes them, instead of 29 scattered 16-byte stores per thread (ncu: lg_throttle 0.64 of 5.8 stall
         * cycles per issue with two warps per sub-partition).  Every thread has read its log by now: in place is safe. */
        __syncthreads();
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const unsigned bytes = (unsigned)ns * (unsigned)p.dvo.rstride * 8u;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(p.dvo.rec(t, s0)), "r"((unsigned)__cvta_generic_to_shared(dsm)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      /* shared memory may go once it has been read */
        }
        __syncthreads();                                    /* ... and be written again by the next block of this CTA */
        }
        continue;
    }
    __syncthreads();
    const double hs = sign1(sX[1] - sX[0]);                 /* every step points in the direction of integration */
    const double* __restrict__ xg = p.ix + (p.ix_per_traj ? t * p.iG : 0L);
    double* __restrict__ yo = p.iy + t * p.iG * p.nI;
    block_step_ranges(sX, sG, ns, s0 == 0, s0 + ns == ntot, xg, p.iG, hs);
    switch (p.nI) {
    case 1: eval_block_steps<FMA, PS, 1, false>(sB, rsp, sX, sG, ns, xg, p.iG, yo, 0); break;
    case 2: eval_block_steps<FMA, PS, 2, false>(sB, rsp, sX, sG, ns, xg, p.iG, yo, 0); break;
    case 3: eval_block_steps<FMA, PS, 3, false>(sB, rsp, sX, sG, ns, xg, p.iG, yo, 0); break;
    case 4: eval_block_steps<FMA, PS, 4, false>(sB, rsp, sX, sG, ns, xg, p.iG, yo, 0); break;
    case 5: eval_block_steps<FMA, PS, 5, false>(sB, rsp, sX, sG, ns, xg, p.iG, yo, 0); break;
    case 6: if constexpr (N >= 6) eval_block_steps<FMA, PS, 6, false>(sB, rsp, sX, sG, ns, xg, p.iG, yo, 0); break;
    default: break;
    }
    __syncthreads();                                        /* the next block of this CTA overwrites what was just read */
    }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

#ifndef __CUDACC_RTC__
/* Host side: the table entry of one instantiation (inst_*.cu; kargs.h: KernelEntry).  The launchers themselves are not
 * templates (registry.cu): they take the entry points and the compile-time facts collected here, so that kernels NVRTC builds
 * at run time from a user's source (rtc.cu) are launched by the same code. */
template <class METHOD, class SYS, bool FMA, int EVMODE>
StepKernel step_kernel_entry()
{
    StepKernel k{};
    k.classic = (const void*)&erk_step_kernel<METHOD, SYS, FMA, EVMODE, false>;
    if constexpr (EVMODE != 1) k.gang = (const void*)&erk_step_kernel<METHOD, SYS, FMA, EVMODE, true>;
    k.block = step_block<SYS>();
    k.smem = kstore_smem_bytes<METHOD, SYS, METHOD::SMEM_SLOTS>();
    return k;
}
/* SYSZ: the plane variant of SYS (systems.cuh: zc), or SYS itself when it has none */
template <class METHOD, class SYS, class SYSZ, bool FMA, bool EVENTS>
KernelEntry kernel_entry()
{
    KernelEntry e{};
    e.method = METHOD::ID; e.system = SYS::ID; e.fma = FMA; e.events = EVENTS;
    e.init = (const void*)&erk_init_kernel<METHOD, SYS, FMA>;
    e.step = step_kernel_entry<METHOD, SYS, FMA, EVENTS ? 1 : 0>();
    if constexpr (EVENTS) {
        e.step_park = step_kernel_entry<METHOD, SYS, FMA, 2>();
        e.event = (const void*)&erk_event_kernel<METHOD, SYS, FMA>;
    }
    e.dense_coeffs = (const void*)&erk_dense_kernel<METHOD, SYS, FMA>;
    if constexpr (SYS::ZPLANE != 0u && !std::is_same<SYS, SYSZ>::value) {
        e.step_plane = step_kernel_entry<METHOD, SYSZ, FMA, EVENTS ? 2 : 0>();
        e.dense_plane = (const void*)&erk_dense_kernel<METHOD, SYSZ, FMA>;
    }
    e.pstar = METHOD::PSTAR;
    e.n_state_doubles = SD<SYS::N, SYS::M_MAX>::COUNT; e.n_state_ints = SI::COUNT;
    e.sys_n = SYS::N; e.sys_m_max = SYS::M_MAX;
    return e;
}
#endif

}  // namespace flint
#endif
