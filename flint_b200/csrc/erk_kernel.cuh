/* erk_kernel.cuh -- the ensemble integrator kernel (one trajectory per thread).
 *
 * Device restatement of ERK_class%Integrate (src/ERKIntegrate.f90:29-802) for N
 * independent trajectories.  B200-first structure, not the reference's loop nest:
 *
 *  * persistent lanes: a thread owns one trajectory at a time; when it finishes it
 *    pulls the next index from a global counter, so lanes whose trajectories need
 *    fewer steps do not idle until the slowest lane of the warp is done;
 *  * the adaptive loop is flattened (one step ATTEMPT per warp iteration for every
 *    lane that owns a trajectory), accept/reject is predicated data flow, the
 *    warp leaves the loop on __all_sync(no lane has work);
 *  * state, stage vectors and dense coefficients live in registers (generated
 *    straight-line stage code, erk_methods_gen.cuh); HBM traffic is the SoA
 *    load of Y0 and the SoA store of the results, once per trajectory;
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

#include "erk_methods_gen.cuh"
#include "systems.cuh"

namespace flint {

constexpr double kEps = 2.220446049250313e-16;   /* FLINT_EPS = epsilon(1.0_WP), src/FLINTUtils.f90:49 */
constexpr int kRootMaxItr = 50;                  /* FLINT_ROOTFIND_MAXITR, src/FLINTUtils.f90:52 */
constexpr int kStiffTestSteps = 1000;            /* STIFFTEST_STEPS, src/ERK.f90:53 */
constexpr double kLastStepExpFac = 1.01;         /* LASTSTEP_EXPFAC, src/ERK.f90:43 */
constexpr int kMaxSS = FLINT_MAXNUMSTEPSZEVENTS; /* MAXNUMSTEPSZEVENTS, src/ERK.f90:48 */

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
    double ks[METHOD::NSLOT][N], Yn[N], Yp[N], e;
    METHOD::template step<SYS, FMA, true>(sys, c.X, c.Y, c.F0old, c.h_stage, c.h, ks, Yn, Yp, p.tol, false, e, p.coef);
    METHOD::template dense_coeffs<SYS, FMA>(ks, c.Y, c.h, c.Y1, c.B, p.coef);
    c.haveB = true;
}
template <class METHOD, class SYS, bool FMA>
__device__ __forceinline__ void ensure_bip(const SYS& sys, const KArgs& p, EvCtx<METHOD, SYS>& c)
{
    if (!c.bip_counted) { c.fcalls += METHOD::FC_DENSE; c.bip_counted = true; }   /* BipComputed / FCalls as the reference */
    remat_bip<METHOD, SYS, FMA>(sys, p, c);
}
/* dense coefficients of an event-truncated step for the InterpOn store (rare) */
template <class METHOD, class SYS, bool FMA>
__device__ __noinline__ void rare_dense(const SYS& sys, const KArgs& p, EvCtx<METHOD, SYS>& c)
{
    c.haveB = false;
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
            double V0[M], V1[M];
            for (int i = 0; i < M; ++i) { V0[i] = EV0[i][0]; V1[i] = EV1[i][0]; }
            const double Eventh = hSign * p.ev_stepsz[ev];
            EventX1 = X + Eventh;
            const double Vh = EV1[ev][0];
            c.det &= ~(1u << ev);
            bool last = false;
            while (EventX1 <= (X + c.h)) {                               /* :372 (quirk Q7) */
                if (EventX1 != (X + c.h)) {
                    interp_y<FMA, PS, N>(EventX1, X, c.h, c.B, c.ENY);
                    sys.events(p.evset, EventX1, c.ENY, 1u << ev, V1, c.EvDir, 0, dummy_action);
                } else V1[ev] = Vh;
                if (detect_event((c.evmask >> ev) & 1u, c.EvDir[ev], V0[ev], V1[ev])) {
                    c.det |= (1u << ev);
                    const int slot = nSS[ev] - 1;
                    EX0[ev][slot] = EventX0; EX1[ev][slot] = EventX1; EV0[ev][slot] = V0[ev]; EV1[ev][slot] = V1[ev];
                    nSS[ev] = nSS[ev] + 1;
                    if (nSS[ev] > kMaxSS) break;
                }
                if (last) break;
                EventX0 = EventX1;
                EventX1 = EventX1 + Eventh;
                if (EventX1 > X + c.h) { EventX1 = X + c.h; last = true; }
                V0[ev] = V1[ev];
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

/* ---- the tail of an accepted step (:555-635): dense store, stale-action quirk, natural-step output,
 * advance, new step size, too-small test.  Shared by the hot path (state in registers) and by the
 * deferred event commit (state in the local-memory context). ---- */
template <class METHOD, class SYS, bool FMA, bool EVENTS, bool DENSE>
__device__ __forceinline__ void commit_tail(const KArgs& p, long idx, bool eventsOn, int evAction, double hSign, double hmax,
        double& X, double (&Y)[SYS::N], double (&F0)[SYS::N], double& h_io, double& hbyhoptOld, bool& lastRej, bool LAST, int& st,
        int& fcalls, int acc, double h, double (&Y1)[SYS::N], const double (&F0new)[SYS::N], double Err, double hbyhopt,
        bool bip_counted, const double (&ENY)[SYS::N], const double (&B)[METHOD::PSTAR + 1][SYS::N])
{
    constexpr int N = SYS::N, PS = METHOD::PSTAR;
    const long NN = p.N;
    const bool const_step = p.const_step != 0;
    if constexpr (DENSE) {                                              /* :555-566 */
        if (!bip_counted) { fcalls += METHOD::FC_DENSE; bip_counted = true; }
        if (acc <= p.dense_cap) {
            p.dxint[(long)acc * NN + idx] = X + h;
            for (int j = 0; j <= PS; ++j)
                for (int ii = 0; ii < p.nI; ++ii) {
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < N; ++c) if (p.istates[ii] - 1 == c) v = B[j][c];
                    p.dbip[(((long)j * p.nI + ii) * p.dense_cap + (acc - 1)) * NN + idx] = v;
                }
        }
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
        if (beta != 0.0) hbyhopt = hbyhopt / flint_det_pow(hbyhoptOld, beta);
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
template <class METHOD, class SYS, bool FMA, bool DENSE>
__device__ __noinline__ void deferred_event_commit(const SYS& sys, const KArgs& p, EvCtx<METHOD, SYS>& c, long idx, double hSign)
{
    constexpr int N = SYS::N, M = SYS::M_MAX, PS = METHOD::PSTAR;
    event_rare_path<METHOD, SYS, FMA>(sys, p, c, idx, hSign);
#pragma unroll
    for (int i = 0; i < M; ++i) c.EV0[i] = c.EV1[i];                  /* :550 */
    if constexpr (DENSE) {
        if (c.h != c.h_stage) c.haveB = false;                         /* truncated: rebuild the dense stages (quirk Q6) */
        remat_bip<METHOD, SYS, FMA>(sys, p, c);
    }
    double Xn = c.X, hio = c.h;
    commit_tail<METHOD, SYS, FMA, true, DENSE>(p, idx, true, c.action, hSign, c.hmax, Xn, c.Y, c.F0old, hio, c.hbyhoptOld, c.lastRej,
                                               c.LAST, c.st, c.fcalls, c.acc, c.h, c.Y1, c.F0new, c.Err, c.hbyhopt, c.bip_counted, c.ENY, c.B);
    c.X = Xn; c.h = hio;       /* new X, next step size; new Y in c.Y, new F0 in c.F0old */
    (void)PS;
}

/* ===================================================================== */
template <class METHOD, class SYS, bool FMA, bool EVENTS, bool DENSE>
__global__ void __launch_bounds__(FLINT_BLOCK_THREADS, FLINT_MIN_BLOCKS)
erk_integrate_kernel(const __grid_constant__ KArgs p)
{
    constexpr int N = SYS::N, M = SYS::M_MAX, PS = METHOD::PSTAR;
    SYS sys;
    sys.load(p.sp);
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long NN = p.N;
    const int m = p.m;
    const bool const_step = p.const_step != 0;
    const double fac = const_step ? 1.0 : kLastStepExpFac;              /* :106-110 */
    const double SFl = p.szp[0], SFmin = p.szp[1], beta = p.szp[3], beta_mult = p.szp[4];
    const double expo = madd<FMA>(-beta, beta_mult, 1.0 / ((double)METHOD::Q + 1.0));   /* src/StepSz.f90:153 */

    /* per-trajectory (loop-carried) state */
    long idx = -1;
    bool have = false, exhausted = false, first = true, pending = false;
    double X = 0, h = 0, Xf = 0, hSign = 1, hmax = 0, hbyhoptOld = 0;
    double Y[N], F0[N];
    int total = 0, acc = 0, rej = 0, fcalls = 0, st = FLINT_SUCCESS;
    bool LAST = false, lastRej = false;
    int stiffMode = 0, stiffThr = 0, nonStiffThr = 0, stiffOut = 0;
    double EV0[M];
    int EvDir[M];
    unsigned evmask = 0;
    int evAction = FLINT_EVENTACTION_CONTINUE, evCount = 0;
    bool eventsOn = false;
    EvCtx<METHOD, SYS> ctx;                 /* local memory; touched only when an event has to be located */
#pragma unroll
    for (int i = 0; i < N; ++i) { Y[i] = 0; F0[i] = 0; }
#pragma unroll
    for (int i = 0; i < M; ++i) { EV0[i] = 0; EvDir[i] = 0; }

    /* final status resolution and result store (:642-708) */
    auto finalize = [&]() {
        int fs = st;
        if (LAST && st == FLINT_SUCCESS) {}
        else if (eventsOn && LAST && st == FLINT_EVENT_TERM) {}
        else if (total >= p.max_steps && st == FLINT_SUCCESS) fs = FLINT_ERROR_MAXSTEPS;
        else if (st == FLINT_ERROR_STEPSZ_TOOSMALL) {}
        else if (eventsOn && st == FLINT_ERROR_EVENTROOT) {}
        else if (st == FLINT_STIFF_PROBLEM) {}
        else fs = FLINT_ERROR;
        p.xf[idx] = X;
#pragma unroll
        for (int i = 0; i < N; ++i) p.yf[(long)i * NN + idx] = Y[i];
        if (!const_step && p.stepsz) p.stepsz[idx] = h;
        p.status[idx] = fs;
        if (p.counters) {
            p.counters[0 * NN + idx] = total; p.counters[1 * NN + idx] = acc;
            p.counters[2 * NN + idx] = rej; p.counters[3 * NN + idx] = fcalls;
        }
        if (p.stifftest) p.stifftest[idx] = stiffOut;
        if (p.ev_count) p.ev_count[idx] = evCount;
        if (p.steps_count) p.steps_count[idx] = acc + 1;
        if (DENSE && p.dense_count) p.dense_count[idx] = acc;
        have = false;
    };

    for (;;) {
        /* ---------------- refill: take the next trajectory and initialise it (:204-259) ---------------- */
        if (!have && !exhausted) {
            const long i = first ? tid : (long)atomicAdd(p.work_counter, 1ULL);
            first = false;
            if (i < NN) {
                idx = i;
                have = true;
                X = p.x0 ? p.x0[idx] : p.x0_scalar;
#pragma unroll
                for (int c = 0; c < N; ++c) Y[c] = p.y0[(long)c * NN + idx];
                Xf = p.xf[idx];
                total = 0; acc = 0; rej = 0; fcalls = 0; st = FLINT_SUCCESS;
                LAST = false; lastRej = false; stiffThr = 0; nonStiffThr = 0;
                hbyhoptOld = p.szp[5];
                evAction = FLINT_EVENTACTION_CONTINUE; evCount = 0;
                hSign = sign1(Xf - X);                                      /* :94 */
                h = p.stepsz ? p.stepsz[idx] : 0.0;
                bool early = false;   /* argument errors: the reference returns before touching Xf/Yf (:196-199) */
                if (const_step && (hSign * h) <= 0.0) { st = FLINT_ERROR_CONSTSTEPSZ; early = true; }   /* :99-103 */
                stiffMode = p.stifftest ? p.stifftest[idx] : p.stiff_mode;
                stiffOut = stiffMode;
                if (p.stiff_present && (stiffMode < 0 || stiffMode > 2)) { st = FLINT_ERROR_PARAMS; early = true; }   /* :143-154 */
                if (!p.stiff_present) stiffMode = 0;
                evmask = p.evmask_bits;
                eventsOn = EVENTS && p.events_on && evmask != 0u;          /* :175-193 */
                if (early) {
                    p.status[idx] = st;
                    if (p.counters) { p.counters[0 * NN + idx] = 0; p.counters[1 * NN + idx] = 0; p.counters[2 * NN + idx] = 0; p.counters[3 * NN + idx] = 0; }
                    have = false;
                } else {
                    if (EVENTS && eventsOn) {                               /* :216-221 */
                        double yy[N]; int da = 0;
#pragma unroll
                        for (int c = 0; c < N; ++c) yy[c] = Y[c];
                        sys.events(p.evset, X, yy, (1u << m) - 1u, EV0, EvDir, 0, da);
                    }
                    {                                                       /* :224-225 */
                        rhs_eval<SYS, FMA>(sys, Y, F0);
                    }
                    fcalls = 1;
                    hmax = p.hmax_opt;
                    if (hmax == 0.0) hmax = fabs(Xf - X);                   /* :237 */
                    if (h == 0.0) {                                         /* StepSz0Hairer, src/StepSz.f90:53-117 */
                        double t0[N], t1[N], sc0[N];
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            sc0[c] = madd<FMA>(p.tol.rt(c), fabs(Y[c]), p.tol.at(c));   /* :230-234 */
                            t0[c] = Y[c] / sc0[c]; t1[c] = F0[c] / sc0[c];
                        }
                        const double d0 = norm2<FMA, N>(t0), d1 = norm2<FMA, N>(t1);
                        double h0 = (d0 <= 1.0e-5 || d1 <= 1.0e-5) ? 1.0e-6 : (0.01 * d0) / d1;
                        h0 = hSign * dmin(h0, hmax);
                        double yh[N], F1[N];
#pragma unroll
                        for (int c = 0; c < N; ++c) yh[c] = madd<FMA>(h0, F0[c], Y[c]);
                        rhs_eval<SYS, FMA>(sys, yh, F1);
#pragma unroll
                        for (int c = 0; c < N; ++c) t0[c] = (F1[c] - F0[c]) / sc0[c];
                        const double d2 = norm2<FMA, N>(t0) / h0;
                        const double dMax = dmax(d1, fabs(d2));
                        double h1;
                        if (dMax <= 1.0e-15) h1 = dmax(1.0e-6, fabs(h0) * 1.0e-3);
                        else h1 = flint_det_pow(0.01 / dMax, 1.0 / (double)METHOD::P);
                        h = hSign * dmin(dmin(100.0 * fabs(h0), h1), hmax);
                        fcalls += 1;
                    }
                    if (p.steps_on && p.steps_cap > 0) {                    /* :254-257 */
                        p.xint[0 * NN + idx] = X;
#pragma unroll
                        for (int c = 0; c < N; ++c) p.yint[((long)c * p.steps_cap + 0) * NN + idx] = Y[c];
                    }
                    if (DENSE && p.dense_cap > 0) p.dxint[0 * NN + idx] = X;   /* :259 */
                    if (!(st == FLINT_SUCCESS && !LAST && total < p.max_steps)) finalize();
                }
            } else exhausted = true;
        }
        if (__all_sync(0xffffffffu, !have && exhausted)) break;

        /* ---------------- deferred: finish an accepted step whose events have to be located ---------------- */
        if (EVENTS && pending) {
            deferred_event_commit<METHOD, SYS, FMA, DENSE>(sys, p, ctx, idx, hSign);
            X = ctx.X; h = ctx.h; hbyhoptOld = ctx.hbyhoptOld; lastRej = ctx.lastRej; LAST = ctx.LAST; st = ctx.st;
            fcalls = ctx.fcalls; evmask = ctx.evmask; evAction = ctx.action; evCount = ctx.evcount;
#pragma unroll
            for (int i = 0; i < N; ++i) { Y[i] = ctx.Y[i]; F0[i] = ctx.F0old[i]; }
#pragma unroll
            for (int i = 0; i < M; ++i) { EV0[i] = ctx.EV0[i]; EvDir[i] = ctx.EvDir[i]; }
            pending = false;
            if (!(st == FLINT_SUCCESS && !LAST && total < p.max_steps)) finalize();
        }

        /* ---------------- one step attempt for every lane that owns a trajectory (:263-637) ---------------- */
        if (have) {
            if (hSign * (madd<FMA>(fac, h, X) - Xf) > 0.0) { h = Xf - X; LAST = true; }   /* :268-271 */
            double ks[METHOD::NSLOT][N], Ynew[N], Ypre[N], Err;
            METHOD::template step<SYS, FMA, DENSE>(sys, X, Y, F0, h, h, ks, Ynew, Ypre, p.tol, !const_step, Err, p.coef);   /* :274 */
            total += 1;
            double hbyhopt = 1.0;
            if (!const_step) hbyhopt = flint_det_pow(Err, expo);            /* src/StepSz.f90:153 */
            if (Err <= 1.0) {                                               /* :283 */
                fcalls += METHOD::FC_ACC;
                acc += 1;
                if (p.step_once) LAST = true;                               /* :289 */
                bool bip_counted = false;                                   /* BipComputed :292 */

                if (stiffMode == 1 || stiffMode == 2) {                     /* CheckStiffness :299-308, :767-798 */
                    bool stiff = false;
                    if ((acc % kStiffTestSteps) == 0 || stiffThr > 0) {
                        double dk[N], dy[N];
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            dk[c] = ks[METHOD::SLOT_KSINT][c] - ks[METHOD::SLOT_KSINT_M1][c];
                            dy[c] = (METHOD::YSINT_ASSIGNED ? Ynew[c] : 0.0) - Ypre[c];
                        }
                        const double sN = norm2<FMA, N>(dk), sD = norm2<FMA, N>(dy);
                        double hLamb = 0.0;
                        if (sD > 0.0) hLamb = (fabs(h) * sN) / sD;
                        if (hLamb > 6.1) { nonStiffThr = 0; stiffThr += 1; if (stiffThr == 15) stiff = true; }
                        else { nonStiffThr += 1; if (nonStiffThr == 6) stiffThr = 0; }
                    }
                    if (stiff) {
                        stiffOut = -1;
                        if (stiffMode == 1) { st = FLINT_STIFF_PROBLEM; LAST = true; }
                    }
                }

                bool rare = false;
                if (EVENTS && eventsOn) {                                   /* :313-552 */
                    double EV1[M], yy[N];
#pragma unroll
                    for (int i = 0; i < M; ++i) EV1[i] = 0.0;
#pragma unroll
                    for (int c = 0; c < N; ++c) yy[c] = Ynew[c];
                    int da = 0;
                    sys.events(p.evset, X + h, yy, (1u << m) - 1u, EV1, EvDir, 0, da);   /* :324-325 */
                    unsigned det = 0, ss = 0;
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        if (i < m) {
                            if (detect_event((evmask >> i) & 1u, EvDir[i], EV0[i], EV1[i])) det |= (1u << i);   /* :326-328 */
                            if (!(p.ev_stepsz[i] == 0.0 || p.ev_stepsz[i] > hSign * h)) ss |= (1u << i);        /* :332-335 */
                        }
                    }
                    if (ss && !bip_counted) { fcalls += METHOD::FC_DENSE; bip_counted = true; }   /* :337-344, quirk Q1 (lazy) */
                    rare = det || (ss & evmask);
                    if (rare) {            /* stash the step; it is committed at the top of the next lane iteration */
                        ctx.X = X; ctx.h = h; ctx.h_stage = h;
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            ctx.Y[i] = Y[i]; ctx.Y1[i] = Ynew[i]; ctx.F0old[i] = F0[i]; ctx.ENY[i] = yy[i];
                            ctx.F0new[i] = ks[METHOD::SLOT_F0NEW][i];
                        }
#pragma unroll
                        for (int i = 0; i < M; ++i) { ctx.EV0[i] = EV0[i]; ctx.EV1[i] = EV1[i]; ctx.EvDir[i] = EvDir[i]; }
                        ctx.evmask = evmask; ctx.det = det; ctx.ss = ss; ctx.action = evAction; ctx.st = st; ctx.fcalls = fcalls;
                        ctx.evcount = evCount; ctx.LAST = LAST; ctx.bip_counted = bip_counted; ctx.haveB = false;
                        ctx.Err = Err; ctx.hbyhopt = hbyhopt; ctx.hbyhoptOld = hbyhoptOld; ctx.hmax = hmax; ctx.acc = acc; ctx.lastRej = lastRej;
                        pending = true;
                    } else {
#pragma unroll
                        for (int i = 0; i < M; ++i) EV0[i] = EV1[i];        /* :550 */
                    }
                }
                if (!rare) {
                    double B[PS + 1][N];       /* dead unless DENSE */
                    if constexpr (DENSE) METHOD::template dense_coeffs<SYS, FMA>(ks, Y, h, Ynew, B, p.coef);
                    double hio = h;
                    commit_tail<METHOD, SYS, FMA, EVENTS, DENSE>(p, idx, eventsOn, evAction, hSign, hmax, X, Y, F0, hio, hbyhoptOld, lastRej, LAST, st,
                                                                 fcalls, acc, h, Ynew, ks[METHOD::SLOT_F0NEW], Err, hbyhopt, bip_counted, Ynew, B);
                    h = hio;
                }
            } else {                                                        /* :607-624 */
                fcalls += METHOD::FC_REJ;
                if (acc >= 1) rej += 1;
                h = h * dmax(SFmin, SFl / hbyhopt);                         /* src/StepSz.f90:176, :627 */
                lastRej = true;
                LAST = false;
                if (!const_step && fabs(h) * 0.01 <= fabs(X) * kEps) st = FLINT_ERROR_STEPSZ_TOOSMALL;   /* :632-635 */
            }
            if (!pending && !(st == FLINT_SUCCESS && !LAST && total < p.max_steps)) finalize();
        }
    }
}

/* Host-side launcher signature shared by all instantiations (inst_*.cu). */

template <class METHOD, class SYS, bool FMA, bool EVENTS, bool DENSE>
cudaError_t launch_erk(const KArgs& a, int grid, int block, cudaStream_t s)
{
    KArgs b = a;
    METHOD::load_coefficients(b.coef);
    erk_integrate_kernel<METHOD, SYS, FMA, EVENTS, DENSE><<<grid, block, 0, s>>>(b);
    return cudaGetLastError();
}
template <class METHOD, class SYS, bool FMA, bool EVENTS, bool DENSE>
cudaError_t occupancy_erk(int* blocks_per_sm, int block)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, erk_integrate_kernel<METHOD, SYS, FMA, EVENTS, DENSE>, block, 0);
}

}  // namespace flint
#endif
