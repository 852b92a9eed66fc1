/* flint_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A line-by-line C++ restatement of the reference's fp64 algorithm (FLINT,
 * /root/reference/src/*.f90).  Nothing in the product (flint_b200/, include/)
 * may include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker
 * or the CPU baseline.
 *
 * The reference is pure Fortran and no Fortran compiler exists in the build
 * image, so it cannot be compiled into oracle/_ref (DESIGN.md "Oracle").  This
 * restatement is pinned instead against the reference's own golden result files
 * (tests/wp_*.txt, tests/results_CR3BP.txt; see tests/test_oracle_golden.py).
 *
 * Every routine cites the reference lines it restates.  Expression ORDER follows
 * the Fortran exactly (left-to-right sums, same parenthesisation) because step
 * counts on the benchmark problems are sensitive to single-ulp changes (SURVEY.md
 * §7.3 H1).  Two arithmetic modes, selected per object:
 *   FMA = false ("strict"): every * and + rounds separately -- what
 *          `gfortran -O3` emits for x86-64 (no FMA in the baseline ISA).
 *   FMA = true : a*b+c sites of the form documented in DESIGN.md ("arithmetic
 *          contract") are fused.  Mirrors the kernels' throughput mode.
 * and two pow implementations: libm pow (what a gfortran build calls) and the
 * deterministic pow of flint_b200/csrc/det_pow.h (the one arithmetic routine the
 * oracle shares with the product, because bit parity with the GPU needs a pow
 * that returns identical bits on both sides).
 *
 * Intrinsics whose code is not in the reference tree are restated as:
 *   norm2(v)   -> sqrt(v1*v1 + v2*v2 + ...) accumulated left to right
 *   x**k (int) -> libgcc __powidf2 square-and-multiply order
 *   sum(v)     -> 0 + v1 + v2 + ... left to right
 *   matmul(K,b)-> per component, sum over stages left to right from 0
 * Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -mfma -fopenmp).
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <limits>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FLINT_TABLEAU_CONST static const
#include "oracle_tableaus.h"
#include "../flint_b200/csrc/det_pow.h"   /* shared arithmetic-contract pow only */

namespace {

/* ---- enums: src/FLINT_base.f90:43-61,86-102,107-120; src/ERK.f90:56-66 ---- */
enum { FLINT_SUCCESS = 0, FLINT_EVENT_TERM = 1, FLINT_STIFF_PROBLEM = 2, FLINT_ERROR = 3,
       FLINT_ERROR_PARAMS = 4, FLINT_ERROR_MAXSTEPS = 5, FLINT_ERROR_MEMALLOC = 6,
       FLINT_ERROR_MEMDEALLOC = 7, FLINT_ERROR_DIMENSION = 8, FLINT_ERROR_STEPSZ_TOOSMALL = 9,
       FLINT_ERROR_INTPSTATES = 10, FLINT_ERROR_INTP_ARRAY = 11, FLINT_ERROR_INTP_OFF = 12,
       FLINT_ERROR_INIT_REQD = 13, FLINT_ERROR_EVENTPARAMS = 14, FLINT_ERROR_EVENTROOT = 15,
       FLINT_ERROR_CONSTSTEPSZ = 16 };
enum { EVACT_CONTINUE = 0, EVACT_TERMINATE = 1, EVACT_CHANGESOLY = 2, EVACT_RESTARTINT = 4, EVACT_MASK = 128 };
enum { EVOPT_ROOTFINDING = 0, EVOPT_STEPBEGIN = 1, EVOPT_STEPEND = 2 };
enum { ERK_DOP853 = 17, ERK_DOP54 = 18, ERK_VERNER98R = 19, ERK_VERNER65E = 20 };

const double FLINT_EPS = std::numeric_limits<double>::epsilon();  /* src/FLINTUtils.f90:49 */
const int FLINT_ROOTFIND_MAXITR = 50;                               /* src/FLINTUtils.f90:52 */
const double LASTSTEP_EXPFAC = 1.01;                                /* src/ERK.f90:43 */
const int MAXNUMSTEPSZEVENTS = 4;                                   /* src/ERK.f90:48 */
const double DEFAULT_EVENTTOL = 1.0e-6;                             /* src/ERK.f90:50 */
const int STIFFTEST_STEPS = 1000;                                   /* src/ERK.f90:53 */
/* src/StepSz.f90:35-44 */
const double SF = 0.9, SFMIN = 0.333, SFMAX = 6.0, HBYHOPTOLD = 1.0e-4;

const int MAXN = 8, MAXM = 4, MAXS = 21, MAXDEG = 8;

/* ---- arithmetic contract primitives ---- */
template <bool FMA> inline double madd(double a, double b, double c)
{
    if (FMA) return __builtin_fma(a, b, c);
    return a * b + c;
}
inline double dmax(double a, double b) { return (a >= b || b != b) ? a : b; }
inline double dmin(double a, double b) { return (a <= b || b != b) ? a : b; }
inline double sign1(double x) { return std::signbit(x) ? -1.0 : 1.0; }   /* sign(1.0, x) */
inline double qnan() { return std::numeric_limits<double>::quiet_NaN(); }

/* libgcc __powidf2 order (x**k with an integer variable exponent) */
inline double powi(double x, int k)
{
    unsigned n = (k < 0) ? (unsigned)(-k) : (unsigned)k;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
    return (k < 0) ? 1.0 / y : y;
}

template <bool FMA> inline double norm2v(const double* v, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s = madd<FMA>(v[i], v[i], s);
    return std::sqrt(s);
}

struct PowFn {
    int mode;  /* 0 = libm pow, 1 = deterministic contract pow */
    double operator()(double x, double y) const { return mode ? flint_det_pow(x, y) : std::pow(x, y); }
};

/* =====================================================================
 * Differential-equation systems (the reference's test programs define
 * them; src/FLINT_base.f90:220-309 is the abstract interface).
 * ===================================================================== */
enum { SYS_CR3BP_PLANAR = 1, SYS_CR3BP_SPATIAL = 2, SYS_LORENZ = 3, SYS_TWOBODY = 4 };
enum { EVSET_NONE = 0, EVSET_CR3BP_SAMPLE3 = 1, EVSET_XY_CROSSING = 2, EVSET_APSIS = 3,
       EVSET_CR3BP_SAMPLE3_TERM = 4, EVSET_CR3BP_SAMPLE3_RESTART = 5 };

struct DiffEqSys {
    int n = 0, m = 0;
    int system_id = 0, eventset_id = 0;
    double sp[8] = {0};   /* system parameters: mu | sigma,rho,beta | GM */

    /* F: the RHS.  params are the optional Integrate(params=...) array. */
    template <bool FMA> void F(double X, const double* Y, const double* params, int nparams, double* f) const
    {
        (void)X;
        switch (system_id) {
        case SYS_CR3BP_PLANAR: {   /* tests/test_CR3BP.f90:167-186 */
            const double mu = sp[0];
            double v1[2] = { Y[0] + mu, Y[1] };
            double n1 = norm2v<FMA>(v1, 2);
            double r1cube = n1 * n1 * n1;
            double v2[2] = { Y[0] - (1.0 - mu), Y[1] };
            double n2 = norm2v<FMA>(v2, 2);
            double r2cube = n2 * n2 * n2;
            f[0] = Y[3]; f[1] = Y[4]; f[2] = Y[5];
            f[3] = madd<FMA>(2.0, Y[4], Y[0]) - (1.0 - mu) * (Y[0] + mu) / r1cube
                   - mu * (Y[0] - (1.0 - mu)) / r2cube;
            f[4] = madd<FMA>(-2.0, Y[3], Y[1]) - (1.0 - mu) * Y[1] / r1cube
                   - mu * Y[1] / r2cube;
            f[5] = 0.0;
            break; }
        case SYS_CR3BP_SPATIAL: {  /* tests/test_WP.f90:88-109 */
            const double mu = sp[0];
            double v1[3] = { Y[0] + mu, Y[1], Y[2] };
            double n1 = norm2v<FMA>(v1, 3);
            double r1cube = n1 * n1 * n1;
            double v2[3] = { Y[0] - (1.0 - mu), Y[1], Y[2] };
            double n2 = norm2v<FMA>(v2, 3);
            double r2cube = n2 * n2 * n2;
            f[0] = Y[3]; f[1] = Y[4]; f[2] = Y[5];
            f[3] = madd<FMA>(2.0, Y[4], Y[0]) - (1.0 - mu) * (Y[0] + mu) / r1cube
                   - mu * (Y[0] - (1.0 - mu)) / r2cube;
            f[4] = madd<FMA>(-2.0, Y[3], Y[1]) - (1.0 - mu) * Y[1] / r1cube
                   - mu * Y[1] / r2cube;
            f[5] = -(1.0 - mu) * Y[2] / r1cube - mu * Y[2] / r2cube;
            break; }
        case SYS_LORENZ: {         /* tests/test_Lorenz.f90:45-57 (object fields, not params) */
            (void)params; (void)nparams;
            const double sigma = sp[0], rho = sp[1], beta = sp[2];
            f[0] = sigma * (Y[1] - Y[0]);
            f[1] = Y[0] * (rho - Y[2]) - Y[1];
            f[2] = madd<FMA>(Y[0], Y[1], -(beta * Y[2]));
            break; }
        case SYS_TWOBODY: {        /* tests/test_TBP.f90:46-59 */
            const double GM = sp[0];
            double nr = norm2v<FMA>(Y, 3);
            double fac = -GM / (nr * nr * nr);
            f[0] = Y[3]; f[1] = Y[4]; f[2] = Y[5];
            f[3] = fac * Y[0]; f[4] = fac * Y[1]; f[5] = fac * Y[2];
            break; }
        default:
            for (int i = 0; i < n; ++i) f[i] = 0.0;
        }
    }

    /* G: event function (src/FLINT_base.f90:257-309).  Y is inout; loc_event is the
     * 1-based located event id or 0 when "not present". */
    void G(double X, double* Y, const int* EvalEvents, double* Value, int* Direction,
           int loc_event, int* loc_action) const
    {
        switch (eventset_id) {
        case EVSET_CR3BP_SAMPLE3:        /* tests/test_CR3BP.f90:97-164 */
        case EVSET_CR3BP_SAMPLE3_TERM:   /* same detectors, event 2 terminates (test variant) */
        case EVSET_CR3BP_SAMPLE3_RESTART:/* same detectors, event 2 restarts    (test variant) */
            if (EvalEvents[0] == 1) {
                Value[0] = (X > 0.5 && X < 0.5005) ? 1.0 : -1.0;
                Direction[0] = 0;
            }
            if (EvalEvents[1] == 1) {
                Value[1] = (Y[0] < 0.0) ? qnan() : Y[1];
                Direction[1] = 1;
            }
            if (EvalEvents[2] == 1) {
                Value[2] = (Y[0] > 0.0) ? qnan() : Y[1];
                Direction[2] = 1;
            }
            if (loc_event > 0 && loc_action) {
                if (loc_event == 1) *loc_action = EVACT_CONTINUE | EVACT_MASK;
                else if (loc_event == 2) {
                    if (eventset_id == EVSET_CR3BP_SAMPLE3_TERM) *loc_action = EVACT_TERMINATE;
                    else if (eventset_id == EVSET_CR3BP_SAMPLE3_RESTART) *loc_action = EVACT_RESTARTINT | EVACT_MASK;
                    else *loc_action = EVACT_CONTINUE | EVACT_MASK;
                }
                else if (loc_event == 3) {
                    Y[0] = 0.994; Y[1] = 0.0; Y[2] = 0.0; Y[3] = 0.0;
                    Y[4] = -2.00158510637908252240537862224; Y[5] = 0.0;
                    *loc_action = EVACT_CHANGESOLY | EVACT_MASK;
                }
            }
            break;
        case EVSET_XY_CROSSING:     /* README.md:121-149 */
            if (EvalEvents[0] == 1) { Value[0] = Y[1]; Direction[0] = 1; }
            if (EvalEvents[1] == 1) { Value[1] = Y[0]; Direction[1] = -1; }
            if (loc_event > 0 && loc_action) *loc_action = EVACT_CONTINUE | EVACT_MASK;
            break;
        case EVSET_APSIS:           /* tests/test_TBP.f90:72-90 */
            if (EvalEvents[0] == 1) {
                Value[0] = Y[0] * Y[3] + Y[1] * Y[4] + Y[2] * Y[5];   /* dot_product */
                Direction[0] = 0;
            }
            if (loc_event > 0 && loc_action) *loc_action = EVACT_CONTINUE;  /* reference leaves it unset */
            break;
        default: break;
        }
    }
};

/* =====================================================================
 * Tableau access (src/ButcherTableaus.f90 via tools/gen_tableaus.py)
 * ===================================================================== */
struct Tableau {
    int s, sint, p, phat, q, pstar, fsal;
    const double *a, *b, *c, *e, *d; const int* dinz; int drows, dcols;
};
Tableau get_tableau(int method)
{
    Tableau t{};
#define FILL(U) t.s = FLINT_##U##_S; t.sint = FLINT_##U##_SINT; t.p = FLINT_##U##_P; t.phat = FLINT_##U##_PHAT; \
    t.q = FLINT_##U##_Q; t.pstar = FLINT_##U##_PSTAR; t.fsal = FLINT_##U##_FSAL; t.a = FLINT_##U##_A; t.b = FLINT_##U##_B; \
    t.c = FLINT_##U##_C; t.e = FLINT_##U##_E; t.d = FLINT_##U##_D; t.dinz = FLINT_##U##_DINZ; \
    t.drows = FLINT_##U##_D_ROWS; t.dcols = FLINT_##U##_D_COLS;
    switch (method) {
    case ERK_DOP853: FILL(DOP853) break;
    case ERK_DOP54: FILL(DOP54) break;
    case ERK_VERNER98R: FILL(VERNER98R) break;
    case ERK_VERNER65E: FILL(VERNER65E) break;
    default: FILL(DOP853)
    }
#undef FILL
    return t;
}

/* =====================================================================
 * StepSize module: src/StepSz.f90
 * ===================================================================== */
/* src/StepSz.f90:53-117 */
template <bool FMA>
void StepSz0Hairer(double& h, int p, int n, double X0, const double* Y0, const double* F0, const double* Sc,
                   double hSign, double hMax, int& FCalls, const DiffEqSys& sys, const double* params, int nparams,
                   const PowFn& powf)
{
    double tmp[MAXN], F1[MAXN], YhF0[MAXN];
    for (int i = 0; i < n; ++i) tmp[i] = Y0[i] / Sc[i];
    double d0 = norm2v<FMA>(tmp, n);                                  /* :79 */
    for (int i = 0; i < n; ++i) tmp[i] = F0[i] / Sc[i];
    double d1 = norm2v<FMA>(tmp, n);                                  /* :80 */
    double h0;
    if (d0 <= 1.0e-5 || d1 <= 1.0e-5) h0 = 1.0e-6;                    /* :83-87 */
    else h0 = 0.01 * d0 / d1;
    h0 = hSign * dmin(h0, hMax);                                      /* :89 */
    for (int i = 0; i < n; ++i) YhF0[i] = madd<FMA>(h0, F0[i], Y0[i]); /* :92 */
    sys.F<FMA>(X0 + h0, YhF0, params, nparams, F1);                   /* :93 */
    for (int i = 0; i < n; ++i) tmp[i] = (F1[i] - F0[i]) / Sc[i];
    double d2 = norm2v<FMA>(tmp, n) / h0;                             /* :96 */
    double dMax = dmax(d1, std::fabs(d2));                            /* :99 */
    double h1;
    if (dMax <= 1.0e-15) h1 = dmax(1.0e-6, std::fabs(h0) * 1.0e-3);   /* :105-109 */
    else h1 = powf(0.01 / dMax, 1.0 / p);
    h = hSign * dmin(dmin(100.0 * std::fabs(h0), h1), hMax);          /* :112 */
    FCalls = 1;                                                       /* :115 */
}

/* src/StepSz.f90:124-180 */
template <bool FMA>
double StepSzHairer(double h, bool IsLastStepRejected, int q, double e, double* StepSzParams, const PowFn& powf)
{
    double SFl = StepSzParams[0], SFmin = StepSzParams[1], SFmax = StepSzParams[2];
    double beta = StepSzParams[3], beta_mult = StepSzParams[4], hbyhoptOld = StepSzParams[5];
    double hbyhopt = powf(e, madd<FMA>(-beta, beta_mult, 1.0 / (q + 1.0)));   /* :153 */
    if (!IsLastStepRejected) {
        hbyhopt = hbyhopt / powf(hbyhoptOld, beta);                   /* :160 */
        double r = h * dmin(SFmax, dmax(SFmin, SFl * 1.0 / hbyhopt)); /* :164 */
        StepSzParams[5] = dmax(e, hbyhoptOld);                        /* :168 */
        return r;
    }
    return h * dmax(SFmin, SFl * 1.0 / hbyhopt);                      /* :176 */
}

/* =====================================================================
 * Brent root finder: src/FLINTUtils.f90:69-207
 * ===================================================================== */
template <class Func>
void Root(double a, double b, double ya, double yb, double tol, Func f, double& x, double& fx, int& niter, int& error)
{
    double aa, bb, cc = 0, fa, fb, fc, xm, ss, dd = 0, ee = 0, pp, qq, rr, tolx;
    const double NEARZERO = 1.0e-20;
    aa = a; bb = b; fa = ya; fb = yb;
    error = 0; niter = 0;
    x = bb;   /* the reference leaves x undefined on the error paths; b is our defined stand-in */
    if (std::fabs(fa) <= NEARZERO) { x = aa; fx = fa; return; }       /* :106-109 */
    else if (std::fabs(fb) <= NEARZERO) { x = bb; fx = fb; return; }  /* :110-114 */
    if ((fa > 0.0 && fb > 0.0) || (fa < 0.0 && fb < 0.0)) { error = 1; niter = 0; return; }  /* :117-121 */
    fc = fb;
    bool done = false;
    int itr = 0;
    while (!done && itr < FLINT_ROOTFIND_MAXITR) {                    /* :127 */
        if ((fc > 0.0 && fb > 0.0) || (fc < 0.0 && fb < 0.0)) {       /* :130-135 */
            cc = aa; fc = fa; dd = bb - aa; ee = dd;
        }
        if (std::fabs(fc) < std::fabs(fb)) {                          /* :137-144 */
            aa = bb; bb = cc; cc = aa; fa = fb; fb = fc; fc = fa;
        }
        tolx = 2.0 * FLINT_EPS * std::fabs(bb) + 0.5 * tol;           /* :146 (2*eps*|b| is exact) */
        xm = 0.5 * (cc - bb);                                         /* :148 */
        if (std::fabs(xm) <= tolx || std::fabs(fa) < NEARZERO) {      /* :150-154 */
            x = bb; done = true; fx = f(x);
        } else {
            if (std::fabs(ee) >= tolx && std::fabs(fa) > std::fabs(fb)) {   /* :156 */
                ss = fb / fa;
                if (std::fabs(aa - cc) < NEARZERO) {                  /* :158-161 */
                    pp = 2.0 * xm * ss; qq = 1.0 - ss;
                } else {                                              /* :163-167 */
                    qq = fa / fc; rr = fb / fc;
                    pp = ss * (2.0 * xm * qq * (qq - rr) - (bb - aa) * (rr - 1.0));
                    qq = (qq - 1.0) * (rr - 1.0) * (ss - 1.0);
                }
                if (pp > NEARZERO) qq = -qq;                          /* :170 */
                pp = std::fabs(pp);
                if ((2.0 * pp) < dmin(3.0 * xm * qq - std::fabs(tolx * qq), std::fabs(ee * qq))) {  /* :172 */
                    ee = dd; dd = pp / qq;
                } else { dd = xm; ee = dd; }
            } else { dd = xm; ee = dd; }                              /* :181-185 */
            aa = bb; fa = fb;                                         /* :187-188 */
            if (std::fabs(dd) > tolx) bb = bb + dd;                   /* :189-190 */
            else { if (xm > 0.0) bb = bb + std::fabs(tolx); else bb = bb - std::fabs(tolx); }
            fb = f(bb);                                               /* :199 */
            itr = itr + 1;
        }
    }
    if (itr >= FLINT_ROOTFIND_MAXITR) error = 2;                      /* :204 */
    niter = itr;
}

/* =====================================================================
 * ERK_class  (src/ERK.f90:72-117, src/FLINT_base.f90:125-211)
 * ===================================================================== */
struct InitOpts {          /* mirrors the optional arguments of Init, src/FLINT_base.f90:318-392 */
    int method = ERK_DOP853;
    int natol = 1, nrtol = 1; double atol[MAXN] = {0}, rtol[MAXN] = {0};
    int has_interp_on = 0, interp_on = 0;
    int n_interp_states = 0; int interp_states[MAXN] = {0};   /* 1-based; 0 = absent */
    int has_min_step = 0; double min_step = 0;
    int has_max_step = 0; double max_step = 0;
    int has_stepsz_params = 0; double stepsz_params[6] = {0};
    int has_events_on = 0, events_on = 0;
    int n_event_stepsz = -1; double event_stepsz[MAXM] = {0};   /* -1 = absent */
    int n_event_options = -1; int event_options[MAXM] = {0};
    int n_event_tol = -1; double event_tol[MAXM] = {0};
};

struct IntegrateOpts {     /* optional arguments of Integrate, src/FLINT_base.f90:400-468 */
    int has_const_step = 0, const_step = 0;
    int has_step_advance = 0, step_advance = 0;
    int has_int_steps_on = 0, int_steps_on = 0;
    int has_xint = 0;                 /* Xint and Yint present */
    int has_event_mask = 0; int event_mask[MAXM] = {0};
    int has_event_states = 0;
    int has_stiff_test = 0;
    int nparams = 0; const double* params = nullptr;
};

template <bool FMA>
struct ERK {
    /* FLINT_class fields, src/FLINT_base.f90:125-211 */
    int status = FLINT_SUCCESS;
    int MaxSteps = 0;
    bool IsScalarTol = true;
    double ATol[MAXN], RTol[MAXN];
    bool InterpOn = false;
    std::vector<int> InterpStates;    /* 1-based */
    double MinStepSize = 0, MaxStepSize = 0;
    double StepSzParams[6];
    bool EventsOn = false;
    double EventStepSz[MAXM]; int EventOptions[MAXM]; double ETol[MAXM];
    int TotalSteps = 0, AcceptedSteps = 0, RejectedSteps = 0, FCalls = 0;
    double X0s = 0, Xf = 0, hf = 0, h0 = 0;
    std::vector<double> Y0s, Yf;
    std::vector<double> Xint;         /* dense storage: Xint(1:acc+1) */
    std::vector<double> Bip;          /* Bip(nI, 0:pstar, step) column-major */
    bool dense_allocated = false;
    /* ERK_class fields */
    int method = ERK_DOP853;
    int p = FLINT_DOP853_P, phat = FLINT_DOP853_PHAT, q = FLINT_DOP853_Q, pstar = FLINT_DOP853_PSTAR;
    int s = 0, sint = 0; bool IsFSALMethod = true;
    double k[MAXS][MAXN];             /* k(:,stage) */
    const DiffEqSys* sys = nullptr;
    Tableau tab;
    PowFn powf{0};

    /* ---------------- Init: src/ERKInit.f90:32-342 ---------------- */
    void Init(const DiffEqSys* DE, int MaxSteps_, const InitOpts& o)
    {
        status = FLINT_SUCCESS;
        method = o.method;
        sys = DE;
        if (MaxSteps_ <= 0) { status = FLINT_ERROR_PARAMS; return; }      /* :68-73 */
        MaxSteps = MaxSteps_;
        if (o.natol == 1 && o.nrtol == 1) {                                /* :76-87 */
            IsScalarTol = true; ATol[0] = o.atol[0]; RTol[0] = o.rtol[0];
        } else if (o.natol == sys->n && o.nrtol == sys->n) {
            IsScalarTol = false;
            for (int i = 0; i < sys->n; ++i) { ATol[i] = o.atol[i]; RTol[i] = o.rtol[i]; }
        } else { status = FLINT_ERROR_PARAMS; return; }
        MinStepSize = o.has_min_step ? std::fabs(o.min_step) : FLINT_EPS;  /* :92-96 */
        MaxStepSize = o.has_max_step ? std::fabs(o.max_step) : 0.0;        /* :100-104 */
        if (o.has_stepsz_params) {                                         /* :108-120 */
            for (int i = 0; i < 6; ++i) StepSzParams[i] = o.stepsz_params[i];
        } else {
            double beta = 0.0, bm = 0.0;
            if (method == ERK_DOP54) { beta = FLINT_DOP54_BETA; bm = FLINT_DOP54_BETA_MULT; }
            else if (method == ERK_DOP853) { beta = FLINT_DOP853_BETA; bm = FLINT_DOP853_BETA_MULT; }
            StepSzParams[0] = SF; StepSzParams[1] = SFMIN; StepSzParams[2] = SFMAX;
            StepSzParams[3] = beta; StepSzParams[4] = bm; StepSzParams[5] = HBYHOPTOLD;
        }
        Xint.clear(); Bip.clear(); dense_allocated = false;               /* :123 */
        if (o.has_interp_on) InterpOn = o.interp_on != 0;                  /* :124-130 */
        EventsOn = false;                                                  /* :135-136 */
        if (o.has_events_on) EventsOn = o.events_on != 0;
        if (EventsOn) {
            if (sys->eventset_id == EVSET_NONE || sys->m <= 0) status = FLINT_ERROR_EVENTPARAMS;  /* :141 */
            InterpStates.clear();
            for (int i = 1; i <= sys->n; ++i) InterpStates.push_back(i);   /* :144 */
            if (o.n_event_stepsz >= 0) {                                   /* :147-158 */
                if (o.n_event_stepsz != sys->m) { status = FLINT_ERROR_DIMENSION; return; }
                for (int i = 0; i < sys->m; ++i) EventStepSz[i] = std::fabs(o.event_stepsz[i]);
            } else for (int i = 0; i < sys->m; ++i) EventStepSz[i] = 0.0;
            if (o.n_event_options >= 0) {                                  /* :161-176 */
                if (o.n_event_options != sys->m) { status = FLINT_ERROR_DIMENSION; return; }
                for (int i = 0; i < sys->m; ++i)
                    if (o.event_options[i] < EVOPT_ROOTFINDING || o.event_options[i] > EVOPT_STEPEND) {
                        status = FLINT_ERROR_EVENTPARAMS; return;
                    }
                for (int i = 0; i < sys->m; ++i) EventOptions[i] = o.event_options[i];
            } else for (int i = 0; i < sys->m; ++i) EventOptions[i] = EVOPT_ROOTFINDING;
            if (o.n_event_tol >= 0) {                                      /* :179-188 */
                if (o.n_event_tol != sys->m) { status = FLINT_ERROR_DIMENSION; return; }
                for (int i = 0; i < sys->m; ++i) ETol[i] = std::fabs(o.event_tol[i]);
            } else for (int i = 0; i < sys->m; ++i) ETol[i] = DEFAULT_EVENTTOL;
        }
        tab = get_tableau(method);                                         /* :198-274 */
        s = tab.s; sint = tab.sint; p = tab.p; phat = tab.phat; q = tab.q; IsFSALMethod = tab.fsal != 0;
        if (InterpOn || EventsOn) pstar = tab.pstar;
        if (InterpOn) {                                                    /* :291-335 */
            if (o.n_interp_states > 0 && !EventsOn) {
                bool bad = o.n_interp_states < 1 || o.n_interp_states > sys->n;
                for (int i = 0; i < o.n_interp_states && !bad; ++i)
                    if (o.interp_states[i] < 1 || o.interp_states[i] > sys->n) bad = true;
                if (bad) status = FLINT_ERROR_INTPSTATES;
                else { InterpStates.assign(o.interp_states, o.interp_states + o.n_interp_states); status = FLINT_SUCCESS; }
            } else {
                InterpStates.clear();
                for (int i = 1; i <= sys->n; ++i) InterpStates.push_back(i);
                status = FLINT_SUCCESS;
            }
            if (status != FLINT_SUCCESS) return;
            IntpMemAllocate();
        }
    }
    void IntpMemAllocate()
    {
        /* the reference allocates MaxSteps+1 entries up front (src/ERKInit.f90:321-332);
         * the oracle grows on demand -- same observable contents. */
        Xint.assign(1, 0.0); Bip.clear(); dense_allocated = true;
    }

    /* ---------------- InterpY: src/ERKInterp.f90:124-140 ---------------- */
    static void InterpY(int np, double X, double X0, double h, int pstar_, const double* bCoeffs, double* Yip)
    {
        double theta = (X - X0) / h;
        for (int i = 0; i < np; ++i) Yip[i] = 0.0;
        for (int kk = 0; kk <= pstar_; ++kk) {
            double tk = powi(theta, kk);
            for (int i = 0; i < np; ++i) Yip[i] = madd<FMA>(bCoeffs[i + np * kk], tk, Yip[i]);
        }
    }

    /* ---------------- generic step: src/ERKStepInt.f90:120-247 ---------------- */
    void stepint_generic(double X0, const double* Y0, double* F0, double h, double* Y1, double* Ysint_pre, double* Ysint,
                         int& FCalls_, bool EstimateErr, double& Err, const double* params, int np)
    {
        const int n = sys->n;
        const double *a = tab.a, *b = tab.b, *c = tab.c, *e = tab.e;
        double Yint[MAXN], aijkj[MAXN];
        for (int i = 0; i < n; ++i) k[0][i] = F0[i];                        /* :149 */
        for (int i = 2; i <= sint - 1; ++i) {                              /* :153-176 */
            int astart = (int)((i - 1) / 2.0 * (i - 2));
            for (int cmp = 0; cmp < n; ++cmp) aijkj[cmp] = 0.0;
            for (int j = 1; j <= i - 1; ++j)
                for (int cmp = 0; cmp < n; ++cmp) aijkj[cmp] = madd<FMA>(k[j - 1][cmp], a[astart + j - 1], aijkj[cmp]);
            for (int cmp = 0; cmp < n; ++cmp) Ysint_pre[cmp] = madd<FMA>(h, aijkj[cmp], Y0[cmp]);
            sys->F<FMA>(madd<FMA>(h, c[i - 2], X0), Ysint_pre, params, np, k[i - 1]);
        }
        {                                                                  /* :182-192 */
            int astart = (int)((sint - 1) / 2.0 * (sint - 2));
            for (int cmp = 0; cmp < n; ++cmp) aijkj[cmp] = 0.0;
            for (int j = 1; j <= sint - 1; ++j)
                for (int cmp = 0; cmp < n; ++cmp) aijkj[cmp] = madd<FMA>(k[j - 1][cmp], a[astart + j - 1], aijkj[cmp]);
            for (int cmp = 0; cmp < n; ++cmp) Yint[cmp] = madd<FMA>(h, aijkj[cmp], Y0[cmp]);
            sys->F<FMA>(madd<FMA>(h, c[sint - 2], X0), Yint, params, np, k[sint - 1]);
        }
        if (!IsFSALMethod) {                                               /* :198 matmul(k, b) */
            for (int cmp = 0; cmp < n; ++cmp) {
                double acc = 0.0;
                for (int j = 0; j < sint; ++j) acc = madd<FMA>(k[j][cmp], b[j], acc);
                Yint[cmp] = madd<FMA>(h, acc, Y0[cmp]);
            }
        }
        if (EstimateErr) {                                                 /* :200-217 */
            double Sc[MAXN], temp[MAXN];
            for (int cmp = 0; cmp < n; ++cmp) {
                double at = IsScalarTol ? ATol[0] : ATol[cmp], rt = IsScalarTol ? RTol[0] : RTol[cmp];
                Sc[cmp] = madd<FMA>(rt, dmax(std::fabs(Y0[cmp]), std::fabs(Yint[cmp])), at);
                temp[cmp] = 0.0;
            }
            for (int j = 0; j < sint; ++j)
                for (int cmp = 0; cmp < n; ++cmp) temp[cmp] = madd<FMA>(k[j][cmp], e[j], temp[cmp]);
            double sm = 0.0;
            for (int cmp = 0; cmp < n; ++cmp) { double qv = h * temp[cmp] / Sc[cmp]; sm = madd<FMA>(qv, qv, sm); }
            Err = std::sqrt(sm / n);
        } else Err = 0.0;
        if (Err <= 1.0) {                                                  /* :224-246 */
            for (int cmp = 0; cmp < n; ++cmp) Y1[cmp] = Yint[cmp];
            if (!IsFSALMethod) {
                sys->F<FMA>(X0 + h, Y1, params, np, F0);
                for (int cmp = 0; cmp < n; ++cmp) Ysint[cmp] = Yint[cmp];
                FCalls_ = sint;
            } else {
                for (int cmp = 0; cmp < n; ++cmp) F0[cmp] = k[sint - 1][cmp];
                /* quirk Q3: the reference never assigns Ysint on this path; the oracle
                 * leaves the caller's previous contents in place (zero-initialised). */
                FCalls_ = sint - 1;
            }
        } else {
            for (int cmp = 0; cmp < n; ++cmp) Y1[cmp] = Y0[cmp];
            FCalls_ = IsFSALMethod ? sint - 2 : sint - 1;
        }
    }

    /* ---------------- generic dense coefficients: src/ERKStepInt.f90:250-323 ---------------- */
    void IntpCoeff_generic(double X0, const double* Y0, double h, double* IpCoeff, int& FCalls_, const double* params, int np)
    {
        const int n = sys->n, nI = (int)InterpStates.size();
        const double *a = tab.a, *c = tab.c, *d = tab.d;
        double aijkj[MAXN];
        for (int i = sint + 1; i <= s; ++i) {                              /* :273-294 */
            int astart = (int)((i - 1) / 2.0 * (i - 2));
            for (int cmp = 0; cmp < n; ++cmp) aijkj[cmp] = 0.0;
            for (int j = 1; j <= i - 1; ++j)
                for (int cmp = 0; cmp < n; ++cmp) aijkj[cmp] = madd<FMA>(k[j - 1][cmp], a[astart + j - 1], aijkj[cmp]);
            for (int cmp = 0; cmp < n; ++cmp) aijkj[cmp] = madd<FMA>(h, aijkj[cmp], Y0[cmp]);
            sys->F<FMA>(madd<FMA>(h, c[i - 2], X0), aijkj, params, np, k[i - 1]);
        }
        FCalls_ = s - sint;                                                /* :296 */
        for (int ii = 0; ii < nI; ++ii) IpCoeff[ii] = Y0[InterpStates[ii] - 1];   /* :304 */
        for (int j = 1; j <= pstar; ++j) {                                 /* :305-322 */
            for (int ii = 0; ii < nI; ++ii) {
                double cont = 0.0;
                for (int i = 0; i < tab.drows; ++i) {
                    int istage = tab.dinz[i];
                    cont = madd<FMA>(d[i + tab.drows * (j - 1)], k[istage - 1][InterpStates[ii] - 1], cont);
                }
                IpCoeff[ii + nI * j] = h * cont;
            }
        }
    }

    /* ---------------- DOP853 step: src/ERKStepInt.f90:327-470 ---------------- */
    void DOP853_stepint(double X0, const double* Y0, double* F0, double h, double* Y1, double* Ysint_pre, double* Ysint,
                        int& FCalls_, bool EstimateErr, double& Err, const double* params, int np)
    {
        const int n = sys->n;
        const double *A = FLINT_DOP853_A, *C = FLINT_DOP853_C, *E = FLINT_DOP853_E, *BHH = FLINT_DOP853_BHH;
        double Yint[MAXN], aijkj[MAXN], Sc[MAXN];
#define KK(st) k[(st) - 1][cmp]
#define AA(idx) A[(idx) - 1]
#define STAGE(st, DEST, SUMEXPR)                                                             \
        for (int cmp = 0; cmp < n; ++cmp) { double acc; SUMEXPR; aijkj[cmp] = acc; DEST[cmp] = madd<FMA>(h, acc, Y0[cmp]); } \
        sys->F<FMA>(madd<FMA>(h, C[(st) - 2], X0), DEST, params, np, k[(st) - 1]);
#define T0(st, ai) acc = KK(st) * AA(ai)
#define T(st, ai) acc = madd<FMA>(KK(st), AA(ai), acc)
        for (int cmp = 0; cmp < n; ++cmp) k[0][cmp] = F0[cmp];              /* :352 */
        STAGE(2, Yint, T0(1, 1))                                            /* :355-357 */
        STAGE(3, Yint, T0(1, 2); T(2, 3))                                   /* :360-362 */
        STAGE(4, Yint, T0(1, 4); T(3, 6))                                   /* :365-367 */
        STAGE(5, Yint, T0(1, 7); T(3, 9); T(4, 10))                         /* :370-372 */
        STAGE(6, Yint, T0(1, 11); T(4, 14); T(5, 15))                       /* :375-378 */
        STAGE(7, Yint, T0(1, 16); T(4, 19); T(5, 20); T(6, 21))             /* :381-384 */
        STAGE(8, Yint, T0(1, 22); T(4, 25); T(5, 26); T(6, 27); T(7, 28))   /* :387-390 */
        STAGE(9, Yint, T0(1, 29); T(4, 32); T(5, 33); T(6, 34); T(7, 35); T(8, 36))            /* :393-396 */
        STAGE(10, Yint, T0(1, 37); T(4, 40); T(5, 41); T(6, 42); T(7, 43); T(8, 44); T(9, 45)) /* :399-403 */
        STAGE(11, Yint, T0(1, 46); T(4, 49); T(5, 50); T(6, 51); T(7, 52); T(8, 53); T(9, 54); T(10, 55))  /* :406-410 */
        STAGE(12, Ysint_pre, T0(1, 56); T(4, 59); T(5, 60); T(6, 61); T(7, 62); T(8, 63); T(9, 64); T(10, 65); T(11, 66)) /* :413-417 */
        for (int cmp = 0; cmp < n; ++cmp) {                                 /* :420-423 */
            double acc;
            T0(1, 67); T(6, 72); T(7, 73); T(8, 74); T(9, 75); T(10, 76); T(11, 77); T(12, 78);
            aijkj[cmp] = acc; Yint[cmp] = madd<FMA>(h, acc, Y0[cmp]);
        }
        if (EstimateErr) {
            double Es = 0.0, Err3 = 0.0;
            for (int cmp = 0; cmp < n; ++cmp) {                             /* :430-434 */
                double at = IsScalarTol ? ATol[0] : ATol[cmp], rt = IsScalarTol ? RTol[0] : RTol[cmp];
                Sc[cmp] = madd<FMA>(rt, dmax(std::fabs(Y0[cmp]), std::fabs(Yint[cmp])), at);
            }
            for (int cmp = 0; cmp < n; ++cmp) {                             /* :439-440 */
                double t = KK(1) * E[0];
                t = madd<FMA>(KK(6), E[5], t); t = madd<FMA>(KK(7), E[6], t); t = madd<FMA>(KK(8), E[7], t);
                t = madd<FMA>(KK(9), E[8], t); t = madd<FMA>(KK(10), E[9], t); t = madd<FMA>(KK(11), E[10], t);
                t = madd<FMA>(KK(12), E[11], t);
                double qv = t / Sc[cmp];
                Es = madd<FMA>(qv, qv, Es);
            }
            for (int cmp = 0; cmp < n; ++cmp) {                             /* :443-444 */
                double t = madd<FMA>(-BHH[0], KK(1), aijkj[cmp]);
                t = madd<FMA>(-BHH[1], KK(9), t);
                t = madd<FMA>(-BHH[2], KK(12), t);
                double qv = t / Sc[cmp];
                Err3 = madd<FMA>(qv, qv, Err3);
            }
            double DenomErr = madd<FMA>(0.01, Err3, Es);                    /* :446 */
            if (DenomErr == 0.0) DenomErr = 1.0;                            /* :447 */
            Err = std::fabs(h) * Es * std::sqrt(1.0 / (n * DenomErr));      /* :448 */
        } else Err = 0.0;
        if (Err <= 1.0) {                                                   /* :456-464 */
            for (int cmp = 0; cmp < n; ++cmp) Y1[cmp] = Yint[cmp];
            sys->F<FMA>(X0 + h, Yint, params, np, k[12]);
            for (int cmp = 0; cmp < n; ++cmp) { F0[cmp] = k[12][cmp]; Ysint[cmp] = Yint[cmp]; }
            FCalls_ = FLINT_DOP853_SINT - 1;
        } else {                                                            /* :465-469 */
            for (int cmp = 0; cmp < n; ++cmp) Y1[cmp] = Y0[cmp];
            FCalls_ = FLINT_DOP853_SINT - 2;
        }
    }

    /* ---------------- DOP853 dense coefficients: src/ERKStepInt.f90:473-563 ---------------- */
    void DOP853_IntpCoeff(double X0, const double* Y0, double h, const double* Y1, double* IntpCoeff, int& FCalls_,
                          const double* params, int np)
    {
        const int n = sys->n, nI = (int)InterpStates.size();
        const double *A = FLINT_DOP853_A, *C = FLINT_DOP853_C, *D = FLINT_DOP853_D;
        double Yint[MAXN], aijkj[MAXN];
        STAGE(14, Yint, T0(1, 79); T(7, 85); T(8, 86); T(9, 87); T(10, 88); T(11, 89); T(12, 90); T(13, 91))      /* :491-494 */
        STAGE(15, Yint, T0(1, 92); T(6, 97); T(7, 98); T(8, 99); T(11, 102); T(12, 103); T(13, 104); T(14, 105))  /* :497-500 */
        STAGE(16, Yint, T0(1, 106); T(6, 111); T(7, 112); T(8, 113); T(9, 114); T(13, 118); T(14, 119); T(15, 120)) /* :503-506 */
        FCalls_ = 3;
        /* D(row, col) for col = 4..7, rows = stages {1,6..16} -> D[row + 12*(col-4)] */
        static const int dst[12] = { 1, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16 };
        for (int ii = 0; ii < nI; ++ii) {
            const int cmp = InterpStates[ii] - 1;
            double cont0 = Y0[cmp];                                          /* :518 */
            double cont1 = Y1[cmp] - Y0[cmp];                                /* :520 */
            double cont2 = madd<FMA>(h, k[0][cmp], -cont1);                  /* :522 */
            double cont3 = madd<FMA>(-h, k[12][cmp], cont1) - cont2;         /* :524 */
            double cont[4];
            for (int col = 0; col < 4; ++col) {                              /* :526-552 */
                double acc = D[0 + 12 * col] * k[0][cmp];
                for (int r = 1; r < 12; ++r) acc = madd<FMA>(D[r + 12 * col], k[dst[r] - 1][cmp], acc);
                cont[col] = h * acc;
            }
            const double cont4 = cont[0], cont5 = cont[1], cont6 = cont[2], cont7 = cont[3];
            IntpCoeff[ii + nI * 0] = cont0;                                  /* :555-562 */
            IntpCoeff[ii + nI * 1] = cont1 + cont2;
            IntpCoeff[ii + nI * 2] = cont3 + cont4 - cont2;
            IntpCoeff[ii + nI * 3] = cont5 + cont6 - 2 * cont4 - cont3;
            IntpCoeff[ii + nI * 4] = madd<FMA>(-3.0, cont6, cont7) - 2 * cont5 + cont4;
            IntpCoeff[ii + nI * 5] = madd<FMA>(3.0, cont6, -3 * cont7) + cont5;
            IntpCoeff[ii + nI * 6] = madd<FMA>(3.0, cont7, -cont6);
            IntpCoeff[ii + nI * 7] = -cont7;
        }
    }
#undef STAGE
#undef T0
#undef T
#undef KK
#undef AA

    /* ---------------- DOP54 step: src/ERKStepInt.f90:568-656 ---------------- */
    void DOP54_stepint(double X0, const double* Y0, double* F0, double h, double* Y1, double* Ysint_pre, double* Ysint,
                       int& FCalls_, bool EstimateErr, double& Err, const double* params, int np)
    {
        const int n = sys->n;
        const double *A = FLINT_DOP54_A, *C = FLINT_DOP54_C, *E = FLINT_DOP54_E;
        double Yint[MAXN], Sc[MAXN];
        for (int cmp = 0; cmp < n; ++cmp) k[0][cmp] = F0[cmp];              /* :590 */
        /* stage i uses the full row (zeros are not skipped by the reference here: a(5) etc. are non-zero) */
        for (int st = 2; st <= 7; ++st) {                                   /* :593-617 */
            int astart = (st - 1) * (st - 2) / 2;
            double* dest = (st == 6) ? Ysint_pre : Yint;
            for (int cmp = 0; cmp < n; ++cmp) {
                double acc = k[0][cmp] * A[astart];
                for (int j = 2; j <= st - 1; ++j) acc = madd<FMA>(k[j - 1][cmp], A[astart + j - 1], acc);
                dest[cmp] = madd<FMA>(h, acc, Y0[cmp]);
            }
            double xs = (st == 7) ? X0 + h : madd<FMA>(h, C[st - 2], X0);
            sys->F<FMA>(xs, dest, params, np, k[st - 1]);
        }
        if (EstimateErr) {                                                  /* :622-636 */
            for (int cmp = 0; cmp < n; ++cmp) {
                double at = IsScalarTol ? ATol[0] : ATol[cmp], rt = IsScalarTol ? RTol[0] : RTol[cmp];
                Sc[cmp] = madd<FMA>(rt, dmax(std::fabs(Y0[cmp]), std::fabs(Yint[cmp])), at);
            }
            double sm = 0.0;
            for (int cmp = 0; cmp < n; ++cmp) {
                double t = k[0][cmp] * E[0];
                t = madd<FMA>(k[2][cmp], E[2], t); t = madd<FMA>(k[3][cmp], E[3], t); t = madd<FMA>(k[4][cmp], E[4], t);
                t = madd<FMA>(k[5][cmp], E[5], t); t = madd<FMA>(k[6][cmp], E[6], t);
                double qv = (h * t) / Sc[cmp];
                sm = madd<FMA>(qv, qv, sm);
            }
            Err = std::sqrt(sm / n);
        } else Err = 0.0;
        if (Err <= 1.0) {                                                   /* :642-655 */
            for (int cmp = 0; cmp < n; ++cmp) { Y1[cmp] = Yint[cmp]; F0[cmp] = k[6][cmp]; Ysint[cmp] = Yint[cmp]; }
            FCalls_ = FLINT_DOP54_SINT - 1;
        } else {
            for (int cmp = 0; cmp < n; ++cmp) Y1[cmp] = Y0[cmp];
            FCalls_ = FLINT_DOP54_SINT - 2;   /* quirk Q10: 6 evaluations were made */
        }
    }

    /* ---------------- DOP54 dense coefficients: src/ERKStepInt.f90:659-700 ---------------- */
    void DOP54_IntpCoeff(const double* Y0, double h, const double* Y1, double* IntpCoeff, int& FCalls_)
    {
        const int nI = (int)InterpStates.size();
        const double* D = FLINT_DOP54_D;   /* column 1 = rows {1,3,4,5,6,7} */
        static const int dst[6] = { 1, 3, 4, 5, 6, 7 };
        FCalls_ = 0;
        for (int ii = 0; ii < nI; ++ii) {
            const int cmp = InterpStates[ii] - 1;
            double cont0 = Y0[cmp];
            double cont1 = Y1[cmp] - Y0[cmp];
            double cont2 = madd<FMA>(h, k[0][cmp], -cont1);
            double cont3 = madd<FMA>(-h, k[6][cmp], cont1) - cont2;
            double acc = D[0] * k[0][cmp];
            for (int r = 1; r < 6; ++r) acc = madd<FMA>(D[r], k[dst[r] - 1][cmp], acc);
            double cont4 = h * acc;
            IntpCoeff[ii + nI * 0] = cont0;
            IntpCoeff[ii + nI * 1] = cont1 + cont2;
            IntpCoeff[ii + nI * 2] = cont3 + cont4 - cont2;
            IntpCoeff[ii + nI * 3] = -2.0 * cont4 - cont3;
            IntpCoeff[ii + nI * 4] = cont4;
        }
    }

    /* ---------------- dispatch: src/ERKStepInt.f90:28-114 ---------------- */
    void StepInt(double X0, const double* Y0, double* F0, double h, double* Y1, double* Ysint_pre, double* Ysint,
                 int& FCalls_, bool EstimateErr, double& Err, const double* params, int np)
    {
        switch (method) {
        case ERK_DOP853: DOP853_stepint(X0, Y0, F0, h, Y1, Ysint_pre, Ysint, FCalls_, EstimateErr, Err, params, np); break;
        case ERK_DOP54: DOP54_stepint(X0, Y0, F0, h, Y1, Ysint_pre, Ysint, FCalls_, EstimateErr, Err, params, np); break;
        case ERK_VERNER65E: case ERK_VERNER98R:
            stepint_generic(X0, Y0, F0, h, Y1, Ysint_pre, Ysint, FCalls_, EstimateErr, Err, params, np); break;
        default: status = FLINT_ERROR;
        }
    }
    void InterpCoeff(double X0, const double* Y0, double h, const double* Y1, double* Bip_, int& FCalls_, const double* params, int np)
    {
        switch (method) {
        case ERK_DOP853: DOP853_IntpCoeff(X0, Y0, h, Y1, Bip_, FCalls_, params, np); break;
        case ERK_DOP54: DOP54_IntpCoeff(Y0, h, Y1, Bip_, FCalls_); break;
        case ERK_VERNER65E: case ERK_VERNER98R: IntpCoeff_generic(X0, Y0, h, Bip_, FCalls_, params, np); break;
        default: status = FLINT_ERROR;
        }
    }

    /* ---------------- Integrate: src/ERKIntegrate.f90:29-802 ---------------- */
    /* Outputs that are `allocatable, intent(out)` in Fortran are std::vectors here. */
    void Integrate(double X0, const double* Y0, double& Xf_io, double* Yf_out, double& StepSz, const IntegrateOpts& o,
                   std::vector<double>* XintOut, std::vector<double>* YintOut, std::vector<double>* EventStates,
                   int* StiffTest)
    {
        const int n = sys->n, m = sys->m;
        const double* params = o.params; const int np = o.nparams;
        int st = FLINT_SUCCESS;
        std::vector<double> BipLoc((size_t)std::max<size_t>(InterpStates.size(), 1) * (MAXDEG + 1), 0.0);
        double* Bipl = BipLoc.data();
        double SzP[6];
        double h, hSign, hnew = 0, hmax, X, Err = 0, LastStepSzFac;
        double Y[MAXN], Y1[MAXN], F0[MAXN], Sc0[MAXN], Ysint_pre[MAXN] = {0}, Ysint[MAXN] = {0};
        int FCalls_ = 0;
        int StiffnessTest, StiffThreshold, NonStiffThreshold, StiffTestSteps;
        bool IntStepsNeeded = false, LastStepRejected = false, LAST_STEP = false, ConstStepSz, IsProblemStiff;
        bool StepOnce, InterpOnL, EventsOnL, BipComputed = false;
        int EventId;
        double EventX0 = 0, EventX1 = 0;
        double EventNewY[MAXN];
        int EvalEvents[MAXM], StepSzEvents[MAXM], EventDir[MAXM] = {0}, nStepSzEvents[MAXM];
        bool EvMask[MAXM], EventsDetected[MAXM];
        int EventAction = EVACT_CONTINUE;   /* quirk Q5: uninitialised in the reference; defined as CONTINUE here */
        double EX0[MAXM][MAXNUMSTEPSZEVENTS], EX1[MAXM][MAXNUMSTEPSZEVENTS];   /* EX0(StepEvId, EventId) */
        double EV0[MAXM][MAXNUMSTEPSZEVENTS], EV1[MAXM][MAXNUMSTEPSZEVENTS];
        std::vector<double> EventData;
        bool EventDataAllocated = false;

        hmax = MaxStepSize;                                                  /* :81-90 */
        InterpOnL = InterpOn; EventsOnL = EventsOn;
        const int nInterpStates = (int)InterpStates.size();
        for (int i = 0; i < 6; ++i) SzP[i] = StepSzParams[i];
        StiffTestSteps = STIFFTEST_STEPS;
        hSign = sign1(Xf_io - X0);                                           /* :94 */
        ConstStepSz = false;                                                 /* :99-103 */
        if (o.has_const_step) {
            ConstStepSz = o.const_step != 0;
            if (ConstStepSz && ((hSign * StepSz) <= 0)) st = FLINT_ERROR_CONSTSTEPSZ;
        }
        LastStepSzFac = ConstStepSz ? 1.0 : LASTSTEP_EXPFAC;                 /* :106-110 */
        StepOnce = o.has_step_advance ? (o.step_advance != 0) : false;       /* :113-114 */
        if (o.has_int_steps_on) {                                            /* :120-127 */
            if (o.int_steps_on && InterpOnL) st = FLINT_ERROR_PARAMS;
            else if (o.int_steps_on && !InterpOnL) IntStepsNeeded = true;
        }
        if (IntStepsNeeded) {                                                /* :129-140 */
            if (o.has_xint && XintOut && YintOut) { XintOut->clear(); YintOut->clear(); }
            else st = FLINT_ERROR_PARAMS;
        }
        if (o.has_stiff_test && StiffTest) {                                 /* :143-154 */
            StiffnessTest = *StiffTest;
            if (StiffnessTest < 0 || StiffnessTest > 2) st = FLINT_ERROR_PARAMS;
        } else StiffnessTest = 0;
        if (InterpOnL) {                                                     /* :157-172 */
            /* the reference sets me%status and then dereferences the unallocated array;
             * the oracle returns the error instead of crashing */
            if (!dense_allocated) { status = FLINT_ERROR_INIT_REQD; return; }
            Xint.assign(1, 0.0); Bip.clear();
        }
        if (EventsOnL) {                                                     /* :175-193 */
            for (int i = 0; i < m; ++i) EvMask[i] = o.has_event_mask ? (o.event_mask[i] != 0) : true;
            bool anym = false; for (int i = 0; i < m; ++i) anym = anym || EvMask[i];
            if (!anym) EventsOnL = false;
            else if (!o.has_event_states || !EventStates) { status = FLINT_ERROR_EVENTPARAMS; return; }  /* reference: sets me%status, then UB */
            else EventDataAllocated = true;
        }
        if (st != FLINT_SUCCESS) { status = st; return; }                    /* :196-199 */

        X = X0;                                                              /* :204-212 */
        for (int i = 0; i < n; ++i) Y[i] = Y0[i];
        FCalls = 0; TotalSteps = 0; AcceptedSteps = 0; RejectedSteps = 0;
        StiffThreshold = 0; NonStiffThreshold = 0;
        st = FLINT_SUCCESS;
        for (int ev = 0; ev < MAXM; ++ev) for (int j = 0; j < MAXNUMSTEPSZEVENTS; ++j) { EV0[ev][j] = 0; EV1[ev][j] = 0; EX0[ev][j] = 0; EX1[ev][j] = 0; }

        if (EventsOnL) {                                                     /* :216-221 */
            EventX0 = X;
            for (int i = 0; i < n; ++i) EventNewY[i] = Y[i];
            for (int i = 0; i < m; ++i) EvalEvents[i] = 1;
            double val[MAXM] = {0};
            sys->G(EventX0, EventNewY, EvalEvents, val, EventDir, 0, nullptr);
            for (int i = 0; i < m; ++i) EV0[i][0] = val[i];
        }
        sys->F<FMA>(X, Y, params, np, F0);                                   /* :224-225 */
        FCalls = FCalls + 1;
        for (int i = 0; i < n; ++i) {                                        /* :230-234 */
            double at = IsScalarTol ? ATol[0] : ATol[i], rt = IsScalarTol ? RTol[0] : RTol[i];
            Sc0[i] = madd<FMA>(rt, std::fabs(Y[i]), at);
        }
        if (hmax == 0.0) hmax = std::fabs(Xf_io - X);                        /* :237 */
        if (StepSz == 0.0) {                                                 /* :240-245 */
            StepSz0Hairer<FMA>(h, p, n, X, Y, F0, Sc0, hSign, hmax, FCalls_, *sys, params, np, powf);
            FCalls = FCalls + FCalls_;
        } else h = StepSz;
        X0s = X; Y0s.assign(Y, Y + n);                                       /* :248-249 */
        if (IntStepsNeeded) { XintOut->push_back(X); YintOut->insert(YintOut->end(), Y, Y + n); }  /* :254-257 */
        if (InterpOnL) Xint[0] = X;                                          /* :259 */

        while (st == FLINT_SUCCESS && !LAST_STEP && TotalSteps < MaxSteps) { /* :263-264 */
            if (hSign * (madd<FMA>(LastStepSzFac, h, X) - Xf_io) > 0) {      /* :268-271 */
                h = Xf_io - X; LAST_STEP = true;
            }
            StepInt(X, Y, F0, h, Y1, Ysint_pre, Ysint, FCalls_, !ConstStepSz, Err, params, np);   /* :274 */
            TotalSteps = TotalSteps + 1;                                     /* :277-278 */
            FCalls = FCalls + FCalls_;
            if (Err <= 1.0) {                                                /* :283 */
                AcceptedSteps = AcceptedSteps + 1;
                if (StepOnce) LAST_STEP = true;                              /* :289 */
                BipComputed = false;                                         /* :292 */
                if (StiffnessTest == 1 || StiffnessTest == 2) {              /* :299-308, CheckStiffness :767-798 */
                    IsProblemStiff = false;
                    double hLamb = 0.0;
                    if ((AcceptedSteps % StiffTestSteps) == 0 || StiffThreshold > 0) {
                        double dk[MAXN], dy[MAXN];
                        for (int i = 0; i < n; ++i) { dk[i] = k[sint - 1][i] - k[sint - 2][i]; dy[i] = Ysint[i] - Ysint_pre[i]; }
                        double StiffN = norm2v<FMA>(dk, n), StiffD = norm2v<FMA>(dy, n);
                        if (StiffD > 0.0) hLamb = std::fabs(h) * StiffN / StiffD;
                        if (hLamb > 6.1) {
                            NonStiffThreshold = 0; StiffThreshold = StiffThreshold + 1;
                            if (StiffThreshold == 15) IsProblemStiff = true;
                        } else {
                            NonStiffThreshold = NonStiffThreshold + 1;
                            if (NonStiffThreshold == 6) StiffThreshold = 0;
                        }
                    }
                    if (IsProblemStiff) {
                        *StiffTest = -1;
                        if (StiffnessTest == 1) { st = FLINT_STIFF_PROBLEM; LAST_STEP = true; }
                    }
                }

                if (EventsOnL) {                                             /* :313-552 */
                    for (int i = 0; i < m; ++i) { EvalEvents[i] = 1; EventsDetected[i] = false; nStepSzEvents[i] = 1; }
                    EventX0 = X; EventX1 = X + h;
                    for (int ev = 0; ev < m; ++ev) for (int j = 0; j < MAXNUMSTEPSZEVENTS; ++j) { EX0[ev][j] = EventX0; EX1[ev][j] = EventX1; }
                    for (int i = 0; i < n; ++i) EventNewY[i] = Y1[i];        /* :324-325 */
                    {
                        double val[MAXM]; for (int i = 0; i < m; ++i) val[i] = EV1[i][0];
                        sys->G(EventX1, EventNewY, EvalEvents, val, EventDir, 0, nullptr);
                        for (int i = 0; i < m; ++i) EV1[i][0] = val[i];
                    }
                    auto DetectEvent = [&](int id, double val0, double val1) -> bool {   /* :715-741 */
                        bool det = false;
                        if (EvMask[id] && !std::isnan(val0) && !std::isnan(val1)) {
                            switch (EventDir[id]) {
                            case 0: if ((val0 <= 0.0 && val1 >= 0.0) || (val0 >= 0.0 && val1 <= 0.0)) det = true; break;
                            case 1: if (val0 <= 0.0 && val1 >= 0.0) det = true; break;
                            case -1: if (val0 >= 0.0 && val1 <= 0.0) det = true; break;
                            }
                        }
                        return det;
                    };
                    for (EventId = 0; EventId < m; ++EventId)               /* :326-328 */
                        EventsDetected[EventId] = DetectEvent(EventId, EV0[EventId][0], EV1[EventId][0]);

                    bool anyss = false;                                      /* :332-337 */
                    for (int i = 0; i < m; ++i) {
                        StepSzEvents[i] = 1;
                        if (EventStepSz[i] == 0 || EventStepSz[i] > hSign * h) StepSzEvents[i] = 0;
                        anyss = anyss || StepSzEvents[i] == 1;
                    }
                    if (anyss) {
                        if (!BipComputed) {                                  /* :340-344 (quirk Q1: mask not consulted) */
                            InterpCoeff(X, Y, h, Y1, Bipl, FCalls_, params, np);
                            FCalls = FCalls + FCalls_; BipComputed = true;
                        }
                        for (EventId = 0; EventId < m; ++EventId) {          /* :347-411 */
                            if (EvMask[EventId] && StepSzEvents[EventId] == 1) {
                                double EventV0[MAXM], EventV1[MAXM], EventVh[MAXM];
                                for (int i = 0; i < m; ++i) EvalEvents[i] = 0;
                                EvalEvents[EventId] = 1;
                                double Eventh = hSign * EventStepSz[EventId];
                                EventX1 = X + Eventh;
                                for (int i = 0; i < m; ++i) { EventV0[i] = EV0[i][0]; EventV1[i] = EV1[i][0]; }
                                EventVh[EventId] = EV1[EventId][0];
                                EventsDetected[EventId] = false;
                                bool LastStepSzEvent = false;
                                while (EventX1 <= (X + h)) {                 /* :372 (quirk Q7: forward only) */
                                    if (EventX1 != (X + h)) {
                                        InterpY(n, EventX1, X, h, pstar, Bipl, EventNewY);
                                        sys->G(EventX1, EventNewY, EvalEvents, EventV1, EventDir, 0, nullptr);
                                    } else EventV1[EventId] = EventVh[EventId];
                                    bool EvStepSzDetected = DetectEvent(EventId, EventV0[EventId], EventV1[EventId]);
                                    if (EvStepSzDetected) {
                                        EventsDetected[EventId] = true;
                                        int slot = nStepSzEvents[EventId] - 1;
                                        EX0[EventId][slot] = EventX0; EX1[EventId][slot] = EventX1;
                                        EV0[EventId][slot] = EventV0[EventId]; EV1[EventId][slot] = EventV1[EventId];
                                        nStepSzEvents[EventId] = nStepSzEvents[EventId] + 1;
                                        if (nStepSzEvents[EventId] > MAXNUMSTEPSZEVENTS) break;
                                    }
                                    if (LastStepSzEvent) break;
                                    EventX0 = EventX1;
                                    EventX1 = EventX1 + Eventh;
                                    if (EventX1 > X + h) { EventX1 = X + h; LastStepSzEvent = true; }
                                    EventV0[EventId] = EventV1[EventId];
                                }
                                if (nStepSzEvents[EventId] > 1) nStepSzEvents[EventId] = nStepSzEvents[EventId] - 1;
                            }
                        }
                    }

                    bool anydet = false; for (int i = 0; i < m; ++i) anydet = anydet || EventsDetected[i];
                    if (anydet) {                                            /* :416-547 */
                        double EXm[MAXM][MAXNUMSTEPSZEVENTS];
                        double EYm[MAXN];
                        for (int ev = 0; ev < MAXM; ++ev) for (int j = 0; j < MAXNUMSTEPSZEVENTS; ++j) EXm[ev][j] = 0.0;
                        for (EventId = 0; EventId < m; ++EventId) {          /* :426-460 */
                            for (int StepEvId = 0; StepEvId < nStepSzEvents[EventId]; ++StepEvId) {
                                if (!EventsDetected[EventId]) continue;
                                switch (EventOptions[EventId]) {
                                case EVOPT_STEPBEGIN: EXm[EventId][StepEvId] = EX0[EventId][StepEvId]; break;
                                case EVOPT_STEPEND: EXm[EventId][StepEvId] = EX1[EventId][StepEvId]; break;
                                default: {
                                    if (!BipComputed) {
                                        InterpCoeff(X, Y, h, Y1, Bipl, FCalls_, params, np);
                                        FCalls = FCalls + FCalls_; BipComputed = true;
                                    }
                                    auto SingleEvent = [&](double xx) -> double {     /* :746-763 */
                                        double ff[MAXN], eval[MAXM] = {0}; int EvalEvent[MAXM], edir[MAXM];
                                        InterpY(nInterpStates, xx, X, h, pstar, Bipl, ff);
                                        for (int i = 0; i < m; ++i) EvalEvent[i] = 0;
                                        EvalEvent[EventId] = 1;
                                        sys->G(xx, ff, EvalEvent, eval, edir, 0, nullptr);
                                        return eval[EventId];
                                    };
                                    int rootiter = 0, rooterror = 0; double fxr = EV0[EventId][StepEvId];
                                    Root(EX0[EventId][StepEvId], EX1[EventId][StepEvId], EV0[EventId][StepEvId], EV1[EventId][StepEvId],
                                         ETol[EventId], SingleEvent, EXm[EventId][StepEvId], fxr, rootiter, rooterror);
                                    EV0[EventId][StepEvId] = fxr;
                                    if (rooterror != 0 && rooterror != 2) st = FLINT_ERROR_EVENTROOT;   /* :452-455 */
                                    break; }
                                }
                            }
                        }
                        for (;;) {                                           /* :465-546 */
                            bool any2 = false; for (int i = 0; i < m; ++i) any2 = any2 || EventsDetected[i];
                            if (!any2) break;
                            int MinEv = -1;                                  /* minloc(EXm(1,:), EventsDetected) :469 */
                            for (int i = 0; i < m; ++i) if (EventsDetected[i] && (MinEv < 0 || EXm[i][0] < EXm[MinEv][0])) MinEv = i;
                            switch (EventOptions[MinEv]) {                   /* :472-480 */
                            case EVOPT_STEPBEGIN: for (int i = 0; i < n; ++i) EYm[i] = Y[i]; break;
                            case EVOPT_STEPEND: for (int i = 0; i < n; ++i) EYm[i] = Y1[i]; break;
                            default: InterpY(nInterpStates, EXm[MinEv][0], X, h, pstar, Bipl, EYm);
                            }
                            for (int i = 0; i < m; ++i) EvalEvents[i] = 0;   /* :484-488 */
                            EvalEvents[MinEv] = 1;
                            for (int i = 0; i < n; ++i) EventNewY[i] = EYm[i];
                            {
                                double val[MAXM]; for (int i = 0; i < m; ++i) val[i] = EV0[i][0];
                                sys->G(EXm[MinEv][0], EventNewY, EvalEvents, val, EventDir, MinEv + 1, &EventAction);
                                for (int i = 0; i < m; ++i) EV0[i][0] = val[i];
                            }
                            EventData.push_back(EXm[MinEv][0]);              /* :493 */
                            EventData.insert(EventData.end(), EYm, EYm + n);
                            EventData.push_back((double)(MinEv + 1));
                            if (EventAction & EVACT_MASK) EvMask[MinEv] = false;   /* :499-501 */
                            int act = EventAction & 0x7f;                    /* :504-529 */
                            if (act == EVACT_TERMINATE) {
                                h = EXm[MinEv][0] - X; for (int i = 0; i < n; ++i) Y1[i] = EYm[i];
                                LAST_STEP = true; BipComputed = false; st = FLINT_EVENT_TERM; break;
                            } else if (act == EVACT_RESTARTINT || act == EVACT_CHANGESOLY) {
                                h = EXm[MinEv][0] - X; for (int i = 0; i < n; ++i) Y1[i] = EYm[i];
                                LAST_STEP = false; BipComputed = false; break;
                            }
                            if (StepSzEvents[MinEv] == 1) {                  /* :534-544 */
                                nStepSzEvents[MinEv] = nStepSzEvents[MinEv] - 1;
                                if (nStepSzEvents[MinEv] == 0) EventsDetected[MinEv] = false;
                                else for (int j = 0; j < nStepSzEvents[MinEv]; ++j) EXm[MinEv][j] = EXm[MinEv][j + 1];
                            } else EventsDetected[MinEv] = false;
                        }
                    }
                    for (int ev = 0; ev < m; ++ev) for (int j = 0; j < MAXNUMSTEPSZEVENTS; ++j) EV0[ev][j] = EV1[ev][j];   /* :550 */
                }

                if (InterpOnL) {                                             /* :555-566 */
                    if (!BipComputed) {
                        InterpCoeff(X, Y, h, Y1, Bipl, FCalls_, params, np);
                        FCalls = FCalls + FCalls_; BipComputed = true;
                    }
                    Xint.push_back(X + h);
                    Bip.insert(Bip.end(), Bipl, Bipl + (size_t)nInterpStates * (pstar + 1));
                }
                if (EventsOnL) {                                             /* :572-582 (quirks Q4, Q5) */
                    int act = EventAction & 0x7f;
                    if (act == EVACT_RESTARTINT) {
                        sys->F<FMA>(X + h, Y1, params, np, k[0]); FCalls = FCalls + 1;
                    } else if (act == EVACT_CHANGESOLY) {
                        for (int i = 0; i < n; ++i) Y1[i] = EventNewY[i];
                        sys->F<FMA>(X + h, Y1, params, np, k[0]); FCalls = FCalls + 1;
                    }
                }
                if (IntStepsNeeded) { XintOut->push_back(X + h); YintOut->insert(YintOut->end(), Y1, Y1 + n); }   /* :585-588 */
                X = X + h;                                                   /* :591-592 */
                for (int i = 0; i < n; ++i) Y[i] = Y1[i];
                if (!ConstStepSz) {                                          /* :595-605 */
                    hnew = StepSzHairer<FMA>(h, false, q, Err, SzP, powf);
                    if (std::fabs(hnew) > hmax) hnew = hSign * hmax;
                    if (LastStepRejected) hnew = hSign * dmin(std::fabs(hnew), std::fabs(h));
                    LastStepRejected = false;
                } else hnew = h;
            } else {                                                         /* :607-624 */
                if (AcceptedSteps >= 1) RejectedSteps = RejectedSteps + 1;
                hnew = StepSzHairer<FMA>(h, true, q, Err, SzP, powf);
                LastStepRejected = true;
                LAST_STEP = false;
            }
            h = hnew;                                                        /* :627 */
            if (!ConstStepSz && !LAST_STEP && std::fabs(h) * 0.01 <= std::fabs(X) * FLINT_EPS)   /* :632-635 */
                st = FLINT_ERROR_STEPSZ_TOOSMALL;
        }

        if (EventsOnL && EventDataAllocated && EventStates) *EventStates = EventData;   /* :679-681 */
        if (LAST_STEP && st == FLINT_SUCCESS) {}                             /* :683-699 */
        else if (EventsOnL && LAST_STEP && st == FLINT_EVENT_TERM) {}
        else if (TotalSteps >= MaxSteps && st == FLINT_SUCCESS) st = FLINT_ERROR_MAXSTEPS;
        else if (st == FLINT_ERROR_STEPSZ_TOOSMALL) {}
        else if (EventsOnL && st == FLINT_ERROR_EVENTROOT) {}
        else if (st == FLINT_STIFF_PROBLEM) {}
        else st = FLINT_ERROR;
        Xf_io = X;                                                           /* :702-708 */
        for (int i = 0; i < n; ++i) Yf_out[i] = Y[i];
        if (!ConstStepSz) StepSz = h;
        Xf = Xf_io; hf = StepSz; Yf.assign(Y, Y + n); status = st;
    }

    /* ---------------- Interpolate: src/ERKInterp.f90:28-121 ---------------- */
    void Interpolate(const double* Xarr, long nX, double* Yarr, bool has_save, bool SaveInterpCoeffs)
    {
        if (!InterpOn) { status = FLINT_ERROR_INTP_OFF; return; }            /* :45-48 */
        bool DeallocMem = true;
        if (has_save) DeallocMem = !SaveInterpCoeffs;                        /* :52-53 */
        long nn = nX;
        if (nn < 1) return;                                                  /* :57 */
        if ((long)Xint.size() < AcceptedSteps + 1) { status = FLINT_ERROR_INTP_OFF; return; }  /* storage freed earlier */
        double hSign = sign1(Xint[AcceptedSteps] - Xint[0]);                 /* :60 */
        for (long i = 0; i + 1 < nn; ++i)                                    /* :61-68 */
            if ((hSign * (Xarr[i] - Xarr[i + 1])) >= 0.0) { status = FLINT_ERROR_INTP_ARRAY; return; }
        if (hSign * (Xarr[0] - Xint[0]) < 0.0 || hSign * (Xarr[nn - 1] - Xint[AcceptedSteps]) > 0.0) {   /* :72-76 */
            status = FLINT_ERROR_INTP_ARRAY; return;
        }
        const int nI = (int)InterpStates.size();
        long X0start = 1, X0loc = 1;                                         /* 1-based as in the reference; :83 */
        for (long ctr = 0; ctr < nn; ++ctr) {                                /* :85-110 */
            while (X0start <= AcceptedSteps) {
                X0start = X0start + 1;
                if (hSign * (Xint[X0start - 1] - Xarr[ctr]) >= 0.0) { X0loc = X0start - 1; break; }
            }
            double X0v, hv;
            if (X0loc <= AcceptedSteps) { X0v = Xint[X0loc - 1]; hv = Xint[X0loc] - Xint[X0loc - 1]; }
            else { X0v = Xarr[ctr]; hv = 0.0; }
            InterpY(nI, Xarr[ctr], X0v, hv, pstar, &Bip[(size_t)(X0loc - 1) * nI * (pstar + 1)], &Yarr[ctr * nI]);
            X0start = X0loc;
        }
        if (DeallocMem) { Xint.clear(); Bip.clear(); dense_allocated = false; }   /* :114-118 */
    }
};

/* the 17 messages of src/FLINT_base.f90:64-82 */
const char* const kErrorStrings[17] = {
    "Successfully Completed", "Termination Event Occurred", "Problem appears to be stiff", "Unknown error",
    "Incorrect user-provided parameters", "Maximum integration steps reached", "Memory allocation error",
    "Memory deallocation error", "Incorrect dimensions", "Step size reduced to very small value",
    "Incorrect interpolation states", "Incorrect interpolation array provided", "Interpolation is not enabled",
    "Init is required", "Incorrect event-related parameters ", "Event root finding failed",
    "Incorrect step size for the non-adaptive option" };

struct Handle {
    DiffEqSys sys;
    bool fma;
    ERK<false> strict;
    ERK<true> fused;
    InitOpts iopt;
    int max_steps = 0;
};

}  // namespace

/* =====================================================================
 * C API used by tests/ and bench.py through ctypes.
 * ===================================================================== */
extern "C" {

struct oracle_init_t {
    int system_id, eventset_id, n, m;
    double sysparams[8];
    int max_steps, method;
    int natol, nrtol; double atol[8], rtol[8];
    int has_interp_on, interp_on;
    int n_interp_states; int interp_states[8];
    int has_min_step; double min_step;
    int has_max_step; double max_step;
    int has_stepsz_params; double stepsz_params[6];
    int has_events_on, events_on;
    int n_event_stepsz; double event_stepsz[4];
    int n_event_options; int event_options[4];
    int n_event_tol; double event_tol[4];
    int arith_fma;   /* 0 strict, 1 fused contract */
    int pow_mode;    /* 0 libm, 1 deterministic */
};

struct oracle_int_t {
    int has_const_step, const_step, has_step_advance, step_advance, has_int_steps_on, int_steps_on, has_xint;
    int has_event_mask; int event_mask[4];
    int has_event_states, has_stiff_test, stiff_test;
    int nparams; double params[8];
};

static void fill_init(const oracle_init_t* c, Handle* H)
{
    H->sys.n = c->n; H->sys.m = c->m; H->sys.system_id = c->system_id; H->sys.eventset_id = c->eventset_id;
    for (int i = 0; i < 8; ++i) H->sys.sp[i] = c->sysparams[i];
    InitOpts& o = H->iopt; o = InitOpts();
    o.method = c->method; o.natol = c->natol; o.nrtol = c->nrtol;
    for (int i = 0; i < 8; ++i) { o.atol[i] = c->atol[i]; o.rtol[i] = c->rtol[i]; o.interp_states[i] = c->interp_states[i]; }
    o.has_interp_on = c->has_interp_on; o.interp_on = c->interp_on; o.n_interp_states = c->n_interp_states;
    o.has_min_step = c->has_min_step; o.min_step = c->min_step; o.has_max_step = c->has_max_step; o.max_step = c->max_step;
    o.has_stepsz_params = c->has_stepsz_params; for (int i = 0; i < 6; ++i) o.stepsz_params[i] = c->stepsz_params[i];
    o.has_events_on = c->has_events_on; o.events_on = c->events_on;
    o.n_event_stepsz = c->n_event_stepsz; o.n_event_options = c->n_event_options; o.n_event_tol = c->n_event_tol;
    for (int i = 0; i < 4; ++i) { o.event_stepsz[i] = c->event_stepsz[i]; o.event_options[i] = c->event_options[i]; o.event_tol[i] = c->event_tol[i]; }
    H->fma = c->arith_fma != 0; H->max_steps = c->max_steps;
    H->strict.powf.mode = c->pow_mode; H->fused.powf.mode = c->pow_mode;
}

static IntegrateOpts make_iopts(const oracle_int_t* c)
{
    IntegrateOpts o;
    o.has_const_step = c->has_const_step; o.const_step = c->const_step; o.has_step_advance = c->has_step_advance;
    o.step_advance = c->step_advance; o.has_int_steps_on = c->has_int_steps_on; o.int_steps_on = c->int_steps_on;
    o.has_xint = c->has_xint; o.has_event_mask = c->has_event_mask;
    for (int i = 0; i < 4; ++i) o.event_mask[i] = c->event_mask[i];
    o.has_event_states = c->has_event_states; o.has_stiff_test = c->has_stiff_test;
    o.nparams = c->nparams; o.params = c->nparams > 0 ? c->params : nullptr;
    return o;
}

void* oracle_create(const oracle_init_t* c)
{
    Handle* H = new Handle();
    fill_init(c, H);
    if (H->fma) H->fused.Init(&H->sys, c->max_steps, H->iopt); else H->strict.Init(&H->sys, c->max_steps, H->iopt);
    return H;
}
void oracle_destroy(void* h) { delete (Handle*)h; }
int oracle_status(void* h) { Handle* H = (Handle*)h; return H->fma ? H->fused.status : H->strict.status; }
const char* oracle_status_string(int s) { return (s >= 0 && s <= 16) ? kErrorStrings[s] : "Unknown error"; }

/* Scalar Integrate.  Variable-length outputs are copied into caller buffers with capacities
 * (xint_cap points, ev_cap events); *_count receive the true counts. */
int oracle_integrate(void* h, double x0, const double* y0, double* xf, double* yf, double* stepsz, const oracle_int_t* io,
                     double* xint, double* yint, long xint_cap, long* xint_count,
                     double* evstates, long ev_cap, long* ev_count, int* stifftest, int* counters /*[4]*/)
{
    Handle* H = (Handle*)h;
    IntegrateOpts o = make_iopts(io);
    std::vector<double> xi, yi, es;
    int stiff = io->stiff_test;
    const int n = H->sys.n;
    if (H->fma) H->fused.Integrate(x0, y0, *xf, yf, *stepsz, o, &xi, &yi, &es, &stiff);
    else H->strict.Integrate(x0, y0, *xf, yf, *stepsz, o, &xi, &yi, &es, &stiff);
    if (stifftest) *stifftest = stiff;
    if (xint_count) *xint_count = (long)xi.size();
    if (xint && yint) for (long i = 0; i < (long)xi.size() && i < xint_cap; ++i) { xint[i] = xi[i]; for (int c = 0; c < n; ++c) yint[i * n + c] = yi[i * n + c]; }
    long nev = (long)es.size() / (n + 2);
    if (ev_count) *ev_count = nev;
    if (evstates) for (long i = 0; i < nev && i < ev_cap; ++i) for (int c = 0; c < n + 2; ++c) evstates[i * (n + 2) + c] = es[i * (n + 2) + c];
    if (counters) {
        if (H->fma) { counters[0] = H->fused.TotalSteps; counters[1] = H->fused.AcceptedSteps; counters[2] = H->fused.RejectedSteps; counters[3] = H->fused.FCalls; }
        else { counters[0] = H->strict.TotalSteps; counters[1] = H->strict.AcceptedSteps; counters[2] = H->strict.RejectedSteps; counters[3] = H->strict.FCalls; }
    }
    return oracle_status(h);
}

int oracle_interpolate(void* h, const double* xarr, long nx, double* yarr, int has_save, int save)
{
    Handle* H = (Handle*)h;
    if (H->fma) H->fused.Interpolate(xarr, nx, yarr, has_save != 0, save != 0);
    else H->strict.Interpolate(xarr, nx, yarr, has_save != 0, save != 0);
    return oracle_status(h);
}

/* dense storage inspection (for parity tests of the stored coefficients) */
long oracle_dense_steps(void* h) { Handle* H = (Handle*)h; return H->fma ? (long)H->fused.Xint.size() - 1 : (long)H->strict.Xint.size() - 1; }
void oracle_dense_copy(void* h, double* xint, double* bip)
{
    Handle* H = (Handle*)h;
    const std::vector<double>& X = H->fma ? H->fused.Xint : H->strict.Xint;
    const std::vector<double>& B = H->fma ? H->fused.Bip : H->strict.Bip;
    if (xint) std::copy(X.begin(), X.end(), xint);
    if (bip) std::copy(B.begin(), B.end(), bip);
}

/* Ensemble driver: N independent Integrate calls, OpenMP over trajectories with one
 * ERK object per thread (objects hold mutable state).
 * Layouts are SoA like the product's batch ABI: y0[c*N + i], counters[j*N + i],
 * events[(c*ev_cap + e)*N + i].  x0 scalar, xf per trajectory in/out.
 * Optional fused interpolation onto a shared grid: yarr[(i*G + g)*nI + c]. */
int oracle_integrate_ensemble(const oracle_init_t* c, const oracle_int_t* io, long N, double x0, const double* y0,
                              double* xf, double* yf, double* stepsz, int* counters, int* status,
                              int ev_cap, int* ev_count, double* events, int* stifftest,
                              const double* xarr, long G, double* yarr, int nthreads)
{
    const int n = c->n;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    int used = 1;
#pragma omp parallel
    {
#ifdef _OPENMP
#pragma omp single
        used = omp_get_num_threads();
#endif
        Handle* H = new Handle();
        fill_init(c, H);
        if (H->fma) H->fused.Init(&H->sys, c->max_steps, H->iopt); else H->strict.Init(&H->sys, c->max_steps, H->iopt);
        IntegrateOpts o = make_iopts(io);
        std::vector<double> xi, yi, es, ybuf;
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < N; ++i) {
            double y0l[MAXN], yfl[MAXN];
            for (int k = 0; k < n; ++k) y0l[k] = y0[(long)k * N + i];
            double xfl = xf[i], hl = stepsz ? stepsz[i] : 0.0;
            int stiff = stifftest ? stifftest[i] : io->stiff_test;
            int st;
            int cnt[4];
            if (H->fma) {
                H->fused.Integrate(x0, y0l, xfl, yfl, hl, o, &xi, &yi, &es, &stiff); st = H->fused.status;
                cnt[0] = H->fused.TotalSteps; cnt[1] = H->fused.AcceptedSteps; cnt[2] = H->fused.RejectedSteps; cnt[3] = H->fused.FCalls;
            } else {
                H->strict.Integrate(x0, y0l, xfl, yfl, hl, o, &xi, &yi, &es, &stiff); st = H->strict.status;
                cnt[0] = H->strict.TotalSteps; cnt[1] = H->strict.AcceptedSteps; cnt[2] = H->strict.RejectedSteps; cnt[3] = H->strict.FCalls;
            }
            xf[i] = xfl; if (stepsz) stepsz[i] = hl; if (stifftest) stifftest[i] = stiff;
            for (int k = 0; k < n; ++k) yf[(long)k * N + i] = yfl[k];
            if (counters) for (int j = 0; j < 4; ++j) counters[(long)j * N + i] = cnt[j];
            if (status) status[i] = st;
            if (ev_count) {
                long nev = (long)es.size() / (n + 2);
                ev_count[i] = (int)nev;
                if (events) for (long e = 0; e < nev && e < ev_cap; ++e)
                    for (int k = 0; k < n + 2; ++k) events[((long)k * ev_cap + e) * N + i] = es[e * (n + 2) + k];
            }
            if (xarr && yarr && G > 0) {
                const int nI = H->fma ? (int)H->fused.InterpStates.size() : (int)H->strict.InterpStates.size();
                if (H->fma) H->fused.Interpolate(xarr, G, &yarr[(size_t)i * G * nI], true, true);
                else H->strict.Interpolate(xarr, G, &yarr[(size_t)i * G * nI], true, true);
            }
        }
        delete H;
    }
    return used;
}

/* direct access to the building blocks, for unit-level parity tests */
void oracle_rhs(int system_id, const double* sysparams, int n, int fma, double x, const double* y, double* f)
{
    DiffEqSys s; s.n = n; s.system_id = system_id; for (int i = 0; i < 8; ++i) s.sp[i] = sysparams[i];
    if (fma) s.F<true>(x, y, nullptr, 0, f); else s.F<false>(x, y, nullptr, 0, f);
}
double oracle_pow(double x, double y, int mode) { PowFn p{mode}; return p(x, y); }
void oracle_pow_vec(const double* x, const double* y, double* out, long n, int mode) { PowFn p{mode}; for (long i = 0; i < n; ++i) out[i] = p(x[i], y[i]); }
int oracle_brent(double a, double b, double tol, const double* poly, int deg, double* x, double* fx, int* niter)
{
    /* root of a polynomial (Horner) -- exercises Root on a known function */
    auto f = [&](double t) { double v = poly[deg]; for (int i = deg - 1; i >= 0; --i) v = v * t + poly[i]; return v; };
    int err = 0;
    Root(a, b, f(a), f(b), tol, f, *x, *fx, *niter, err);
    return err;
}
int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  /* extern "C" */
