/* systems.cuh -- DiffEqSys functors compiled into the library.
 *
 * The reference's user code extends `DiffEqSys` with a type-bound RHS `F` and an
 * event-function pointer `G` (src/FLINT_base.f90:220-309).  On the GPU those
 * become __device__ functors selected by id through flint_sys_create(); the RHS
 * and event definitions below are the ones of the reference's own test programs
 * (the workloads of BASELINE.json).  Expression order follows the Fortran so the
 * result is bit-identical to the oracle (arith.cuh).
 *
 * A system provides:
 *   static constexpr int N       state dimension
 *   static constexpr int M_MAX   largest number of events among its event sets
 *   F, split into its trivial and its computed part so that the integrator neither marshals nor
 *   multiplies what it already knows:
 *     NIN, in_idx(i)             the states the computed part reads
 *     NOUT                       number of computed derivative components
 *     fsrc(c)                    f[c] = y[fsrc(c)] if >= 0; +0.0 if -1; out[-2 - fsrc(c)] otherwise
 *     accel<FMA>(in, out)        the computed components (branch-free division / square root, arith.cuh)
 *     INLINE_ACCEL               accel is a handful of operations: inline it at every stage instead of one out-of-line copy
 *     zc(c)                      component c is identically +0.0 in state and derivative (kernel variant for
 *                                trajectories that live in an invariant coordinate plane; false in the general systems)
 *     ZPLANE                     bit mask of the components a plane variant of this system may drop (0: none)
 *   rhs<FMA>(x, y, f)            F assembled from the above (eval_rhs, below)
 *   events(evset, x, y, eval_bits, value, dir, loc_event, action)   G
 * All four systems are autonomous (F does not read x).
 * `params` of Integrate(params=...) is forwarded in sp_params (unused by the
 * reference's own systems).
 */
#ifndef FLINT_B200_SYSTEMS_CUH
#define FLINT_B200_SYSTEMS_CUH

#include "arith.cuh"
#include "../../include/flint_b200.h"

namespace flint {

/* F assembled from a system's trivial/computed split */
template <bool FMA, class SYS>
__device__ __forceinline__ void eval_rhs(const SYS& sys, const double (&y)[SYS::N], double (&f)[SYS::N])
{
    double in[SYS::NIN], out[SYS::NOUT];
#pragma unroll
    for (int i = 0; i < SYS::NIN; ++i) in[i] = y[SYS::in_idx(i)];
    sys.template accel<FMA>(in, out);
    sys.template assemble<FMA>(y, out, f);
}

/* the four CR3BP quotients with the plain operators (cold path of accel) */
struct Quot4 { double q[4]; };
static __device__ __noinline__ Quot4 cr3bp_quotients_plain(double s1, double s2, double u1, double u2, double u3, double u4)
{
    const double n1 = sqrt(s1), n2 = sqrt(s2);
    const double r1c = (n1 * n1) * n1, r2c = (n2 * n2) * n2;
    Quot4 r;
    r.q[0] = u1 / r1c; r.q[1] = u2 / r2c; r.q[2] = u3 / r1c; r.q[3] = u4 / r2c;
    return r;
}

static __device__ __noinline__ double twobody_fac_plain(double GM, double s)
{
    const double nr = sqrt(s);
    return -GM / ((nr * nr) * nr);
}

/* ---- planar CR3BP: tests/test_CR3BP.f90:167-186; events :97-164 and README.md:121-149 ---- */
struct SysCR3BPPlanar {
    static constexpr int N = 6;
    static constexpr int M_MAX = 3;
    static constexpr int ID = FLINT_SYS_CR3BP_PLANAR;
    double mu, omm;   /* omm = 1 - mu (one rounding, as `1.0_WP-mu` at run time; formed on the host, capi.cu) */
    bool pok;         /* both parameters in [2^-60, 2^60): part of the fast-path range test of accel */
    __device__ __forceinline__ void load(const double* sp) { mu = sp[0]; omm = sp[7]; pok = small60(mu) & small60(omm); }

    /* out of line: the four quotients (they need only the position); inline (assemble): the Coriolis terms and the
     * subtractions -- passing the velocities through the call made the callee spill and reload them (ncu: 2.7 % of the
     * samples on that reload) */
    static constexpr int NIN = 2, NOUT = 4;
    static constexpr bool INLINE_ACCEL = false;
    /* general (6-component) kernels, 2^19 orbits, strict, ms with parked events / in-loop events / no events:
     * out of line 128 x 4 with barrier 85.1 / 114.6 / 75.7; inlined in the hot loop 128 x 4: 85.4 / 129.2 / 74.7,
     * 256 x 2: 81.4 / 125.5 / 70.5, 512 x 1 without barrier: 80.9 / 115.8 / 67.6 */
    static constexpr bool INLINE_ACCEL_HOT = true;
    static constexpr int STEP_BLOCK = 512, STEP_MIN_BLOCKS = 1, STEP_BARRIER = 0;   /* a barrier every 4 / 16 / 64 iterations instead of
                                                                                       none: 131.0 / 133.9 / 133.5 vs 128.5 ms (plane variant);
                                                                                       384 threads at 168 registers / 448 at 146: 125.8 / 132.0 vs 121.4 ms */
    static constexpr unsigned ZPLANE = (1u << 2) | (1u << 5);   /* z = vz = 0 is invariant under F and under every event action below */
    __host__ __device__ static constexpr bool zc(int) { return false; }
    __host__ __device__ static constexpr int in_idx(int i) { return i; }                          /* y0 y1 */
    __host__ __device__ static constexpr int fsrc(int c) { return c < 3 ? c + 3 : (c == 5 ? -1 : -2 - (c - 3)); }

    static constexpr bool LAZY_RANGE = true;
    mutable bool bad = false;                                     /* accel_lazy: an operand left the fast-path range during this attempt */
    template <bool FMA> __device__ __forceinline__ void accel(const double (&in)[NIN], double (&out)[NOUT]) const { accel_impl<FMA, false>(in, out); }
    template <bool FMA> __device__ __forceinline__ void accel_lazy(const double (&in)[NIN], double (&out)[NOUT]) const { accel_impl<FMA, true>(in, out); }
    template <bool FMA, bool LAZY>
    __device__ __forceinline__ void accel_impl(const double (&in)[NIN], double (&out)[NOUT]) const
    {
        const double y0 = in[0], y1 = in[1];
        const double a = y0 + mu;
        const double b = y0 - omm;
        double s1 = __dmul_rn(a, a)   /* == 0 + x*x: x*x is never -0 */;
        s1 = madd<FMA>(y1, y1, s1);
        double s2 = __dmul_rn(b, b);
        s2 = madd<FMA>(y1, y1, s2);
        const double u1 = omm * a, u2 = mu * b, u3 = omm * y1, u4 = mu * y1;
        /* fast-path range: s1, s2 in [2^-200, 2^200) and the four numerators in [2^-300, 2^300) -- implied by a, b in
         * [2^-100, 2^99) (each carries its sum of squares), y1 in [2^-240, 2^99) and the two parameters in [2^-60, 2^60):
         * 3 tests per call instead of 6 (arith.cuh: small100).  The floor of the other coordinates is as low as the numerators
         * allow: measured, a 2^-100 floor on y and z sent the halo-orbit ensembles of test_WP to the out-of-line operators often
         * enough to cost 50 % (12.0 -> 18.3 ms per 2^18 orbits); with 2^-240 they never leave the fast path */
        const bool ok = small100(a) & small100(b) & small240(y1) & pok;
        double n1, n2, i1, i2;
        sqrt2_nochk(s1, s2, n1, n2);
        const double r1c = (n1 * n1) * n1, r2c = (n2 * n2) * n2;
        rcp2_newton(r1c, r2c, i1, i2);
        double q1 = div_rcp_nochk(u1, r1c, i1), q2 = div_rcp_nochk(u2, r2c, i2);
        double q3 = div_rcp_nochk(u3, r1c, i1), q4 = div_rcp_nochk(u4, r2c, i2);
        if (LAZY) bad = bad | !ok;                   /* the lane loop repeats the attempt through accel */
        else if (!ok) {                              /* operands outside the fast-path range: plain operators, out of line */
            const Quot4 q = cr3bp_quotients_plain(s1, s2, u1, u2, u3, u4);
            q1 = q.q[0]; q2 = q.q[1]; q3 = q.q[2]; q4 = q.q[3];
        }
        out[0] = q1; out[1] = q2; out[2] = q3; out[3] = q4;
    }
    template <bool FMA>
    __device__ __forceinline__ void assemble(const double (&y)[N], const double (&q)[NOUT], double (&f)[N]) const
    {
        f[0] = y[3]; f[1] = y[4]; f[2] = y[5];
        f[3] = (madd<FMA>(2.0, y[4], y[0]) - q[0]) - q[1];
        f[4] = (madd<FMA>(-2.0, y[3], y[1]) - q[2]) - q[3];
        f[5] = 0.0;
    }
    template <bool FMA>
    __device__ __forceinline__ void rhs(double /*x*/, const double (&y)[N], double (&f)[N]) const { eval_rhs<FMA>(*this, y, f); }

    /* eval_bits: bit i set <=> EvalEvents(i+1) == 1.  loc_event: 1-based located id, 0 = absent. */
    __device__ __forceinline__ void events(int evset, double x, double (&y)[N], unsigned eval_bits,
                                           double (&value)[M_MAX], int (&dir)[M_MAX], int loc_event, int& action) const
    {
        if (evset == FLINT_EVSET_XY_CROSSING) {
            if (eval_bits & 1u) { value[0] = y[1]; dir[0] = 1; }
            if (eval_bits & 2u) { value[1] = y[0]; dir[1] = -1; }
            if (loc_event > 0) action = FLINT_EVENTACTION_CONTINUE | FLINT_EVENTACTION_MASK;
            return;
        }
        /* SAMPLE3 family */
        if (eval_bits & 1u) { value[0] = (x > 0.5 && x < 0.5005) ? 1.0 : -1.0; dir[0] = 0; }
        if (eval_bits & 2u) { value[1] = (y[0] < 0.0) ? qnan() : y[1]; dir[1] = 1; }
        if (eval_bits & 4u) { value[2] = (y[0] > 0.0) ? qnan() : y[1]; dir[2] = 1; }
        if (loc_event > 0) {
            if (loc_event == 1) action = FLINT_EVENTACTION_CONTINUE | FLINT_EVENTACTION_MASK;
            else if (loc_event == 2) {
                if (evset == FLINT_EVSET_CR3BP_SAMPLE3_TERM) action = FLINT_EVENTACTION_TERMINATE;
                else if (evset == FLINT_EVSET_CR3BP_SAMPLE3_RESTART) action = FLINT_EVENTACTION_RESTARTINT | FLINT_EVENTACTION_MASK;
                else action = FLINT_EVENTACTION_CONTINUE | FLINT_EVENTACTION_MASK;
            } else if (loc_event == 3) {
                y[0] = 0.994; y[1] = 0.0; y[2] = 0.0; y[3] = 0.0;
                y[4] = -2.00158510637908252240537862224; y[5] = 0.0;
                action = FLINT_EVENTACTION_CHANGESOLY | FLINT_EVENTACTION_MASK;
            }
        }
    }
};

/* Planar CR3BP for trajectories whose z and vz are exactly +0.0: both stay +0.0 through every stage
 * ((+0) + h*(+-0) = +0, f[2] = y[5], f[5] = 0), contribute (+-0/Sc)^2 = +0 to the error norms, and no event
 * action of this system moves them, so the integrator may drop the two components without changing a bit
 * of any result.  erk_init_kernel checks the condition per trajectory; the generic kernel runs otherwise. */
struct SysCR3BPPlanarZ : SysCR3BPPlanar {
    __host__ __device__ static constexpr bool zc(int c) { return c == 2 || c == 5; }
    /* launch configuration inherited (right-hand side inlined in the hot loop, 512 x 1, no barrier).  Inlined, the attempt is
     * 2525 instructions instead of 3020, but the hot loop is 42 KB against a 32 KB instruction cache: one CTA of 16 warps per
     * SM (they start together and stay roughly in step, so a line is fetched once for all of them) and no barrier (exact lock
     * step starves the fp64 pipe).  Measured, 2^20 orbits, strict / events: 128 x 4 with barrier 141.6 ms, 256 x 2 with 131.7,
     * 512 x 1 with 139.1, 512 x 1 without 131.3 (out of line, 128 x 4 with barrier: 143.3); without events 131.5 / 126.9 /
     * 119.1 / 119.3 (130.3). */
};

/* ---- spatial CR3BP: tests/test_WP.f90:88-109 (no events in the reference) ---- */
struct SysCR3BPSpatial {
    static constexpr int N = 6;
    static constexpr int M_MAX = 1;
    static constexpr int ID = FLINT_SYS_CR3BP_SPATIAL;
    double mu, omm;
    bool pok;
    __device__ __forceinline__ void load(const double* sp) { mu = sp[0]; omm = sp[7]; pok = small60(mu) & small60(omm); }

    static constexpr int NIN = 3, NOUT = 6;                       /* position in, the six quotients out (see the planar system) */
    static constexpr bool INLINE_ACCEL = false;
    /* halo orbit, 2^18 trajectories, DOP853 / Verner98R ms: out of line 128 x 4: 13.8 / 19.3, 256 x 2: 13.3 / 18.5; inlined in the
     * hot loop 128 x 4: 13.1 / 19.3, 256 x 2: 12.0 / 18.0, 512 x 1 without barrier: 13.0 / 18.7 */
    static constexpr bool INLINE_ACCEL_HOT = true;
    static constexpr int STEP_BLOCK = 256, STEP_MIN_BLOCKS = 2;
    static constexpr unsigned ZPLANE = 0;
    __host__ __device__ static constexpr bool zc(int) { return false; }
    __host__ __device__ static constexpr int in_idx(int i) { return i; }                           /* y0..y2 */
    __host__ __device__ static constexpr int fsrc(int c) { return c < 3 ? c + 3 : -2 - (c - 3); }

    static constexpr bool LAZY_RANGE = true;
    mutable bool bad = false;
    template <bool FMA> __device__ __forceinline__ void accel(const double (&in)[NIN], double (&out)[NOUT]) const { accel_impl<FMA, false>(in, out); }
    template <bool FMA> __device__ __forceinline__ void accel_lazy(const double (&in)[NIN], double (&out)[NOUT]) const { accel_impl<FMA, true>(in, out); }
    template <bool FMA, bool LAZY>
    __device__ __forceinline__ void accel_impl(const double (&in)[NIN], double (&out)[NOUT]) const
    {
        const double y0 = in[0], y1 = in[1], y2 = in[2];
        const double a = y0 + mu;
        const double b = y0 - omm;
        double s1 = __dmul_rn(a, a)   /* == 0 + x*x: x*x is never -0 */;
        s1 = madd<FMA>(y1, y1, s1);
        s1 = madd<FMA>(y2, y2, s1);
        double s2 = __dmul_rn(b, b);
        s2 = madd<FMA>(y1, y1, s2);
        s2 = madd<FMA>(y2, y2, s2);
        const double u1 = omm * a, u2 = mu * b, u3 = omm * y1, u4 = mu * y1, u5 = -omm * y2, u6 = mu * y2;
        const bool ok = small100(a) & small100(b) & small240(y1) & small240(y2) & pok;     /* see the planar system */
        double n1, n2, i1, i2;
        sqrt2_nochk(s1, s2, n1, n2);
        const double r1c = (n1 * n1) * n1, r2c = (n2 * n2) * n2;
        rcp2_newton(r1c, r2c, i1, i2);
        double q1 = div_rcp_nochk(u1, r1c, i1), q2 = div_rcp_nochk(u2, r2c, i2);
        double q3 = div_rcp_nochk(u3, r1c, i1), q4 = div_rcp_nochk(u4, r2c, i2);
        double q5 = div_rcp_nochk(u5, r1c, i1), q6 = div_rcp_nochk(u6, r2c, i2);
        if (LAZY) bad = bad | !ok;
        else if (!ok) {                              /* operands outside the fast-path range: plain operators, out of line */
            const Quot4 q = cr3bp_quotients_plain(s1, s2, u1, u2, u3, u4);
            const Quot4 z = cr3bp_quotients_plain(s1, s2, u5, u6, u5, u6);
            q1 = q.q[0]; q2 = q.q[1]; q3 = q.q[2]; q4 = q.q[3]; q5 = z.q[0]; q6 = z.q[1];
        }
        out[0] = q1; out[1] = q2; out[2] = q3; out[3] = q4; out[4] = q5; out[5] = q6;
    }
    template <bool FMA>
    __device__ __forceinline__ void assemble(const double (&y)[N], const double (&q)[NOUT], double (&f)[N]) const
    {
        f[0] = y[3]; f[1] = y[4]; f[2] = y[5];
        f[3] = (madd<FMA>(2.0, y[4], y[0]) - q[0]) - q[1];
        f[4] = (madd<FMA>(-2.0, y[3], y[1]) - q[2]) - q[3];
        f[5] = q[4] - q[5];
    }
    template <bool FMA>
    __device__ __forceinline__ void rhs(double, const double (&y)[N], double (&f)[N]) const { eval_rhs<FMA>(*this, y, f); }
    __device__ __forceinline__ void events(int, double, double (&)[N], unsigned, double (&)[M_MAX], int (&)[M_MAX], int, int&) const {}
};

/* ---- Lorenz: tests/test_Lorenz.f90:45-57 ---- */
struct SysLorenz {
    static constexpr int N = 3;
    static constexpr int M_MAX = 1;
    static constexpr int ID = FLINT_SYS_LORENZ;
    double sigma, rho, beta;
    __device__ __forceinline__ void load(const double* sp) { sigma = sp[0]; rho = sp[1]; beta = sp[2]; }

    static constexpr int NIN = 3, NOUT = 3;
    static constexpr bool INLINE_ACCEL = true;                    /* 9 flops */
    static constexpr int STEP_BLOCK = 256, STEP_MIN_BLOCKS = 2;   /* DOP54, 2^18 trajectories: 83.8 ms at 128 x 4, 79.5 ms at 256 x 2; 256 x 3 (80
                                                                     registers) 88.8, 256 x 4 (64 registers) 89.2: the lane state spills */
    static constexpr unsigned ZPLANE = 0;
    __host__ __device__ static constexpr bool zc(int) { return false; }
    __host__ __device__ static constexpr int in_idx(int i) { return i; }
    __host__ __device__ static constexpr int fsrc(int c) { return -2 - c; }

    template <bool FMA>
    __device__ __forceinline__ void accel(const double (&y)[NIN], double (&f)[NOUT]) const
    {
        f[0] = sigma * (y[1] - y[0]);
        f[1] = __dsub_rn(__dmul_rn(y[0], rho - y[2]), y[1]);   /* never fused, in both modes */
        f[2] = madd<FMA>(y[0], y[1], -(beta * y[2]));
    }
    template <bool FMA>
    __device__ __forceinline__ void assemble(const double (&y)[N], const double (&o)[NOUT], double (&f)[N]) const { scatter_rhs<meta::remove_cvref_t<decltype(*this)>>(y, o, f); }
    template <bool FMA>
    __device__ __forceinline__ void rhs(double, const double (&y)[N], double (&f)[N]) const { eval_rhs<FMA>(*this, y, f); }
    __device__ __forceinline__ void events(int, double, double (&)[N], unsigned, double (&)[M_MAX], int (&)[M_MAX], int, int&) const {}
};

/* ---- two-body: tests/test_TBP.f90:46-59; apsis event :72-90 ---- */
struct SysTwoBody {
    static constexpr int N = 6;
    static constexpr int M_MAX = 1;
    static constexpr int ID = FLINT_SYS_TWOBODY;
    double GM;
    __device__ __forceinline__ void load(const double* sp) { GM = sp[0]; }

    static constexpr int NIN = 3, NOUT = 3;
    static constexpr bool INLINE_ACCEL = false;
    static constexpr bool INLINE_ACCEL_HOT = true;                /* DOP853, 2^17 orbits x 10 periods: 12.8 -> 11.2 ms (128 x 4 with barrier;
                                                                     256 x 2: 12.5, 512 x 1 without barrier: 11.3) */
    static constexpr unsigned ZPLANE = 0;
    __host__ __device__ static constexpr bool zc(int) { return false; }
    __host__ __device__ static constexpr int in_idx(int i) { return i; }                           /* position */
    __host__ __device__ static constexpr int fsrc(int c) { return c < 3 ? c + 3 : -2 - (c - 3); }

    static constexpr bool LAZY_RANGE = true;                      /* see the planar CR3BP system */
    mutable bool bad = false;
    template <bool FMA> __device__ __forceinline__ void accel(const double (&y)[NIN], double (&f)[NOUT]) const { accel_impl<FMA, false>(y, f); }
    template <bool FMA> __device__ __forceinline__ void accel_lazy(const double (&y)[NIN], double (&f)[NOUT]) const { accel_impl<FMA, true>(y, f); }
    template <bool FMA, bool LAZY>
    __device__ __forceinline__ void accel_impl(const double (&y)[NIN], double (&f)[NOUT]) const
    {
        double s = __dmul_rn(y[0], y[0])   /* == 0 + x*x: x*x is never -0 */;
        s = madd<FMA>(y[1], y[1], s);
        s = madd<FMA>(y[2], y[2], s);
        const bool ok = moderate_sq(s) & moderate(GM);
        const double nr = sqrt_nochk(s);
        const double d = (nr * nr) * nr;
        double fac = div_rcp_nochk(-GM, d, rcp_newton(d));
        if (LAZY) bad = bad | !ok;
        else if (!ok) fac = twobody_fac_plain(GM, s);
        f[0] = fac * y[0]; f[1] = fac * y[1]; f[2] = fac * y[2];
    }
    template <bool FMA>
    __device__ __forceinline__ void assemble(const double (&y)[N], const double (&o)[NOUT], double (&f)[N]) const { scatter_rhs<meta::remove_cvref_t<decltype(*this)>>(y, o, f); }
    template <bool FMA>
    __device__ __forceinline__ void rhs(double, const double (&y)[N], double (&f)[N]) const { eval_rhs<FMA>(*this, y, f); }
    __device__ __forceinline__ void events(int, double, double (&y)[N], unsigned eval_bits, double (&value)[M_MAX],
                                           int (&dir)[M_MAX], int loc_event, int& action) const
    {
        if (eval_bits & 1u) {   /* dot_product(Y(1:3), Y(4:6)): left-to-right, never fused (as the oracle) */
            value[0] = __dadd_rn(__dadd_rn(__dmul_rn(y[0], y[3]), __dmul_rn(y[1], y[4])), __dmul_rn(y[2], y[5]));
            dir[0] = 0;
        }
        if (loc_event > 0) action = FLINT_EVENTACTION_CONTINUE;
    }
};

/* ---- user systems ------------------------------------------------------------------------------
 * A user right-hand side / event function is a struct like the ones above in a header of its own, registered at build
 * time:  python -m flint_b200.build --user-systems my_systems.cuh   (flint_b200/build.py; tests/user_systems/ has an
 * example).  Deriving from UserSystem<n, m_max> supplies the bookkeeping of a system whose n derivative components
 * are all computed from all n states; the struct then provides
 *     static constexpr int ID = FLINT_SYS_USER_BASE + k;                       the id flint_sys_create() selects
 *     void load(const double* sp)                                              system parameters (<= 7 doubles)
 *     template <bool FMA> void accel(const double (&y)[N], double (&f)[N]) const   F  (use madd<FMA> for a*b+c)
 *     void events(evset, x, y, eval_bits, value, dir, loc_event, action) const      G  (optional)
 * and the line  FLINT_REGISTER_SYSTEM(Name, has_events)  which the build script reads.  Optional tuning members (defaults in
 * parentheses; measured per built-in system above):
 *     static constexpr bool INLINE_ACCEL      (false)  inline accel everywhere: a right-hand side of a handful of operations
 *     static constexpr bool INLINE_ACCEL_HOT  (absent = false)  inline it in the hot step kernel only
 *     static constexpr int  STEP_BLOCK, STEP_MIN_BLOCKS  (128, 4)  CTA size and resident CTAs of the step kernel
 *     static constexpr int  STEP_BARRIER      (absent: 1, or 0 without events when INLINE_ACCEL)  CTA barrier every K-th attempt */
#define FLINT_REGISTER_SYSTEM(NAME, HAS_EVENTS)
template <int N_, int M_MAX_>
struct UserSystem {
    static constexpr int N = N_, M_MAX = (M_MAX_ > 0 ? M_MAX_ : 1), NIN = N_, NOUT = N_;
    static constexpr bool INLINE_ACCEL = false;                   /* a user system may override: true for a very small F */
    static constexpr unsigned ZPLANE = 0;
    __host__ __device__ static constexpr bool zc(int) { return false; }
    __host__ __device__ static constexpr int in_idx(int i) { return i; }
    __host__ __device__ static constexpr int fsrc(int c) { return -2 - c; }
    __device__ __forceinline__ void events(int, double, double (&)[N_], unsigned, double (&)[M_MAX], int (&)[M_MAX], int, int&) const {}
    template <bool FMA>
    __device__ __forceinline__ void assemble(const double (&)[N_], const double (&o)[N_], double (&f)[N_]) const
    {
#pragma unroll
        for (int c = 0; c < N_; ++c) f[c] = o[c];
    }
};

}  // namespace flint
#endif
