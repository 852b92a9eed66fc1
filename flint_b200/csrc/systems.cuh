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
 *   rhs<FMA>(x, y, f)            F
 *   events(evset, x, y, eval_bits, value, dir, loc_event, action)   G
 * `params` of Integrate(params=...) is forwarded in sp_params (unused by the
 * reference's own systems).
 */
#ifndef FLINT_B200_SYSTEMS_CUH
#define FLINT_B200_SYSTEMS_CUH

#include "arith.cuh"
#include "../../include/flint_b200.h"

namespace flint {

/* ---- planar CR3BP: tests/test_CR3BP.f90:167-186; events :97-164 and README.md:121-149 ---- */
struct SysCR3BPPlanar {
    static constexpr int N = 6;
    static constexpr int M_MAX = 3;
    static constexpr int ID = FLINT_SYS_CR3BP_PLANAR;
    double mu, omm;   /* omm = 1 - mu (one rounding, as `1.0_WP-mu` at run time) */
    __device__ __forceinline__ void load(const double* sp) { mu = sp[0]; omm = 1.0 - mu; }

    template <bool FMA>
    __device__ __forceinline__ void rhs(double /*x*/, const double (&y)[N], double (&f)[N]) const
    {
        const double a = y[0] + mu;
        const double b = y[0] - omm;
        double s1 = __dmul_rn(a, a)   /* == 0 + x*x: x*x is never -0 */;
        s1 = madd<FMA>(y[1], y[1], s1);
        const double n1 = sqrt(s1);
        const double r1c = (n1 * n1) * n1;
        double s2 = __dmul_rn(b, b)   /* == 0 + x*x: x*x is never -0 */;
        s2 = madd<FMA>(y[1], y[1], s2);
        const double n2 = sqrt(s2);
        const double r2c = (n2 * n2) * n2;
        f[0] = y[3]; f[1] = y[4]; f[2] = y[5];
        f[3] = (madd<FMA>(2.0, y[4], y[0]) - (omm * a) / r1c) - (mu * b) / r2c;
        f[4] = (madd<FMA>(-2.0, y[3], y[1]) - (omm * y[1]) / r1c) - (mu * y[1]) / r2c;
        f[5] = 0.0;
    }

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

/* ---- spatial CR3BP: tests/test_WP.f90:88-109 (no events in the reference) ---- */
struct SysCR3BPSpatial {
    static constexpr int N = 6;
    static constexpr int M_MAX = 1;
    static constexpr int ID = FLINT_SYS_CR3BP_SPATIAL;
    double mu, omm;
    __device__ __forceinline__ void load(const double* sp) { mu = sp[0]; omm = 1.0 - mu; }

    template <bool FMA>
    __device__ __forceinline__ void rhs(double, const double (&y)[N], double (&f)[N]) const
    {
        const double a = y[0] + mu;
        const double b = y[0] - omm;
        double s1 = __dmul_rn(a, a)   /* == 0 + x*x: x*x is never -0 */;
        s1 = madd<FMA>(y[1], y[1], s1);
        s1 = madd<FMA>(y[2], y[2], s1);
        const double n1 = sqrt(s1);
        const double r1c = (n1 * n1) * n1;
        double s2 = __dmul_rn(b, b)   /* == 0 + x*x: x*x is never -0 */;
        s2 = madd<FMA>(y[1], y[1], s2);
        s2 = madd<FMA>(y[2], y[2], s2);
        const double n2 = sqrt(s2);
        const double r2c = (n2 * n2) * n2;
        f[0] = y[3]; f[1] = y[4]; f[2] = y[5];
        f[3] = (madd<FMA>(2.0, y[4], y[0]) - (omm * a) / r1c) - (mu * b) / r2c;
        f[4] = (madd<FMA>(-2.0, y[3], y[1]) - (omm * y[1]) / r1c) - (mu * y[1]) / r2c;
        f[5] = (-omm * y[2]) / r1c - (mu * y[2]) / r2c;
    }
    __device__ __forceinline__ void events(int, double, double (&)[N], unsigned, double (&)[M_MAX], int (&)[M_MAX], int, int&) const {}
};

/* ---- Lorenz: tests/test_Lorenz.f90:45-57 ---- */
struct SysLorenz {
    static constexpr int N = 3;
    static constexpr int M_MAX = 1;
    static constexpr int ID = FLINT_SYS_LORENZ;
    double sigma, rho, beta;
    __device__ __forceinline__ void load(const double* sp) { sigma = sp[0]; rho = sp[1]; beta = sp[2]; }

    template <bool FMA>
    __device__ __forceinline__ void rhs(double, const double (&y)[N], double (&f)[N]) const
    {
        f[0] = sigma * (y[1] - y[0]);
        f[1] = __dsub_rn(__dmul_rn(y[0], rho - y[2]), y[1]);   /* never fused, in both modes */
        f[2] = madd<FMA>(y[0], y[1], -(beta * y[2]));
    }
    __device__ __forceinline__ void events(int, double, double (&)[N], unsigned, double (&)[M_MAX], int (&)[M_MAX], int, int&) const {}
};

/* ---- two-body: tests/test_TBP.f90:46-59; apsis event :72-90 ---- */
struct SysTwoBody {
    static constexpr int N = 6;
    static constexpr int M_MAX = 1;
    static constexpr int ID = FLINT_SYS_TWOBODY;
    double GM;
    __device__ __forceinline__ void load(const double* sp) { GM = sp[0]; }

    template <bool FMA>
    __device__ __forceinline__ void rhs(double, const double (&y)[N], double (&f)[N]) const
    {
        double s = __dmul_rn(y[0], y[0])   /* == 0 + x*x: x*x is never -0 */;
        s = madd<FMA>(y[1], y[1], s);
        s = madd<FMA>(y[2], y[2], s);
        const double nr = sqrt(s);
        const double fac = -GM / ((nr * nr) * nr);
        f[0] = y[3]; f[1] = y[4]; f[2] = y[5];
        f[3] = fac * y[0]; f[4] = fac * y[1]; f[5] = fac * y[2];
    }
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

}  // namespace flint
#endif
