/* runtime_systems.cuh -- user DiffEqSys functors registered at RUN time from this text (flint_sys_register_source, csrc/rtc.cu):
 * the stock library compiles them with NVRTC; nothing is rebuilt.
 *
 *  RtLorenz      the Lorenz system against the user-facing concept: must equal the built-in functor, and therefore the CPU
 *                oracle, bit for bit (tests/test_gpu_round2.py::test_runtime_registered_systems)
 *  RtVanDerPol   x'' - mu (1 - x^2) x' + x = 0 with one event (x crosses zero upwards, CONTINUE); mu can also arrive through
 *                Integrate(params=...) (load_params); checked against scipy's DOP853 */
#include "systems.cuh"

namespace flint {

struct RtLorenz : UserSystem<3, 0> {
    static constexpr int ID = FLINT_SYS_USER_BASE + 10;
    static constexpr bool INLINE_ACCEL = true;
    double sigma, rho, beta;
    __device__ __forceinline__ void load(const double* sp) { sigma = sp[0]; rho = sp[1]; beta = sp[2]; }
    template <bool FMA>
    __device__ __forceinline__ void accel(const double (&y)[3], double (&f)[3]) const
    {
        f[0] = sigma * (y[1] - y[0]);
        f[1] = __dsub_rn(__dmul_rn(y[0], rho - y[2]), y[1]);
        f[2] = madd<FMA>(y[0], y[1], -(beta * y[2]));
    }
};

struct RtVanDerPol : UserSystem<2, 1> {
    static constexpr int ID = FLINT_SYS_USER_BASE + 11;
    double mu;
    __device__ __forceinline__ void load(const double* sp) { mu = sp[0]; }
    __device__ __forceinline__ void load_params(const double* params, int n_params) { if (n_params > 0) mu = params[0]; }
    template <bool FMA>
    __device__ __forceinline__ void accel(const double (&y)[2], double (&f)[2]) const
    {
        f[0] = y[1];
        f[1] = madd<FMA>(mu * (1.0 - y[0] * y[0]), y[1], -y[0]);
    }
    __device__ __forceinline__ void events(int /*evset*/, double /*x*/, double (&y)[2], unsigned eval_bits, double (&value)[1], int (&dir)[1],
                                           int loc_event, int& action) const
    {
        if (eval_bits & 1u) { value[0] = y[0]; dir[0] = 1; }
        if (loc_event > 0) action = FLINT_EVENTACTION_CONTINUE;
    }
};

}  // namespace flint
