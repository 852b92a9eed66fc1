/* example_systems.cuh -- user DiffEqSys functors registered at build time:
 *     python -m flint_b200.build --tag _user --only NONE --user-systems tests/user_systems/example_systems.cuh
 * builds flint_b200/libflint_b200_user.so with these two systems (and none of the built-in ones).
 *
 *  UserLorenz      the Lorenz system written against the user-facing concept: must reproduce the built-in functor, and
 *                  therefore the CPU oracle, bit for bit (tests/test_gpu_parity.py::test_user_registered_systems)
 *  UserVanDerPol   x'' - mu (1 - x^2) x' + x = 0 with one event (x crosses zero upwards, CONTINUE): no counterpart in the
 *                  reference; checked against scipy's DOP853 */
#include "systems.cuh"

namespace flint {

struct UserLorenz : UserSystem<3, 0> {
    static constexpr int ID = FLINT_SYS_USER_BASE + 0;
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
FLINT_REGISTER_SYSTEM(UserLorenz, false)

struct UserVanDerPol : UserSystem<2, 1> {
    static constexpr int ID = FLINT_SYS_USER_BASE + 1;
    double mu;
    __device__ __forceinline__ void load(const double* sp) { mu = sp[0]; }
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
FLINT_REGISTER_SYSTEM(UserVanDerPol, true)

}  // namespace flint
