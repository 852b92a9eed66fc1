/* cr3bp_abi.c -- the call sequence of the reference's tests/test_CR3BP.f90 (Init :390-392, Integrate :399-404, Info)
 * through the C ABI, from plain C: what fortran/flint_gpu.f90 binds with ISO_C_BINDING.
 *
 *   gcc -O2 -I include examples/cr3bp_abi.c -o cr3bp_abi -L flint_b200 -lflint_b200 -Wl,-rpath,$PWD/flint_b200 -lm
 *
 * Prints one line "key value" per result.  Without a CUDA device Init still works (host-side validation, as the
 * reference) and Integrate reports FLINT_ERROR with the reason in flint_last_error(): there is no CPU fallback. */
#include <math.h>
#include <stdio.h>
#include "flint_b200.h"

int main(void)
{
    const double mu = 0.012277471;                                   /* tests/test_CR3BP.f90:33 */
    const double y0[6] = {0.994, 0.0, 0.0, 0.0, -2.00158510637908252240537862224, 0.0};
    const double period = 17.0652165601579625588917206249;
    const double rtol = 1.0e-11, atol = 1.0e-14;                     /* :248 */
    flint_sys_t* de = flint_sys_create(FLINT_SYS_CR3BP_PLANAR, FLINT_EVSET_CR3BP_SAMPLE3, 6, 3, &mu, 1);
    flint_erk_t* erk = flint_erk_create();
    if (!de || !erk) { printf("create failed\n"); return 2; }

    const double ev_stepsz[3] = {1.0e-5, 0.0, 0.0}, ev_tol[3] = {1.0e-3, 1.0e-9, 1.0e-9};   /* :380-386 */
    flint_init_opts io = {0};
    io.present = FLINT_INIT_METHOD | FLINT_INIT_INTERP_ON | FLINT_INIT_EVENTS_ON | FLINT_INIT_EVENT_STEPSZ | FLINT_INIT_EVENT_TOL;
    io.method = ERK_DOP853;
    io.n_atol = 1; io.atol = &atol; io.n_rtol = 1; io.rtol = &rtol;
    io.interp_on = 0; io.events_on = 1;
    io.n_event_stepsz = 3; io.event_stepsz = ev_stepsz; io.n_event_tol = 3; io.event_tol = ev_tol;
    int st = flint_erk_init(erk, de, 100000, &io);
    printf("init_status %d %s\n", st, flint_status_string(st));
    if (st != FLINT_SUCCESS) return 1;

    flint_int_opts o = {0};
    o.present = FLINT_INT_CONST_STEPSZ | FLINT_INT_EVENT_MASK | FLINT_INT_EVENT_STATES;
    o.use_const_stepsz = 0;
    o.event_mask[0] = 0; o.event_mask[1] = 1; o.event_mask[2] = 0;   /* EvMask = [.FALSE., .TRUE., .FALSE.] */
    double xf = 2.5 * period, yf[6], stepsz = 0.0, evrec[4 * 8];
    int nev = 0, stiff = 1;
    flint_events_out ev = {4, evrec, &nev};
    st = flint_erk_integrate(erk, 0.0, y0, &xf, yf, &stepsz, &o, NULL, &ev, &stiff);
    printf("integrate_status %d %s\n", st, flint_status_string(st));
    if (st == FLINT_ERROR) { printf("detail %s\n", flint_last_error()); goto done; }

    int last, nsteps, nacc, nrej, nfc, ready;
    flint_erk_info(erk, &last, &nsteps, &nacc, &nrej, &nfc, &ready, NULL, NULL, NULL, NULL, NULL);
    printf("xf %.17g\n", xf);
    printf("steps %d accepted %d rejected %d fcalls %d\n", nsteps, nacc, nrej, nfc);
    printf("events %d\n", nev);
    if (nev > 0) printf("event1 x %.9f y1 %.9f id %d\n", evrec[0], evrec[1], (int)evrec[7]);   /* EventStates(:,1) = [X, Y(1:6), id] */
    {   /* Jacobi constant drift, tests/test_CR3BP.f90:197-209 */
        double c[2];
        const double* yy[2] = {y0, yf};
        for (int k = 0; k < 2; ++k) {
            const double* y = yy[k];
            const double r1 = sqrt((y[0] + mu) * (y[0] + mu) + y[1] * y[1] + y[2] * y[2]);
            const double r2 = sqrt((y[0] - 1.0 + mu) * (y[0] - 1.0 + mu) + y[1] * y[1] + y[2] * y[2]);
            c[k] = y[0] * y[0] + y[1] * y[1] + 2.0 * (1.0 - mu) / r1 + 2.0 * mu / r2 - (y[3] * y[3] + y[4] * y[4] + y[5] * y[5]);
        }
        printf("jacobi_drift %.3e\n", fabs(c[1] - c[0]));
    }
done:
    flint_erk_destroy(erk);
    flint_sys_destroy(de);
    return 0;
}
