"""GPU parity tests: the sm_100a kernels, called through the C ABI, against the CPU oracle
on identical seeded inputs.  Bar: BIT-EXACT final states, step sizes, counters, status codes,
event records and interpolated values (the path is deterministic IEEE arithmetic; see
DESIGN.md "Arithmetic contract")."""
import os

import numpy as np
import pytest

import flint_b200 as F
from flint_b200 import problems as P
import cases as K

pytestmark = pytest.mark.gpu

METHODS = [F.ERK_DOP853, F.ERK_DOP54, F.ERK_VERNER98R, F.ERK_VERNER65E]
MNAME = {17: "DOP853", 18: "DOP54", 19: "Verner98R", 20: "Verner65E"}


def test_div_sqrt_helpers_equal_ieee_operators():
    """The right-hand sides and error norms use branch-free division / square root (csrc/arith.cuh)
    that must equal `/` and `sqrt` bit for bit wherever their range test passes: 2^28 operand sets
    (every exponent, NaN/Inf/subnormals, physical magnitudes, zero numerators), two seeds."""
    for seed in (12345, 987654321):
        bad, ndiv, nsqrt = F.selftest_arith(1 << 28, seed)
        assert bad == 0
        assert ndiv > (1 << 27) and nsqrt > (1 << 27)      # the range test is not vacuous


@pytest.mark.parametrize("arith", [F.FLINT_ARITH_STRICT, F.FLINT_ARITH_FMA])
def test_config3_northstar_parity(arith):
    """Config 3a (the metric's configuration) on 4096 seeded orbits."""
    case = K.cr3bp_case(F.ERK_DOP853, 1e-12, arith)
    y0 = P.cr3bp_ensemble(4096)
    g, o = case.run_gpu(y0), case.run_oracle(y0)
    K.compare_batch(g, o, case, "config3 arith=%d" % arith)
    # a few of the 1e-6-perturbed Arenstorf orbits pass so close to the Moon that the controller gives up
    # (FLINT_ERROR_STEPSZ_TOOSMALL, src/ERKIntegrate.f90:632-635) -- in the oracle and on the GPU alike
    ok = g["status"] == 0
    assert ok.mean() > 0.99 and set(np.unique(g["status"])) <= {0, F.FLINT_ERROR_STEPSZ_TOOSMALL}
    assert (g["nEvents"] == 1).all()
    # known answer (SURVEY.md App. C.6 / tests/results_CR3BP.txt:11): event 2 near x = 0.399136, y1 = 0.748351
    assert np.allclose(g["EventStates"][0, 0], 0.399136, atol=1e-3) and np.allclose(g["EventStates"][1, 0], 0.748351, atol=1e-3)
    assert (g["EventStates"][7, 0] == 2.0).all()


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("method", METHODS)
def test_cr3bp_all_methods_events(method, arith):
    case = K.cr3bp_case(method, 1e-11, arith)
    y0 = P.cr3bp_ensemble(384, start=1000)
    K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, MNAME[method])


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("eventset", [F.FLINT_EVSET_XY_CROSSING, F.FLINT_EVSET_CR3BP_SAMPLE3_TERM, F.FLINT_EVSET_CR3BP_SAMPLE3_RESTART])
def test_event_sets_and_actions(eventset, arith):
    """README events (two root-finds per orbit), TERMINATE and RESTARTINT|MASK actions."""
    case = K.cr3bp_case(F.ERK_DOP853, 1e-11, arith, eventset=eventset)
    y0 = P.cr3bp_ensemble(256, start=5000)
    g, o = case.run_gpu(y0), case.run_oracle(y0)
    K.compare_batch(g, o, case, "evset %d" % eventset)
    if eventset == F.FLINT_EVSET_CR3BP_SAMPLE3_TERM:
        assert (g["status"] == F.FLINT_EVENT_TERM).all()
    if eventset == F.FLINT_EVSET_XY_CROSSING:
        assert (g["nEvents"] == 2).all()


@pytest.mark.parametrize("arith", [0, 1])
def test_unmasked_substep_scanning_and_changesoly(arith):
    """All three SAMPLE3 events unmasked: the square-pulse event needs EventStepSz scanning (a10),
    event 3 exercises CHANGESOLY|MASK; STEPBEGIN/STEPEND options on events 1 and 3."""
    case = K.cr3bp_case(F.ERK_DOP853, 1e-10, arith, norb=0.6, event_mask=[1, 1, 1], event_stepsz=[1e-4, 0.0, 0.0],
                        event_options=[F.FLINT_EVENTOPTION_STEPBEGIN, F.FLINT_EVENTOPTION_ROOTFINDING, F.FLINT_EVENTOPTION_STEPEND],
                        max_events=8)
    y0 = P.cr3bp_ensemble(128, start=9000)
    g, o = case.run_gpu(y0), case.run_oracle(y0)
    K.compare_batch(g, o, case, "scan")
    assert (g["nEvents"] >= 2).all()


@pytest.mark.parametrize("event_mode", [1, 2])
def test_event_modes_are_equivalent(event_mode):
    """Events located inside the stepping kernel (1) and parked + located in batches (2) both equal the
    oracle: README two-event set, TERMINATE, RESTARTINT, unmasked sub-step scanning with CHANGESOLY (in
    mode 2 that parks on nearly every step: many rounds, then the in-loop mop-up pass), the two-body
    apsis event (CONTINUE without MASK: an event per half period) and events + dense output."""
    for eventset in (F.FLINT_EVSET_XY_CROSSING, F.FLINT_EVSET_CR3BP_SAMPLE3_TERM, F.FLINT_EVSET_CR3BP_SAMPLE3_RESTART):
        case = K.cr3bp_case(F.ERK_DOP853, 1e-11, 0, eventset=eventset)
        y0 = P.cr3bp_ensemble(200, start=7000)
        K.compare_batch(case.run_gpu(y0, event_mode=event_mode), case.run_oracle(y0), case, "mode %d evset %d" % (event_mode, eventset))
    case = K.cr3bp_case(F.ERK_DOP54, 1e-9, 1, norb=0.3, event_mask=[1, 1, 1], event_stepsz=[1e-4, 0.0, 0.0],
                        event_options=[F.FLINT_EVENTOPTION_STEPBEGIN, F.FLINT_EVENTOPTION_ROOTFINDING, F.FLINT_EVENTOPTION_STEPEND],
                        max_events=8)
    y0 = P.cr3bp_ensemble(64, start=9100)
    K.compare_batch(case.run_gpu(y0, event_mode=event_mode), case.run_oracle(y0), case, "mode %d scan" % event_mode)
    case = K.twobody_case(F.ERK_VERNER65E, 1e-10, 0, nper=6, eventset=F.FLINT_EVSET_APSIS, max_events=16)
    y0 = P.perturbed_ensemble(P.TBP_Y0, 128)
    g = case.run_gpu(y0, event_mode=event_mode)
    K.compare_batch(g, case.run_oracle(y0), case, "mode %d apsis" % event_mode)
    assert (g["nEvents"] >= 11).all()
    case = K.cr3bp_case(F.ERK_VERNER98R, 1e-10, 1, norb=1.0, interp_on=True, eventset=F.FLINT_EVSET_XY_CROSSING)
    y0 = P.cr3bp_ensemble(64, start=444)
    xarr = np.linspace(0.0, case.xf, 33)
    xarr[-1] = case.xf
    g, o = case.run_gpu(y0, event_mode=event_mode), case.run_oracle(y0, xarr=xarr)
    K.compare_batch(g, o, case, "mode %d dense" % event_mode)
    st, yarr, ist = g["erk"].InterpolateBatch(xarr, 64)
    assert st == 0 and (ist == 0).all()
    K.assert_bit_equal(yarr, o["yarr"], "Yarr mode %d" % event_mode)


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("method", METHODS)
def test_plane_variant_is_bit_identical(method, arith):
    """Planar CR3BP ensembles with z = vz = +0 run the kernel that drops the two identically-zero components
    (automatic); forcing the general kernel (plane_variant=2) and the oracle give the same bits.  An ensemble with
    one out-of-plane orbit, or with a -0.0, must take the general kernel by itself."""
    for eventset in (F.FLINT_EVSET_NONE, F.FLINT_EVSET_CR3BP_SAMPLE3, F.FLINT_EVSET_XY_CROSSING):
        case = K.cr3bp_case(method, 1e-10, arith, norb=1.2, eventset=eventset)
        y0 = P.cr3bp_ensemble(300, start=2500)
        o = case.run_oracle(y0)
        K.compare_batch(case.run_gpu(y0), o, case, "plane auto evset %d" % eventset)
        K.compare_batch(case.run_gpu(y0, plane_variant=2), o, case, "plane off evset %d" % eventset)
    case = K.cr3bp_case(method, 1e-10, arith, norb=1.2)
    y0 = P.cr3bp_ensemble(300, start=2500)
    y0[2, 17] = 1e-3
    y0[5, 99] = -0.0
    K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "out of plane")


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("method", METHODS)
def test_lorenz_no_events(method, arith):
    case = K.lorenz_case(method, 1e-11, arith, xf=10.0)
    y0 = P.lorenz_ensemble(512)
    K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "lorenz " + MNAME[method])


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("method", METHODS)
def test_halo_and_twobody(method, arith):
    case = K.halo_case(method, 1e-10, arith)
    y0 = P.perturbed_ensemble(P.HALO_Y0, 256)
    K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "halo " + MNAME[method])
    case = K.twobody_case(method, 1e-11, arith, nper=2, eventset=F.FLINT_EVSET_APSIS, max_events=8)
    y0 = P.perturbed_ensemble(P.TBP_Y0, 256)
    g, o = case.run_gpu(y0), case.run_oracle(y0)
    K.compare_batch(g, o, case, "twobody " + MNAME[method])
    assert (g["nEvents"] >= 3).all()      # apsis crossings over two periods


@pytest.mark.parametrize("arith", [0, 1])
@pytest.mark.parametrize("method", METHODS)
def test_dense_output_and_interpolate(method, arith):
    """InterpOn + Interpolate on a 257-point grid (test_CR3BP part B shape), events on."""
    case = K.cr3bp_case(method, 1e-10, arith, norb=1.0, interp_on=True)
    N = 96
    y0 = P.cr3bp_ensemble(N, start=333)
    G = 257
    xarr = np.array([case.xf / (G - 1) * i for i in range(G)])
    xarr[-1] = case.xf
    g = case.run_gpu(y0)
    o = case.run_oracle(y0, xarr=xarr)
    K.compare_batch(g, o, case, "dense " + MNAME[method])
    for layout in (0, 1):
        st, yarr, ist = g["erk"].InterpolateBatch(xarr, N, SaveInterpCoeffs=True, layout=layout)
        assert st == 0 and (ist == 0).all()
        ya = yarr if layout == 0 else np.transpose(yarr, (2, 1, 0))
        K.assert_bit_equal(ya, o["yarr"], "Yarr layout %d" % layout)
    # grid outside the integrated span -> FLINT_ERROR_INTP_ARRAY per trajectory
    st, _, ist = g["erk"].InterpolateBatch(xarr * 1.5, N, SaveInterpCoeffs=True)
    assert (ist == F.FLINT_ERROR_INTP_ARRAY).all()
    # default (no Save): storage is released, a second Interpolate reports it
    st, _, _ = g["erk"].InterpolateBatch(xarr, N)
    assert st == 0
    st, _, _ = g["erk"].InterpolateBatch(xarr, N)
    assert st == F.FLINT_ERROR_INTP_OFF


def test_interpolate_mappings_chunks_and_dense_grids():
    """Both interpolation kernels (warp per trajectory for [N][G][nI] and small ensembles, thread per trajectory for
    [nI][G][N] with N >= 4096), grids split into chunks (few trajectories, many points), non-uniform grids, and a grid
    much finer / much coarser than the steps (several steps between two points near the lunar fly-by)."""
    case = K.cr3bp_case(F.ERK_DOP853, 1e-9, 1, norb=1.0, interp_on=True)
    N = 4096
    y0 = P.cr3bp_ensemble(N, start=50)
    rng = np.random.default_rng(7)
    xarr = np.sort(np.concatenate([[0.0, case.xf], rng.uniform(0.0, case.xf, 60), np.linspace(case.xf * 0.49, case.xf * 0.51, 40)]))
    g = case.run_gpu(y0)
    o = case.run_oracle(y0, xarr=xarr)
    K.compare_batch(g, o, case, "dense 4096")
    for layout in (0, 1):
        st, yarr, ist = g["erk"].InterpolateBatch(xarr, N, SaveInterpCoeffs=True, layout=layout)
        assert st == 0
        ya = yarr if layout == 0 else np.transpose(yarr, (2, 1, 0))
        okt = ist == 0
        assert okt.mean() > 0.99
        K.assert_bit_equal(ya[okt], o["yarr"][okt], "Yarr 4096 layout %d" % layout)
    for G in (5, 7001):                      # coarser than any step; many chunks of a fine grid
        case = K.cr3bp_case(F.ERK_VERNER65E, 1e-8, 0, norb=1.0, interp_on=True, interp_states=[2, 5, 1], eventset=F.FLINT_EVSET_NONE)
        y0 = P.cr3bp_ensemble(5, start=900)
        xarr = np.linspace(0.0, case.xf, G)
        xarr[-1] = case.xf
        g, o = case.run_gpu(y0), case.run_oracle(y0, xarr=xarr)
        for layout in (0, 1):
            st, yarr, ist = g["erk"].InterpolateBatch(xarr, 5, SaveInterpCoeffs=True, layout=layout)
            assert st == 0 and (ist == 0).all()
            K.assert_bit_equal(yarr if layout == 0 else np.transpose(yarr, (2, 1, 0)), o["yarr"], "Yarr G=%d layout %d" % (G, layout))


def test_interp_states_subset():
    case = K.halo_case(F.ERK_DOP853, 1e-9, 0, interp_on=True, interp_states=[1, 3, 5])
    y0 = P.perturbed_ensemble(P.HALO_Y0, 64)
    xarr = np.linspace(0.0, case.xf, 50)
    xarr[-1] = case.xf
    g, o = case.run_gpu(y0), case.run_oracle(y0, xarr=xarr)
    K.compare_batch(g, o, case, "subset")
    st, yarr, ist = g["erk"].InterpolateBatch(xarr, 64, SaveInterpCoeffs=False)
    K.assert_bit_equal(yarr, o["yarr"], "subset Yarr")


@pytest.mark.parametrize("arith", [0, 1])
def test_natural_steps_output(arith):
    """IntStepsOn: Xint/Yint at the integrator's own steps (test_WP's sparse run)."""
    case = K.halo_case(F.ERK_VERNER98R, 1e-9, arith, int_steps_on=True)
    y0 = P.perturbed_ensemble(P.HALO_Y0, 40)
    g = case.run_gpu(y0, StepsCap=400)
    o = case.run_oracle(y0)
    K.compare_batch(g, o, case, "intsteps")
    sc = K.halo_case(F.ERK_VERNER98R, 1e-9, arith, int_steps_on=True).oracle_scalar()
    for i in (0, 17, 39):
        r = sc.integrate(0.0, y0[:, i], case.xf, io=case.oracle_int(), xint_cap=400)
        k = int(g["nXint"][i])
        assert k == r["n_xint"] == r["accepted"] + 1
        K.assert_bit_equal(g["Xint"][:k, i], r["xint"], "Xint")
        K.assert_bit_equal(g["Yint"][:, :k, i].T, r["yint"], "Yint")


def test_ragged_sizes_and_backward():
    """N = 1, N not a multiple of the warp/block size, per-trajectory Xf, backward integration."""
    for N in (1, 31, 33, 257):
        case = K.twobody_case(F.ERK_DOP853, 1e-10, 0, nper=1)
        y0 = P.perturbed_ensemble(P.TBP_Y0, N, start=N)
        K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "N=%d" % N)
    case = K.twobody_case(F.ERK_DOP54, 1e-9, 1, nper=1)
    y0 = P.perturbed_ensemble(P.TBP_Y0, 100)
    case.xf = -np.linspace(0.3, 1.0, 100) * P.tbp_period()       # ragged, negative direction
    K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "backward ragged")


def test_options_const_step_advance_maxsteps_vector_tol():
    y0 = P.perturbed_ensemble(P.TBP_Y0, 64)
    # constant step size
    case = K.twobody_case(F.ERK_VERNER65E, 1e-9, 0, nper=0.05, const_step=True, stepsz=7.5)
    K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "const step")
    # wrong sign of the constant step -> FLINT_ERROR_CONSTSTEPSZ
    case = K.twobody_case(F.ERK_DOP853, 1e-9, 0, nper=0.05, const_step=True, stepsz=-7.5)
    g, o = case.run_gpu(y0), case.run_oracle(y0)
    assert (g["status"] == F.FLINT_ERROR_CONSTSTEPSZ).all() and (o["status"] == F.FLINT_ERROR_CONSTSTEPSZ).all()
    # StepAdvance: exactly one accepted step
    case = K.twobody_case(F.ERK_DOP853, 1e-9, 1, nper=1, step_advance=True)
    g, o = case.run_gpu(y0), case.run_oracle(y0)
    K.compare_batch(g, o, case, "step advance")
    assert (g["counters"][1] == 1).all()
    # MaxSteps exhausted
    case = K.twobody_case(F.ERK_DOP54, 1e-12, 0, nper=10)
    case.max_steps = 50
    g, o = case.run_gpu(y0), case.run_oracle(y0)
    K.compare_batch(g, o, case, "maxsteps")
    assert (g["status"] == F.FLINT_ERROR_MAXSTEPS).all()
    # vector tolerances
    case = K.twobody_case(F.ERK_DOP853, 1e-10, 0, nper=1)
    case.rtol = np.array([1e-10, 1e-9, 1e-10, 1e-8, 1e-9, 1e-10])
    case.atol = case.rtol * 1e-3
    K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "vector tol")


def test_scalar_api_reproduces_reference_goldens():
    """The FLINT-shaped scalar calls (Init/Integrate/Interpolate/Info) through the C ABI reproduce
    the reference's work-precision golden FCalls (tests/wp_DOP853.txt:2-10, rounding-robust cells)."""
    gold = {1: 1825, 2: 1744, 3: 1789, 4: 1793, 5: 1860, 6: 1975, 7: 2113, 8: 2266, 9: 2542}
    sys_ = F.DiffEqSys(F.FLINT_SYS_CR3BP_SPATIAL, F.FLINT_EVSET_NONE, [P.HALO_MU])
    tip = np.array([P.HALO_PERIOD / 14 * i for i in range(15)])
    for i, fc in gold.items():
        erk = F.ERK()
        assert erk.Init(sys_, 99000, Method=F.ERK_DOP853, InterpOn=True, RTol=[P.wp_rtol(i)], ATol=[1e-80]) == 0
        r = erk.Integrate(0.0, P.HALO_Y0, P.HALO_PERIOD, StepSz=0.0, StiffTest=1)
        assert r["status"] == 0
        info = erk.Info()
        assert info["nFCalls"] == fc, (i, info["nFCalls"], fc)
        st, yip = erk.Interpolate(tip, SaveInterpCoeffs=False)
        assert st == 0 and np.isfinite(yip).all()
        assert np.array_equal(yip[0], P.HALO_Y0)


def test_device_resident_buffers_match_host_path():
    """FLINT_MEM_DEVICE (torch CUDA tensors, nothing copied) gives the same bits as FLINT_MEM_HOST."""
    import torch
    case = K.cr3bp_case(F.ERK_DOP853, 1e-11, 1)
    y0 = P.cr3bp_ensemble(1000)
    g = case.run_gpu(y0)
    d = case.run_gpu(torch.from_numpy(y0).cuda())
    torch.cuda.synchronize()
    for key in ("Yf", "Xf", "StepSz", "status", "counters", "nEvents", "EventStates"):
        K.assert_bit_equal(d[key].cpu().numpy(), g[key], key)


def test_async_schedule_matches_synchronous_call():
    """flint_exec.async: no host read-back between passes -- a fixed schedule of step / event kernels (empty ones return at
    once) that ends with an in-loop-event pass.  One event per orbit (one round), two (two rounds) and an event per
    half period (more rounds than the schedule has: the last pass mops up) must equal the synchronous call and the oracle."""
    import torch
    cases = [(K.cr3bp_case(F.ERK_DOP853, 1e-11, 0), P.cr3bp_ensemble(700, start=31)),
             (K.cr3bp_case(F.ERK_DOP54, 1e-9, 1, eventset=F.FLINT_EVSET_XY_CROSSING), P.cr3bp_ensemble(300, start=77)),
             (K.twobody_case(F.ERK_VERNER98R, 1e-10, 0, nper=5, eventset=F.FLINT_EVSET_APSIS, max_events=16), P.perturbed_ensemble(P.TBP_Y0, 200)),
             (K.lorenz_case(F.ERK_VERNER65E, 1e-9, 1, xf=5.0), P.lorenz_ensemble(300))]
    for case, y0 in cases:
        o = case.run_oracle(y0)
        g = case.run_gpu(torch.from_numpy(y0).cuda(), async_=True)
        torch.cuda.synchronize()
        g = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in g.items()}
        K.compare_batch(g, o, case, "async")


def test_results_do_not_depend_on_the_launch_shape():
    """flint_exec.block_threads / blocks_per_sm lower the CTA size / the resident CTAs of the persistent step kernels (whose own
    shape is a property of the system: 512 x 1, 256 x 2 or 128 x 4); every item goes through the work counter, so any shape
    gives the same bits."""
    cases = [(K.cr3bp_case(F.ERK_DOP853, 1e-11, 0), P.cr3bp_ensemble(3000, start=11)),
             (K.halo_case(F.ERK_VERNER98R, 1e-9, 1), P.perturbed_ensemble(P.HALO_Y0, 700)),
             (K.lorenz_case(F.ERK_DOP54, 1e-9, 0, xf=5.0), P.lorenz_ensemble(900))]
    for case, y0 in cases:
        o = case.run_oracle(y0)
        for kw in (dict(), dict(block_threads=32), dict(block_threads=128, blocks_per_sm=1), dict(block_threads=96, blocks_per_sm=3),
                   dict(blocks_per_sm=1), dict(block_threads=1024)):
            K.compare_batch(case.run_gpu(y0, **kw), o, case, "launch shape %r" % (kw,))


@pytest.mark.parametrize("arith", [0, 1])
def test_attempts_outside_the_fast_path_range_are_repeated_with_the_plain_operators(arith):
    """The hot loop's right-hand sides use branch-free division / square root and only record that an operand left their
    range; the lane loop then repeats the attempt with the stage-by-stage fallback to the plain operators.  A mass parameter
    outside [2^-60, 2^60) fails the range test on every attempt, so every step of these runs takes that path -- plane
    variant, general planar kernel (out-of-plane members), spatial system -- and must still equal the oracle bit for bit."""
    mu = 1e-30
    for method in (F.ERK_DOP853, F.ERK_DOP54):
        case = K.cr3bp_case(method, 1e-10, arith, norb=0.4, max_events=8)
        case.sysparams = [mu]
        y0 = P.cr3bp_ensemble(200, start=3)
        K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "planar, mu = 1e-30")
        y0[2] = 1e-3
        K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "planar general kernel, mu = 1e-30")
        case = K.halo_case(method, 1e-9, arith)
        case.sysparams = [mu]
        y0 = P.perturbed_ensemble(P.HALO_Y0, 150)
        K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "spatial, mu = 1e-30")
    # two-body in other units: GM' = GM * L^3 with lengths scaled by L = 2.15e30 is the same orbit with the same period, but
    # GM' = 4e96 and r^2 = 4e68 are outside the range of the branch-free operators
    case = K.twobody_case(F.ERK_DOP853, 1e-10, arith, nper=1)
    L = 2.1544346900318754e30
    case.sysparams = [P.TBP_GM * L ** 3]
    y0 = P.perturbed_ensemble(P.TBP_Y0, 120) * L
    K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "two-body, GM = %g" % case.sysparams[0])


def test_host_pipeline_chunks_match_single_chunk_and_oracle():
    """FLINT_MEM_HOST with flint_exec.pipeline = k: the ensemble runs as k chunks of trajectories on two streams (pitched
    H2D / D2H copies of column ranges of the SoA arrays, fixed launch schedule).  Ragged chunk sizes, more chunks than
    streams, events with one and with many rounds, no events: every output equals the one-chunk call and the oracle."""
    cases = [(K.cr3bp_case(F.ERK_DOP853, 1e-11, 0), P.cr3bp_ensemble(1000, start=5), (3, 7)),
             (K.twobody_case(F.ERK_VERNER98R, 1e-10, 0, nper=5, eventset=F.FLINT_EVSET_APSIS, max_events=16), P.perturbed_ensemble(P.TBP_Y0, 300), (2,)),
             (K.lorenz_case(F.ERK_DOP54, 1e-9, 1, xf=5.0), P.lorenz_ensemble(517), (4,))]
    for case, y0, chunkings in cases:
        o = case.run_oracle(y0)
        one = case.run_gpu(y0, pipeline=1)
        K.compare_batch(one, o, case, "one chunk")
        for k in chunkings:
            g = case.run_gpu(y0, pipeline=k)
            K.compare_batch(g, o, case, "pipeline=%d" % k)


@pytest.mark.parametrize("seed", range(int(os.environ.get("FLINT_FUZZ_SEEDS", "6"))))   # more seeds for a stress run
def test_randomised_option_combinations_against_the_oracle(seed):
    """Differential test: seeded random combinations of system, method, arithmetic, tolerances (scalar / vector), event set,
    event options / step sizes / masks, direction, constant-step, single-step, MaxSteps, natural-step output, stiffness mode,
    event-handling mode and the host-buffer chunk pipeline -- every output bit for bit equal to the CPU oracle."""
    rng = np.random.default_rng(20240607 + seed)
    for trial in range(8):
        method = int(rng.choice(METHODS))
        arith = int(rng.integers(0, 2))
        rtol = float(10.0 ** -rng.integers(6, 12))
        which = int(rng.integers(0, 4))
        kw = {}
        extra = {}
        if which == 0:                                   # planar CR3BP with one of its event sets
            evs = int(rng.choice([F.FLINT_EVSET_NONE, F.FLINT_EVSET_CR3BP_SAMPLE3, F.FLINT_EVSET_XY_CROSSING,
                                  F.FLINT_EVSET_CR3BP_SAMPLE3_TERM, F.FLINT_EVSET_CR3BP_SAMPLE3_RESTART]))
            if evs in (1, 4, 5):
                kw.update(event_mask=[int(v) for v in rng.integers(0, 2, 3)],
                          event_stepsz=[float(rng.choice([0.0, 1e-5, 0.05])), 0.0, float(rng.choice([0.0, 0.1]))],
                          event_options=[int(v) for v in rng.integers(0, 3, 3)],
                          event_tol=[1e-3, float(rng.choice([1e-9, 1e-6])), 1e-9])
            elif evs == 2:
                kw.update(event_options=[int(v) for v in rng.integers(0, 3, 2)], event_tol=[float(rng.choice([1e-9, 1e-7]))] * 2)
            case = K.cr3bp_case(method, rtol, arith, eventset=evs, norb=float(rng.choice([0.3, 1.0])), max_events=8, **kw)
            y0 = P.cr3bp_ensemble(int(rng.integers(40, 200)), start=int(rng.integers(0, 1 << 20)))
            if rng.integers(0, 3) == 0:
                y0[2] = 1e-3 * rng.standard_normal(y0.shape[1])       # some members out of the plane: the general kernel
            if evs != F.FLINT_EVSET_NONE:
                extra["event_mode"] = int(rng.integers(0, 3))
        elif which == 1:
            case = K.lorenz_case(method, rtol, arith, xf=float(rng.choice([2.0, 6.0])))
            y0 = P.lorenz_ensemble(int(rng.integers(40, 200)), start=int(rng.integers(0, 1 << 20)))
        elif which == 2:
            case = K.halo_case(method, rtol, arith)
            y0 = P.perturbed_ensemble(P.HALO_Y0, int(rng.integers(40, 200)), start=int(rng.integers(0, 1 << 20)))
        else:
            evs = int(rng.choice([F.FLINT_EVSET_NONE, F.FLINT_EVSET_APSIS]))
            case = K.twobody_case(method, rtol, arith, nper=float(rng.choice([0.5, 2.0])), eventset=evs, max_events=8)
            y0 = P.perturbed_ensemble(P.TBP_Y0, int(rng.integers(40, 200)), start=int(rng.integers(0, 1 << 20)))
            if rng.integers(0, 2):
                case.xf = -case.xf                                    # backward
        N = y0.shape[1]
        opt = int(rng.integers(0, 6))
        if opt == 0:
            case.rtol = rtol * 10.0 ** rng.integers(0, 3, case.n)     # vector tolerances
            case.atol = case.rtol * 1e-3
        elif opt == 1:
            case.step_advance = True
        elif opt == 2:
            case.max_steps = int(rng.integers(5, 60))
        elif opt == 3 and case.interp_on is None:
            case.int_steps_on = True
            extra["StepsCap"] = 400
            case.max_steps = 399
        elif opt == 4:
            case.stiff_test = int(rng.integers(0, 3))
        if rng.integers(0, 3) == 0:
            case.xf = case.xf * rng.uniform(0.5, 1.0, N)              # ragged end points
        if rng.integers(0, 2):
            extra["pipeline"] = int(rng.integers(2, 5))
        what = "seed %d trial %d: system %d method %d arith %d rtol %g opt %d %r %r" % (seed, trial, which, method, arith, rtol, opt, kw, extra)
        o = case.run_oracle(y0)
        g = case.run_gpu(y0, **extra)
        K.compare_batch(g, o, case, what)
        if case.int_steps_on:
            sc = case.oracle_scalar()
            for t in (0, N // 2, N - 1):
                r = sc.integrate(case.x0, y0[:, t], float(np.broadcast_to(case.xf, (N,))[t]), io=case.oracle_int(), xint_cap=400)
                c = int(g["nXint"][t])
                assert c == r["n_xint"], what
                K.assert_bit_equal(g["Xint"][:c, t], r["xint"], what + " Xint")
                K.assert_bit_equal(g["Yint"][:, :c, t].T, r["yint"], what + " Yint")


def test_full_size_properties():
    """BASELINE size (2^20 orbits): size-independent properties -- every trajectory succeeds, reaches
    Xf exactly, locates exactly one event of id 2 on the y-axis crossing, Jacobi drift stays at the
    tolerance level, and a seeded sample of the big run equals the oracle bit for bit."""
    import torch
    N = 1 << 20
    case = K.cr3bp_case(F.ERK_DOP853, 1e-12, 0)
    y0 = P.cr3bp_ensemble(N)
    g = case.run_gpu(torch.from_numpy(y0).cuda())
    torch.cuda.synchronize()
    st, cnt, xf = g["status"].cpu().numpy(), g["counters"].cpu().numpy(), g["Xf"].cpu().numpy()
    yf, ev, nev = g["Yf"].cpu().numpy(), g["EventStates"].cpu().numpy(), g["nEvents"].cpu().numpy()
    ok = st == 0
    assert ok.mean() > 0.995 and set(np.unique(st)) <= {0, F.FLINT_ERROR_STEPSZ_TOOSMALL}   # near-collision orbits give up, as in the oracle
    assert (xf[ok] == case.xf).all() and (nev == 1).all()
    assert (ev[7, 0] == 2.0).all() and (np.abs(ev[2, 0]) < 1e-8).all() and (ev[1, 0] > 0).all()
    assert (cnt[0] >= cnt[1] + cnt[2]).all()          # attempts = accepted + rejected (+ rejections before the first acceptance)

    def jac(y):
        mu = P.CR3BP_MU
        r1 = np.sqrt((y[0] + mu) ** 2 + y[1] ** 2 + y[2] ** 2)
        r2 = np.sqrt((y[0] - 1 + mu) ** 2 + y[1] ** 2 + y[2] ** 2)
        return (y[3] ** 2 + y[4] ** 2 + y[5] ** 2) / 2 - (0.5 * (y[0] ** 2 + y[1] ** 2) + (1 - mu) / r1 + mu / r2)
    drift = np.abs(jac(yf) - jac(y0))[ok]
    assert np.median(drift) < 1e-10 and np.quantile(drift, 0.99) < 1e-6, (np.median(drift), np.quantile(drift, 0.99))   # near-collision passes amplify            # reference DOP853 @1e-11: 2.9e-11 (results_CR3BP.txt:6); close approaches amplify
    sel = np.arange(0, N, N // 512)[:512]
    o = case.run_oracle(np.ascontiguousarray(y0[:, sel]))
    K.assert_bit_equal(st[sel], o["status"], "sampled status")
    K.assert_bit_equal(yf[:, sel], o["yf"], "sampled Yf")
    K.assert_bit_equal(cnt[:, sel], o["counters"], "sampled counters")
