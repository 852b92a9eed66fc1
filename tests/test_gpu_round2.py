"""GPU parity tests added in round 2: the gang-scheduled step kernel (small shards), the ensemble sharded over devices inside one
call, handle re-use on the host-buffer path with multi-round events, defined outputs of rejected trajectories, and the
tolerance / size corners of BASELINE.json's configs that round 1 left without a bit-parity check.  Bar as in
test_gpu_parity.py: BIT-EXACT against the CPU oracle."""
import numpy as np
import pytest

import flint_b200 as F
from flint_b200 import problems as P
import cases as K

pytestmark = pytest.mark.gpu

METHODS = [F.ERK_DOP853, F.ERK_DOP54, F.ERK_VERNER98R, F.ERK_VERNER65E]
MNAME = {17: "DOP853", 18: "DOP54", 19: "Verner98R", 20: "Verner65E"}


# ------------------------------------------------------------------------------------------------ gang-scheduled step kernel
@pytest.mark.parametrize("arith", [0, 1])
def test_gang_scheduled_step_kernel_is_bit_identical(arith):
    """flint_exec.round_len = k: the step kernel time-shares its lanes in rounds of k attempts, each round on the trajectories
    furthest behind in x (erk_kernel.cuh: GANG).  Parked events (one round, two rounds, an event per half period), no events,
    systems with and without the per-attempt barrier, backward and ragged spans: same bits as the oracle for every k."""
    cases = [(K.cr3bp_case(F.ERK_DOP853, 1e-12, arith), P.cr3bp_ensemble(3000, start=21)),
             (K.cr3bp_case(F.ERK_VERNER98R, 1e-10, arith, eventset=F.FLINT_EVSET_XY_CROSSING, norb=1.0), P.cr3bp_ensemble(700, start=4)),
             (K.cr3bp_case(F.ERK_DOP54, 1e-9, arith, eventset=F.FLINT_EVSET_NONE, norb=1.0), P.cr3bp_ensemble(900, start=77)),
             (K.twobody_case(F.ERK_VERNER65E, 1e-10, arith, nper=4, eventset=F.FLINT_EVSET_APSIS, max_events=16), P.perturbed_ensemble(P.TBP_Y0, 300)),
             (K.lorenz_case(F.ERK_DOP54, 1e-9, arith, xf=6.0), P.lorenz_ensemble(1100)),
             (K.halo_case(F.ERK_DOP853, 1e-10, arith), P.perturbed_ensemble(P.HALO_Y0, 600))]
    for case, y0 in cases:
        o = case.run_oracle(y0)
        for k in (2, 7, 64, 1000):
            K.compare_batch(case.run_gpu(y0, round_len=k), o, case, "round_len=%d" % k)
    # out-of-plane members (general planar kernel), backward ragged spans, tiny CTAs (many more trajectories than lanes per CTA: the histogram selection)
    case = K.cr3bp_case(F.ERK_DOP853, 1e-10, arith, norb=0.8)
    y0 = P.cr3bp_ensemble(500, start=9)
    y0[2, ::7] = 1e-3
    K.compare_batch(case.run_gpu(y0, round_len=5, block_threads=64), case.run_oracle(y0), case, "general kernel, rounds")
    case = K.twobody_case(F.ERK_DOP54, 1e-9, arith, nper=1)
    y0 = P.perturbed_ensemble(P.TBP_Y0, 333)
    case.xf = -np.linspace(0.3, 1.0, 333) * P.tbp_period()
    K.compare_batch(case.run_gpu(y0, round_len=9, block_threads=32, blocks_per_sm=1), case.run_oracle(y0), case, "backward ragged, rounds")


def test_gang_scheduled_dense_output_and_device_buffers():
    """The gang-scheduled kernel with InterpOn (step logs), and on device-resident buffers."""
    import torch
    case = K.cr3bp_case(F.ERK_VERNER98R, 1e-10, 0, norb=1.0, interp_on=True, eventset=F.FLINT_EVSET_NONE)
    N = 400
    y0 = P.cr3bp_ensemble(N, start=600)
    xarr = np.linspace(0.0, case.xf, 301)
    xarr[-1] = case.xf
    o = case.run_oracle(y0, xarr=xarr)
    g = case.run_gpu(y0, round_len=6)
    K.compare_batch(g, o, case, "rounds, dense")
    st, yarr, ist = g["erk"].InterpolateBatch(xarr, N)
    assert st == 0 and (ist == 0).all()
    K.assert_bit_equal(yarr, o["yarr"], "rounds, Yarr")
    case = K.cr3bp_case(F.ERK_DOP853, 1e-11, 1)
    y0 = P.cr3bp_ensemble(2000, start=5)
    d = case.run_gpu(torch.from_numpy(y0).cuda(), round_len=4)
    torch.cuda.synchronize()
    d = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in d.items()}
    K.compare_batch(d, case.run_oracle(y0), case, "rounds, device buffers")


def test_small_shard_takes_the_gang_scheduled_kernel_by_itself():
    """The strong-scaling shard of the metric (2^20 orbits / 8 GPUs = 2^17 per GPU, 1.73 orbits per lane): the automatic choice
    (rounds) and round_len=1 (a trajectory stays on its lane until it finishes) give the same bits, and a sample equals the oracle."""
    import torch
    N = 1 << 17
    case = K.cr3bp_case(F.ERK_DOP853, 1e-12, 0)
    y0 = P.cr3bp_ensemble(N)
    y0d = torch.from_numpy(y0).cuda()
    a = case.run_gpu(y0d)
    b = case.run_gpu(y0d, round_len=1)
    torch.cuda.synchronize()
    for key in ("Yf", "Xf", "StepSz", "status", "counters", "nEvents", "EventStates", "StiffTest"):
        K.assert_bit_equal(a[key].cpu().numpy(), b[key].cpu().numpy(), key)
    sel = np.arange(0, N, N // 256)[:256]
    o = case.run_oracle(np.ascontiguousarray(y0[:, sel]))
    K.assert_bit_equal(a["Yf"].cpu().numpy()[:, sel], o["yf"], "sampled Yf")
    K.assert_bit_equal(a["counters"].cpu().numpy()[:, sel], o["counters"], "sampled counters")


# ------------------------------------------------------------------------------------------------ sharded call
def _devices():
    import torch
    n = torch.cuda.device_count()
    return [list(range(n))] if n >= 2 else []


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0]] + _devices())
def test_sharded_call_equals_the_one_device_call(devices):
    """flint_exec.n_devices > 1: chunks of shard_chunk trajectories dealt round-robin over the devices, each shard on its own
    stream from its own host thread, results at the caller's offsets.  A device listed twice shares the GPU between two shards,
    which is how a one-GPU box exercises the deal; with >= 2 GPUs the real device list runs too.  Every output equals the
    one-device call bit for bit: full chunks, a ragged tail, N smaller than one chunk, rows that do / do not line up with the deal."""
    D = len(devices)
    for N, chunk, case in ((3 * 1000 + 417, 1000, K.cr3bp_case(F.ERK_DOP853, 1e-11, 0)),
                           (D * 512 * 2, 512, K.cr3bp_case(F.ERK_DOP54, 1e-9, 1, eventset=F.FLINT_EVSET_XY_CROSSING, norb=1.0)),
                           (300, 4096, K.lorenz_case(F.ERK_VERNER65E, 1e-9, 0, xf=4.0)),
                           (2049, 256, K.twobody_case(F.ERK_DOP853, 1e-10, 0, nper=2, eventset=F.FLINT_EVSET_APSIS, max_events=8))):
        y0 = {1: P.cr3bp_ensemble, 3: P.lorenz_ensemble}.get(case.system_id, lambda n: P.perturbed_ensemble(P.TBP_Y0, n))(N)
        case.xf = case.xf * np.linspace(0.8, 1.0, N)                    # per-trajectory end points travel with their trajectories
        one = case.run_gpu(y0)
        sh = case.run_gpu(y0, devices=devices, shard_chunk=chunk)
        for key in ("Yf", "Xf", "StepSz", "status", "counters", "StiffTest") + (("nEvents", "EventStates") if case.events_on else ()):
            if one.get(key) is not None:
                K.assert_bit_equal(sh[key], one[key], "%s (N=%d chunk=%d devices=%r)" % (key, N, chunk, devices))
    case = K.cr3bp_case(F.ERK_DOP853, 1e-11, 0)
    y0 = P.cr3bp_ensemble(2500)
    K.compare_batch(case.run_gpu(y0, devices=devices, shard_chunk=300), case.run_oracle(y0), case, "sharded vs oracle")


@pytest.mark.parametrize("devices", [[0, 0]] + _devices())
def test_sharded_dense_output_and_interpolate(devices):
    """InterpOn on a sharded call: every device keeps the dense output of its shard; InterpolateBatch evaluates each shard where
    it lives and writes the caller's Yarr at the trajectories' own offsets (both layouts)."""
    case = K.cr3bp_case(F.ERK_VERNER98R, 1e-10, 0, norb=1.0, interp_on=True, eventset=F.FLINT_EVSET_NONE)
    N = 1100
    y0 = P.cr3bp_ensemble(N, start=40)
    xarr = np.linspace(0.0, case.xf, 129)
    xarr[-1] = case.xf
    o = case.run_oracle(y0, xarr=xarr)
    g = case.run_gpu(y0, devices=devices, shard_chunk=200)
    K.compare_batch(g, o, case, "sharded dense")
    for layout in (0, 1):
        st, yarr, ist = g["erk"].InterpolateBatch(xarr, N, SaveInterpCoeffs=True, layout=layout)
        assert st == 0 and (ist == 0).all()
        K.assert_bit_equal(yarr if layout == 0 else np.transpose(yarr, (2, 1, 0)), o["yarr"], "sharded Yarr layout %d" % layout)


# ------------------------------------------------------------------------------------------------ host path, handle re-use
def test_handle_reuse_on_the_host_path_with_multi_round_events():
    """One ERK handle integrates twice with host buffers; unmasked XY_CROSSING-style events recur, so trajectories park in
    several rounds and the event records grow after the first event kernel.  Both calls must return every record (the
    early copy of the records during the resume pass may only be kept when no trajectory parked again)."""
    case = K.twobody_case(F.ERK_DOP853, 1e-10, 0, nper=5, eventset=F.FLINT_EVSET_APSIS, max_events=16)
    y0 = P.perturbed_ensemble(P.TBP_Y0, 500)
    o = case.run_oracle(y0)
    sys_, erk, st = case.make_erk()
    assert st == 0
    for rep in range(3):
        g = erk.IntegrateBatch(case.x0, y0, case.xf, StepSz=0.0, MaxEvents=case.max_events, **case.int_kwargs())
        K.compare_batch(g, o, case, "call %d on one handle" % rep)
        assert (g["nEvents"] >= 9).all()
    # same handle, a different ensemble with the same event counts: stale staging data must not be taken for progress
    y1 = P.perturbed_ensemble(P.TBP_Y0, 500, start=500)
    g = erk.IntegrateBatch(case.x0, y1, case.xf, StepSz=0.0, MaxEvents=case.max_events, **case.int_kwargs())
    K.compare_batch(g, case.run_oracle(y1), case, "second ensemble on one handle")
    case = K.cr3bp_case(F.ERK_DOP853, 1e-11, 0, eventset=F.FLINT_EVSET_XY_CROSSING, max_events=4)
    y0 = P.cr3bp_ensemble(600, start=123)
    o = case.run_oracle(y0)
    sys_, erk, st = case.make_erk()
    for rep in range(2):
        g = erk.IntegrateBatch(case.x0, y0, case.xf, StepSz=0.0, MaxEvents=4, **case.int_kwargs())
        K.compare_batch(g, o, case, "xy crossing, call %d" % rep)


def test_rejected_trajectories_have_defined_outputs():
    """Per-trajectory argument errors (wrong sign of a constant step, StiffTest out of range) return before the integration
    (src/ERKIntegrate.f90:196-199): status set, Xf / StepSz as passed in, Yf = Y0, counters and event count 0 -- also on the
    host-buffer path, whose staging buffers hold the previous call's results."""
    case = K.twobody_case(F.ERK_DOP853, 1e-9, 0, nper=0.05, eventset=F.FLINT_EVSET_APSIS, max_events=4)
    y0 = P.perturbed_ensemble(P.TBP_Y0, 64)
    sys_, erk, st = case.make_erk()
    good = erk.IntegrateBatch(0.0, y0, case.xf, StepSz=7.5, UseConstStepSz=True, EventStates=True, StiffTest=1, MaxEvents=4)
    assert (good["status"] == 0).all() and (good["counters"][0] > 0).all()
    h = np.full(64, 7.5)
    h[::2] = -7.5
    stiff = np.ones(64, dtype=np.int32)
    stiff[3] = 7
    bad = erk.IntegrateBatch(0.0, y0, case.xf, StepSz=h, UseConstStepSz=True, EventStates=True, StiffTest=stiff, MaxEvents=4)
    rej = np.zeros(64, dtype=bool)
    rej[::2] = True
    assert (bad["status"][rej] == F.FLINT_ERROR_CONSTSTEPSZ).all() and bad["status"][3] == F.FLINT_ERROR_PARAMS
    rej[3] = True
    assert (bad["status"][~rej] == 0).all()
    K.assert_bit_equal(bad["Yf"][:, rej], y0[:, rej], "Yf of rejected trajectories")
    assert (bad["counters"][:, rej] == 0).all() and (bad["nEvents"][rej] == 0).all()
    assert (bad["Xf"][rej] == case.xf).all()
    K.assert_bit_equal(bad["Yf"][:, ~rej], good["Yf"][:, ~rej], "the others are unaffected")


def test_scalar_interpolate_refuses_a_batch_store():
    """flint_erk_interpolate writes G*nI doubles: it must not be handed the dense output of a batch Integrate."""
    case = K.cr3bp_case(F.ERK_DOP853, 1e-9, 0, norb=0.3, interp_on=True, eventset=F.FLINT_EVSET_NONE)
    g = case.run_gpu(P.cr3bp_ensemble(8))
    st, _ = g["erk"].Interpolate(np.linspace(0.0, case.xf, 5), SaveInterpCoeffs=True)
    assert st == F.FLINT_ERROR_INTP_OFF
    st, yarr, ist = g["erk"].InterpolateBatch(np.linspace(0.0, case.xf, 5), 8)
    assert st == 0 and (ist == 0).all()


# ------------------------------------------------------------------------------------------------ parity corners of the configs
@pytest.mark.parametrize("method", METHODS)
def test_config5_tight_tolerances(method):
    """Config 5 (work-precision sweep) at the round-off dominated end, rtol 1e-12 .. 1e-14, where the error estimate is noise
    and step counts are most arithmetic-sensitive: halo orbit (tests/test_WP.f90 setup, atol 1e-80) and the
    two-body orbit, both arithmetic modes."""
    for arith in (0, 1):
        for i in (12, 13, 14):
            rtol = P.wp_rtol(i)
            case = K.halo_case(method, rtol, arith)
            y0 = P.perturbed_ensemble(P.HALO_Y0, 48, start=100 * i)
            K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "halo %s rtol 1e-%d arith %d" % (MNAME[method], i, arith))
            case = K.twobody_case(method, rtol, arith, nper=2)
            case.max_steps = 100000
            y0 = P.perturbed_ensemble(P.TBP_Y0, 48, start=100 * i)
            K.compare_batch(case.run_gpu(y0), case.run_oracle(y0), case, "two-body %s rtol 1e-%d arith %d" % (MNAME[method], i, arith))


def test_config2_lorenz_full_span():
    """Config 2 at its real span x in [0, 100] (41 000 steps per trajectory, e^90 error amplification: any single
    differing bit would show) on 256 trajectories."""
    case = K.lorenz_case(F.ERK_DOP54, 1e-11, 0, xf=100.0)
    y0 = P.lorenz_ensemble(256)
    g, o = case.run_gpu(y0), case.run_oracle(y0)
    K.compare_batch(g, o, case, "lorenz x=100")
    assert (g["counters"][1] > 30000).all()


def test_config1_scalar_abi_forward_and_backward():
    """Config 1 through the FLINT-shaped scalar calls: two-body, DOP853, rtol 1e-11 and 1e-12, ten periods forward, then the
    explicit backward leg from the final state with StepSz = 0 (tests/test_TBP.f90:166-245 without quirk Q9)."""
    for rtol in (1e-11, 1e-12):
        case = K.twobody_case(F.ERK_DOP853, rtol, 0, nper=10)
        sys_, erk, st = case.make_erk()
        assert st == 0
        sc = case.oracle_scalar()
        r = erk.Integrate(0.0, P.TBP_Y0, case.xf, StepSz=0.0, StiffTest=1)
        o = sc.integrate(0.0, P.TBP_Y0, case.xf, io=case.oracle_int())
        assert r["status"] == 0 == o["status"]
        K.assert_bit_equal(r["Yf"], o["yf"], "forward Yf")
        info = erk.Info()
        assert (info["nSteps"], info["nAccept"], info["nReject"], info["nFCalls"]) == (o["total"], o["accepted"], o["rejected"], o["fcalls"])
        rb = erk.Integrate(case.xf, r["Yf"], 0.0, StepSz=0.0, StiffTest=1)
        ob = sc.integrate(case.xf, o["yf"], 0.0, io=case.oracle_int())
        assert rb["status"] == 0 == ob["status"] and rb["Xf"] == 0.0
        K.assert_bit_equal(rb["Yf"], ob["yf"], "backward Yf")
        assert np.abs(rb["Yf"] - P.TBP_Y0).max() < 1e-4 * np.abs(P.TBP_Y0).max()       # the orbit closes


def test_params_are_validated():
    """Integrate(params=...): up to 8 values are forwarded to user functors; more is FLINT_ERROR_PARAMS (never dropped silently)."""
    case = K.lorenz_case(F.ERK_DOP54, 1e-8, 0, xf=1.0)
    y0 = P.lorenz_ensemble(32)
    sys_, erk, st = case.make_erk()
    a = erk.IntegrateBatch(0.0, y0, 1.0, StepSz=0.0, params=[1.0, 2.0, 3.0])
    b = erk.IntegrateBatch(0.0, y0, 1.0, StepSz=0.0)
    K.assert_bit_equal(a["Yf"], b["Yf"], "built-in systems ignore params, as the reference's F")
    c = erk.IntegrateBatch(0.0, y0, 1.0, StepSz=0.0, params=np.arange(9.0))
    assert c["rc"] == F.FLINT_ERROR_PARAMS


# ------------------------------------------------------------------------------------------------ dense output, round 2
@pytest.mark.parametrize("dense_mode", [0, 1])
@pytest.mark.parametrize("method", METHODS)
def test_dense_store_grows_instead_of_overflowing(method, dense_mode):
    """The first level of the dense store is far too small (dense_cap = 128 steps): trajectories are suspended when it is full,
    get deeper levels and resume -- no step is lost (the reference sizes Xint/Bip for MaxSteps, src/ERKInit.f90:321-332), every
    trajectory interpolates, final states and counters are untouched.  Both store kinds (step logs + fused Interpolate;
    coefficient records), both layouts, events on (the event code's steps are logged with their truncated h)."""
    case = K.cr3bp_case(method, 1e-10, 0, norb=1.0, interp_on=True)
    N = 300
    y0 = P.cr3bp_ensemble(N, start=77)
    G = 401
    xarr = np.array([case.xf / (G - 1) * i for i in range(G)])
    xarr[-1] = case.xf
    o = case.run_oracle(y0, xarr=xarr)
    assert o["counters"][1].max() > 128                # more than the first level holds (DOP54 / Verner65E: > 2000 steps, four levels)
    g = case.run_gpu(y0, dense_cap=128, dense_mode=dense_mode)
    K.compare_batch(g, o, case, "grown store")
    for layout in (0, 1):
        st, yarr, ist = g["erk"].InterpolateBatch(xarr, N, SaveInterpCoeffs=True, layout=layout)
        assert st == 0 and (ist == 0).all()
        K.assert_bit_equal(yarr if layout == 0 else np.transpose(yarr, (2, 1, 0)), o["yarr"], "Yarr layout %d" % layout)


@pytest.mark.parametrize("arith", [0, 1])
def test_config4_shape_fused_and_separate(arith):
    """Config 4 at its real shape on a 512-orbit sample: Verner98R rtol 1e-11, InterpOn, G = 10 000 uniform points with the last one
    forced to Xf (tests/test_CR3BP.f90:345-347).  The fused path (step logs -> coefficients in registers -> Yarr) and the separate
    phases (coefficient records, then the evaluation kernel) both equal the oracle bit for bit; nobody returns INTP_ARRAY."""
    case = K.cr3bp_case(F.ERK_VERNER98R, 1e-11, arith, interp_on=True, eventset=F.FLINT_EVSET_NONE)
    N, G = 512, 10000
    y0 = P.cr3bp_ensemble(N, start=2048)
    xarr = np.array([case.xf / (G - 1) * i for i in range(G)])
    xarr[-1] = case.xf
    o = case.run_oracle(y0, xarr=xarr)
    for dense_mode in (0, 1):
        g = case.run_gpu(y0, dense_mode=dense_mode, dense_cap=384)
        K.compare_batch(g, o, case, "config 4, dense_mode %d" % dense_mode)
        st, yarr, ist = g["erk"].InterpolateBatch(xarr, N, SaveInterpCoeffs=True)
        assert st == 0 and (ist == 0).all()
        K.assert_bit_equal(yarr, o["yarr"], "config 4 Yarr, dense_mode %d" % dense_mode)
        st, yarr, ist = g["erk"].InterpolateBatch(xarr, N, layout=1)
        K.assert_bit_equal(np.transpose(yarr, (2, 1, 0)), o["yarr"], "config 4 Yarr [nI][G][N], dense_mode %d" % dense_mode)


@pytest.mark.parametrize("method,arith", [(F.ERK_VERNER98R, 0), (F.ERK_VERNER98R, 1), (F.ERK_DOP853, 0), (F.ERK_DOP54, 0), (F.ERK_VERNER65E, 1)])
def test_interpolation_kernels_agree_bit_for_bit(method, arith, monkeypatch):
    """The [N][G][nI] evaluation kernels -- thread-per-item with bulk stores (the default for an even nI), step-major CTAs and the
    persistent tile pipeline (FLINT_B200_INTERP_KERNEL) -- give the oracle's Yarr bit for bit: a non-uniform grid that leaves
    steps without points and packs hundreds of points into others (every item length 1..P, rounds that straddle tiles), grid
    points that coincide with the ends of the span, nI = 6 and the subset nI = 4 (InterpStates), a store that grew a level."""
    for istates in (None, [1, 2, 4, 5]):
        case = K.cr3bp_case(method, 1e-10, arith, interp_on=True, interp_states=istates, eventset=F.FLINT_EVSET_NONE)
        N, G = 200, 4001
        y0 = P.cr3bp_ensemble(N, start=512)
        u = np.linspace(0.0, 1.0, G)
        xarr = case.xf * (0.5 * u + 0.5 * u ** 7)                          # dense early, sparse late
        xarr[-1] = case.xf
        o = case.run_oracle(y0, xarr=xarr)
        g = case.run_gpu(y0, dense_mode=1, dense_cap=256)
        for kern in ("items", "steps", "tiles"):
            monkeypatch.setenv("FLINT_B200_INTERP_KERNEL", kern)
            st, yarr, ist = g["erk"].InterpolateBatch(xarr, N, SaveInterpCoeffs=True)
            ok = g["status"] == 0                                           # a trajectory that stopped early has no dense output up to Xf
            assert ok.sum() > N // 2 and (ist[ok] == 0).all()
            K.assert_bit_equal(yarr[ok], o["yarr"][ok], "Yarr, kernel %s, method %d, nI %d" % (kern, method, yarr.shape[2]))
        monkeypatch.delenv("FLINT_B200_INTERP_KERNEL")


@pytest.mark.parametrize("method,cap", [(F.ERK_VERNER98R, 1024), (F.ERK_DOP54, 128), (F.ERK_DOP853, 256)])
def test_staged_and_fused_evaluation_from_step_logs(method, cap, monkeypatch):
    """Interpolate [N][G][nI] from a store of step logs (dense_mode 0): STAGED through a temporary coefficient store in parts of
    trajectories (default; parts of 256 here, the last one ragged, a store that grew levels for DOP54) and FUSED in one kernel
    (FLINT_B200_DENSE_EVAL=fused) both equal the oracle bit for bit, with a shared grid and with one grid per trajectory."""
    case = K.cr3bp_case(method, 1e-9, 0, interp_on=True, eventset=F.FLINT_EVSET_NONE)
    N, G = 600, 1501
    y0 = P.cr3bp_ensemble(N, start=8192)
    xarr = np.linspace(0.0, case.xf, G); xarr[-1] = case.xf
    o = case.run_oracle(y0, xarr=xarr)
    g = case.run_gpu(y0, dense_mode=0, dense_cap=cap)
    ok = g["status"] == 0
    assert ok.sum() > N // 2
    grids = np.repeat(xarr[None, :], N, axis=0)
    monkeypatch.setenv("FLINT_B200_DENSE_PART", "256")
    for how in ("staged", "fused"):
        monkeypatch.setenv("FLINT_B200_DENSE_EVAL", how)
        st, yarr, ist = g["erk"].InterpolateBatch(xarr, N, SaveInterpCoeffs=True)
        assert st == 0 and (ist[ok] == 0).all()
        K.assert_bit_equal(yarr[ok], o["yarr"][ok], "Yarr from step logs, %s, method %d" % (how, method))
        st, yarr, ist = g["erk"].InterpolateBatch(grids, N, SaveInterpCoeffs=True)
        assert st == 0 and (ist[ok] == 0).all()
        K.assert_bit_equal(yarr[ok], o["yarr"][ok], "Yarr from step logs, per-trajectory grids, %s, method %d" % (how, method))


@pytest.mark.parametrize("bulk", ["0", "1"])
def test_coefficient_records_thread_by_thread_and_as_bulk_stores(bulk, monkeypatch):
    """erk_dense_kernel writes a block's coefficient records with one bulk store from shared memory (default) or thread by thread
    (FLINT_B200_DENSE_BULK=0): same records, same Yarr, with the in-place pass of Integrate (dense_mode 1) and with the temporary
    store of the staged Interpolate (dense_mode 0); DOP54's store grows two levels from 128 steps."""
    monkeypatch.setenv("FLINT_B200_DENSE_BULK", bulk)
    for method, cap in ((F.ERK_VERNER65E, 256), (F.ERK_DOP54, 128)):
        case = K.cr3bp_case(method, 1e-9, 1, interp_on=True, eventset=F.FLINT_EVSET_NONE)
        N, G = 300, 777
        y0 = P.cr3bp_ensemble(N, start=4096)
        xarr = np.linspace(0.0, case.xf, G); xarr[-1] = case.xf
        o = case.run_oracle(y0, xarr=xarr)
        for dense_mode in (1, 0):
            g = case.run_gpu(y0, dense_mode=dense_mode, dense_cap=cap)
            ok = g["status"] == 0
            st, yarr, ist = g["erk"].InterpolateBatch(xarr, N)
            assert st == 0 and ok.sum() > N // 2 and (ist[ok] == 0).all()
            K.assert_bit_equal(yarr[ok], o["yarr"][ok], "Yarr, bulk %s, dense_mode %d, method %d" % (bulk, dense_mode, method))


def test_interpolation_edge_grids(monkeypatch):
    """Grids the thread-per-item kernel's planning has to get right: one, two and three points; every point inside the first step;
    every point inside the last step; the exact ends of the span; forward and backward integration."""
    for sign in (1.0, -1.0):
        case = K.twobody_case(F.ERK_DOP853, 1e-10, 0, nper=1, interp_on=True)
        N = 40
        y0 = P.perturbed_ensemble(P.TBP_Y0, N)
        xf = sign * 0.9 * P.tbp_period()
        case.xf = xf
        g = case.run_gpu(y0, dense_mode=1)
        sc = case.oracle_scalar()
        grids = [np.array([0.37 * xf]), np.array([0.0, xf]), np.array([0.0, 0.5 * xf, xf]),
                 np.linspace(0.0, 1e-7 * xf, 9), np.linspace((1.0 - 1e-7) * xf, xf, 9), np.array([xf]), np.array([0.0])]
        for kern in ("items", "steps"):
            monkeypatch.setenv("FLINT_B200_INTERP_KERNEL", kern)
            for xarr in grids:
                xarr = xarr.copy()
                if len(xarr) > 1: xarr[-1] = xarr[-1] if abs(xarr[-1]) < abs(xf) else xf
                st, yarr, ist = g["erk"].InterpolateBatch(xarr, N, SaveInterpCoeffs=True)
                assert st == 0 and (ist == 0).all(), (kern, xarr, ist)
                for t in (0, 13, N - 1):
                    sc.integrate(0.0, y0[:, t], float(xf), io=case.oracle_int())
                    so, yo = sc.interpolate(xarr, save=True)
                    assert so == 0
                    K.assert_bit_equal(yarr[t], yo, "edge grid %s, kernel %s, trajectory %d" % (xarr[:3], kern, t))
        monkeypatch.delenv("FLINT_B200_INTERP_KERNEL")


def test_per_trajectory_grids():
    """flint_erk_interpolate_batch_grids: one grid per trajectory (ragged spans: every trajectory has its own Xf), a backward
    ensemble, and a row that leaves its trajectory's span (FLINT_ERROR_INTP_ARRAY for that trajectory only)."""
    for sign in (1.0, -1.0):
        case = K.twobody_case(F.ERK_DOP853, 1e-10, 0, nper=1, interp_on=True)
        N, G = 96, 57
        y0 = P.perturbed_ensemble(P.TBP_Y0, N)
        xf = sign * np.linspace(0.4, 1.0, N) * P.tbp_period()
        case.xf = xf
        grids = np.linspace(0.0, 1.0, G)[None, :] ** 1.5 * xf[:, None]
        grids[:, -1] = xf
        for dense_mode in (0, 1):
            g = case.run_gpu(y0, dense_mode=dense_mode)
            bad = grids.copy()
            bad[5] = grids[5] * 1.01
            st, yarr, ist = g["erk"].InterpolateBatch(bad, N, SaveInterpCoeffs=True)
            assert ist[5] == F.FLINT_ERROR_INTP_ARRAY and (np.delete(ist, 5) == 0).all()
            st, yarr, ist = g["erk"].InterpolateBatch(grids, N)
            assert st == 0 and (ist == 0).all()
            sc = case.oracle_scalar()
            for t in (0, 17, N - 1):
                sc.integrate(0.0, y0[:, t], float(xf[t]), io=case.oracle_int())
                so, yo = sc.interpolate(grids[t], save=True)
                assert so == 0
                K.assert_bit_equal(yarr[t], yo, "per-trajectory grid, trajectory %d" % t)


# ------------------------------------------------------------------------------------------------ run-time registered systems
def test_runtime_registered_systems(tmp_path, monkeypatch):
    """user F / G handed to the STOCK library as source at run time (flint_sys_register_source; NVRTC, csrc/rtc.cu).
    The run-time Lorenz functor equals the oracle's Lorenz bit for bit (all launch paths: persistent lanes, gang-scheduled rounds,
    two shards, dense output + fused Interpolate); a Van der Pol oscillator with an upward zero-crossing event agrees with scipy's
    DOP853, with mu from the system parameters and from Integrate(params=...)."""
    import os
    from scipy.integrate import solve_ivp
    monkeypatch.setenv("FLINT_B200_RTC_CACHE", str(tmp_path))
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "user_systems", "runtime_systems.cuh")).read()
    RT_LORENZ = F.register_system_source("RtLorenz", src, False)
    RT_VDP = F.register_system_source("RtVanDerPol", src, True)
    n0 = F.lib().flint_b200_kernel_count()
    for method, arith in ((F.ERK_DOP54, 0), (F.ERK_DOP853, 1), (F.ERK_VERNER98R, 0), (F.ERK_VERNER65E, 1)):
        case = K.lorenz_case(method, 1e-10, arith, xf=8.0)
        y0 = P.lorenz_ensemble(300)
        o = case.run_oracle(y0)
        sys_ = F.DiffEqSys(RT_LORENZ, F.FLINT_EVSET_NONE, list(P.LORENZ_PARAMS), n=3, m=0)
        erk = F.ERK()
        assert erk.Init(sys_, case.max_steps, Method=method, ATol=[case.atol], RTol=[case.rtol], Arith=arith) == 0
        for kw in (dict(), dict(round_len=16), dict(devices=[0, 0], shard_chunk=64)):
            g = erk.IntegrateBatch(0.0, y0, 8.0, StepSz=0.0, StiffTest=0, **kw)
            for a, b, what in ((g["Yf"], o["yf"], "Yf"), (g["counters"], o["counters"], "counters"), (g["status"], o["status"], "status")):
                K.assert_bit_equal(a, b, "run-time Lorenz %s %r %s" % (MNAME[method], kw, what))
    assert F.lib().flint_b200_kernel_count() == n0 + 4
    # dense output of a run-time system: coefficients rebuilt from the step logs by the run-time module's own dense kernel
    case = K.lorenz_case(F.ERK_DOP853, 1e-9, 0, xf=3.0, interp_on=True)
    y0 = P.lorenz_ensemble(64)
    xarr = np.linspace(0.0, 3.0, 301)
    o = case.run_oracle(y0, xarr=xarr)
    sys_ = F.DiffEqSys(RT_LORENZ, F.FLINT_EVSET_NONE, list(P.LORENZ_PARAMS), n=3, m=0)
    for dense_mode in (0, 1):
        erk = F.ERK()
        assert erk.Init(sys_, case.max_steps, Method=F.ERK_DOP853, ATol=[case.atol], RTol=[case.rtol], InterpOn=True) == 0
        g = erk.IntegrateBatch(0.0, y0, 3.0, StepSz=0.0, StiffTest=0, dense_mode=dense_mode)
        st, yarr, ist = erk.InterpolateBatch(xarr, 64)
        assert st == 0 and (ist == 0).all()
        K.assert_bit_equal(yarr, o["yarr"], "run-time Lorenz dense output, dense_mode %d" % dense_mode)
    # Van der Pol with an event, against scipy
    mu, xf, N = 1.5, 20.0, 64
    y0 = np.stack([2.0 + 1e-3 * P.uniform01(np.arange(N), 5), np.zeros(N)])
    up = lambda t, y: y[0]  # noqa: E731
    up.direction = 1
    for sp, params in (([mu], None), ([0.3], [mu])):
        sys_ = F.DiffEqSys(RT_VDP, 1, sp, n=2, m=1)
        erk = F.ERK()
        assert erk.Init(sys_, 100000, Method=F.ERK_DOP853, ATol=[1e-13], RTol=[1e-12], EventsOn=True, EventTol=[1e-11]) == 0
        g = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, EventStates=True, MaxEvents=8, params=params)
        assert (g["status"] == 0).all()
        for i in (0, 17, 63):
            r = solve_ivp(lambda t, y: [y[1], mu * (1 - y[0] ** 2) * y[1] - y[0]], (0.0, xf), y0[:, i], method="DOP853", rtol=1e-12, atol=1e-13, events=up)
            assert np.allclose(g["Yf"][:, i], r.y[:, -1], rtol=0, atol=1e-8), (g["Yf"][:, i], r.y[:, -1])
            te = r.t_events[0]
            assert g["nEvents"][i] == len(te), (g["nEvents"][i], len(te))
            assert np.allclose(g["EventStates"][0, :len(te), i], te, rtol=0, atol=1e-8)
            assert (g["EventStates"][3, :len(te), i] == 1.0).all()


# ------------------------------------------------------------------------------------------------ dense output spilled to the host
@pytest.mark.parametrize("kw", [dict(dense_spill=3), dict(dense_spill=2, devices=[0, 0]), dict(dense_spill=1, dense_mode=1),
                                dict(dense_spill=4, dense_cap=128)])
def test_dense_output_spilled_to_host_memory(kw, monkeypatch):
    """an ensemble whose dense output does not fit in device memory runs as parts, one after the other; each
    part's store waits in host memory between Integrate and Interpolate and comes back one part at a time (the reference keeps
    Xint / Bip in host memory, src/ERKInit.f90:321-332).  Same results as the oracle, bit for bit: final states, events,
    both Yarr layouts, a second Interpolate after SaveInterpCoeffs, storage gone after the last one.  dense_spill=1 picks the
    number of parts from the memory budget (here a budget of 24 MB per part stands in for a full device)."""
    if kw.get("dense_spill") == 1:
        monkeypatch.setenv("FLINT_B200_SPILL_BUDGET", str(24 << 20))
    case = K.cr3bp_case(F.ERK_VERNER98R, 1e-10, 0, norb=1.0, interp_on=True)
    N = 1500
    y0 = P.cr3bp_ensemble(N, start=11)
    xarr = np.linspace(0.0, case.xf, 257)
    xarr[-1] = case.xf
    o = case.run_oracle(y0, xarr=xarr)
    g = case.run_gpu(y0, shard_chunk=100, **kw)
    K.compare_batch(g, o, case, "spilled dense %r" % kw)
    erk = g["erk"]
    for layout in (0, 1):
        st, yarr, ist = erk.InterpolateBatch(xarr, N, SaveInterpCoeffs=True, layout=layout)
        assert st == 0 and (ist == 0).all()
        K.assert_bit_equal(yarr if layout == 0 else np.transpose(yarr, (2, 1, 0)), o["yarr"], "spilled Yarr layout %d %r" % (layout, kw))
    st, yarr, ist = erk.InterpolateBatch(xarr[:40].copy(), N)          # the last one frees the storage
    assert st == 0 and (ist == 0).all()
    K.assert_bit_equal(yarr, o["yarr"][:, :40], "spilled Yarr, last call")
    st, _, _ = erk.InterpolateBatch(xarr, N)
    assert st == F.FLINT_ERROR_INTP_OFF
