#!/usr/bin/env python3
"""Secondary measurements on one GPU: BASELINE.json configs 2, 4 and 5.  Prints one JSON line per
measurement (inputs resident in HBM, CUDA-event timing through the library's own kernel timer, best of `reps`).

    python tools/bench_configs.py [lorenz] [dense] [wp] [--small]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P

small = "--small" in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["lorenz", "dense", "wp"]
peak_tf, _ = F.measure_fp64_peak(0)
HBM_GBS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6548.8


def emit(**kw):
    print(json.dumps(kw), flush=True)


def run(erk, x0, y0, xf, reps=2, **kw):
    best, r = 1e30, None
    for _ in range(reps):
        r = erk.IntegrateBatch(x0, y0, xf, StepSz=0.0, **kw)
        best = min(best, r["kernel_ms"])
    return best, r


if "lorenz" in which:      # config 2: Lorenz, 2^20 perturbed ICs, DOP54, rtol 1e-11 atol 1e-14, x in [0,100], no events
    N = 1 << (16 if small else 20)
    for arith in (0, 1):
        sys_ = F.DiffEqSys(F.FLINT_SYS_LORENZ, F.FLINT_EVSET_NONE, list(P.LORENZ_PARAMS))
        erk = F.ERK()
        assert erk.Init(sys_, 100000, Method=F.ERK_DOP54, ATol=[1e-14], RTol=[1e-11], Arith=arith) == 0
        y0 = torch.from_numpy(P.lorenz_ensemble(N)).cuda()
        ms, r = run(erk, 0.0, y0, 100.0, StiffTest=0)
        cnt = r["counters"].cpu().numpy().astype(np.int64)
        fl = P.algorithmic_flops(F.ERK_DOP54, F.FLINT_SYS_LORENZ, 3, cnt[0].sum(), cnt[1].sum())
        emit(config=2, what="Lorenz DOP54 trajectories/s", arith=arith, N=N, ms=ms, value=N / ms * 1e3, unit="trajectories/s",
             attempts_per_traj=float(cnt[0].mean()), tflops_alg=fl / ms / 1e9, frac_fp64_peak=fl / ms / 1e9 / peak_tf,
             ok=float((r["status"] == 0).float().mean()))

if "dense" in which:       # config 4: CR3BP Verner98R, dense output, interpolation onto a shared uniform grid
    # Reported three ways: the phases SEPARATELY (Integrate builds coefficient records, dense_mode 1; Interpolate evaluates them,
    # both layouts), FUSED (Integrate keeps step logs; Interpolate rebuilds the coefficients in registers and evaluates from
    # there, Bip never in memory), and the fused path on all 2^18 orbits as ONE pair of calls (Yarr 125.8 GB + 29.5 GB of logs).
    # The separate phases need Yarr + 37 GB of records per 2^17 orbits, so they run as two shards back to back.
    import gc
    SH = 1 << (14 if small else 17)
    NSH = 2
    G = 10000
    CAP = 640

    def timed_interp(erk, xarr, n, layout, yarr):
        best, ist = 1e30, None
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st, _, ist = erk.InterpolateBatch(xarr, n, SaveInterpCoeffs=True, layout=layout, out=yarr)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, ist

    for arith in (0, 1):
        # "logs": Integrate keeps step logs (dense_mode 0); Interpolate goes through a temporary coefficient store in parts of
        # trajectories (staged, the default) or rebuilds the coefficients in registers and evaluates in the same kernel (fused)
        for mode, mname in ((1, "separate"), (0, "logs, staged"), (0, "logs, fused")):
            os.environ["FLINT_B200_DENSE_EVAL"] = "fused" if mname.endswith("fused") else "staged"
            t_int, t_itp, fl_tot, acc_tot, n_ok, n_int_ok = 0.0, [0.0, 0.0], 0.0, 0, 0, 0
            for sh in range(NSH):
                sys_, erk, ikw, xf = P.cr3bp_config3(method=F.ERK_VERNER98R, rtol=1e-11, arith=arith, interp_on=True)
                y0 = torch.from_numpy(P.cr3bp_ensemble(SH, start=sh * SH)).cuda()
                ms, r = run(erk, 0.0, y0, xf, reps=2, dense_cap=CAP, dense_mode=mode, **ikw)      # best of 2: the first call of a process also loads the kernels
                cnt = r["counters"].cpu().numpy().astype(np.int64)
                fl_tot += P.algorithmic_flops(F.ERK_VERNER98R, F.FLINT_SYS_CR3BP_PLANAR, 6, cnt[0].sum(), cnt[1].sum(), cnt[1].sum())
                acc_tot += int(cnt[1].sum())
                t_int += ms
                n_int_ok += int((r["status"] == 0).sum())
                xarr = torch.linspace(0.0, xf, G, dtype=torch.float64).cuda()
                xarr[-1] = xf
                for layout in ((0, 1) if mode == 1 else (0,)):
                    yarr = torch.empty((SH, G, 6) if layout == 0 else (6, G, SH), dtype=torch.float64, device="cuda")
                    best, ist = timed_interp(erk, xarr, SH, layout, yarr)
                    t_itp[layout] += best
                    if layout == 0:
                        n_ok += int((ist == 0).sum())
                    del yarr
                del erk, y0, xarr
                gc.collect(); torch.cuda.empty_cache()
            N = SH * NSH
            gb_w = N * G * 6 * 8 / 1e9
            gb_r = acc_tot * (54 + 2) * 8 / 1e9          # separate: every coefficient record (X0, X1, 9 x 6 coefficients) read once
            gb_log = acc_tot * 14 * 8 / 1e9              # fused: every step log (X, h, Y, F0) read once
            emit(config=4, mode=mname, what="Integrate phase (%s)" % ("steps + coefficient records" if mode == 1 else "steps, logs only"), arith=arith,
                 N=N, shards=NSH, ms=t_int, value=N / t_int * 1e3, unit="orbits/s", accepted_per_orbit=acc_tot / N, tflops_alg=fl_tot / t_int / 1e9,
                 frac_fp64_peak=fl_tot / t_int / 1e9 / peak_tf, integrated_fraction=n_int_ok / N)
            for layout in ((0, 1) if mode == 1 else (0,)):
                rd = gb_r if mode == 1 else gb_log
                emit(config=4, mode=mname, what="Interpolate onto %d-point grid, layout %d%s" % (G, layout, "" if mode == 1 else (" (coefficients rebuilt in registers)" if mname.endswith("fused") else " (from step logs, through a temporary coefficient store)")),
                     arith=arith, N=N, shards=NSH, ms=t_itp[layout], yarr_gb=gb_w, read_gb=rd, gbs_written=gb_w / t_itp[layout] * 1e3,
                     gbs_total=(gb_w + rd) / t_itp[layout] * 1e3, frac_hbm_peak_written=gb_w / t_itp[layout] * 1e3 / HBM_GBS,
                     frac_hbm_peak=(gb_w + rd) / t_itp[layout] * 1e3 / HBM_GBS, points_per_s=N * G / t_itp[layout] * 1e3,
                     interpolated_fraction=n_ok / N)
            emit(config=4, mode=mname, what="Integrate + Interpolate [N][G][nI], total", arith=arith, N=N, ms=t_int + t_itp[0],
                 orbits_per_s=N / (t_int + t_itp[0]) * 1e3)
    if not small:          # the whole of config 4 as one Integrate and one Interpolate call on one GPU (step logs: they fit)
        os.environ["FLINT_B200_DENSE_EVAL"] = "staged"
        try:
            N = 1 << 18
            sys_, erk, ikw, xf = P.cr3bp_config3(method=F.ERK_VERNER98R, rtol=1e-11, arith=0, interp_on=True)
            y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
            ms, r = run(erk, 0.0, y0, xf, reps=1, dense_cap=CAP, dense_mode=0, **ikw)
            xarr = torch.linspace(0.0, xf, G, dtype=torch.float64).cuda()
            xarr[-1] = xf
            yarr = torch.empty((N, G, 6), dtype=torch.float64, device="cuda")
            best, ist = timed_interp(erk, xarr, N, 0, yarr)
            emit(config=4, mode="logs, staged, one call", what="2^18 orbits: one Integrate + one Interpolate call", arith=0, N=N, integrate_ms=ms, interpolate_ms=best,
                 interpolated_fraction=float((ist == 0).float().mean()), yarr_gb=N * G * 48 / 1e9,
                 device_mem_gb=torch.cuda.mem_get_info()[1] / 1e9, orbits_per_s=N / (ms + best) * 1e3)
            del yarr, y0, erk
            gc.collect(); torch.cuda.empty_cache()
        except Exception as e:      # noqa: BLE001
            emit(config=4, mode="logs, staged, one call", error=str(e)[:300])

if "wp" in which:          # config 5: work-precision sweep, spatial CR3BP halo and two-body, all four methods
    N = 1 << (12 if small else 16)
    for sysname in ("halo", "twobody"):
        for method, mname in ((F.ERK_DOP54, "DOP54"), (F.ERK_DOP853, "DOP853"), (F.ERK_VERNER65E, "Verner65E"), (F.ERK_VERNER98R, "Verner98R")):
            row = {}
            for i in range(6, 15, 2):
                rtol = P.wp_rtol(i)
                if sysname == "halo":
                    sys_ = F.DiffEqSys(F.FLINT_SYS_CR3BP_SPATIAL, F.FLINT_EVSET_NONE, [P.HALO_MU])
                    y0h, xf, atol, sid = P.perturbed_ensemble(P.HALO_Y0, N), P.HALO_PERIOD, 1e-80, F.FLINT_SYS_CR3BP_SPATIAL
                else:
                    sys_ = F.DiffEqSys(F.FLINT_SYS_TWOBODY, F.FLINT_EVSET_NONE, [P.TBP_GM])
                    y0h, xf, atol, sid = P.perturbed_ensemble(P.TBP_Y0, N), P.tbp_period() * 10, rtol * 1e-3, F.FLINT_SYS_TWOBODY
                erk = F.ERK()
                assert erk.Init(sys_, 99000, Method=method, ATol=[atol], RTol=[rtol], Arith=0) == 0
                ms, r = run(erk, 0.0, torch.from_numpy(y0h).cuda(), xf, StiffTest=1)
                cnt = r["counters"].cpu().numpy().astype(np.int64)
                fl = P.algorithmic_flops(method, sid, 6, cnt[0].sum(), cnt[1].sum())
                row["1e-%d" % i] = dict(ms=round(ms, 3), traj_per_s=round(N / ms * 1e3), fcalls=float(cnt[3].mean()),
                                        frac_fp64_peak=round(fl / ms / 1e9 / peak_tf, 4))
            emit(config=5, system=sysname, method=mname, N=N, arith="strict", sweep=row)
