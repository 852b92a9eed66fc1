#!/usr/bin/env python3
"""profiles/r01_configs.jsonl (tools/bench_configs.py output) -> profiles/r01_configs.md"""
import json, sys
src, dst = sys.argv[1], sys.argv[2]
rows = [json.loads(l) for l in open(src) if l.strip()]
an = {0: "strict", 1: "FMA"}
L = ["# Round 1: BASELINE.json configs 2, 4, 5 on one B200 (`tools/bench_configs.py`, inputs in HBM, CUDA-event timing, best of 2)\n",
     "fp64 fractions are algorithmic flops (SURVEY.md §8d formulas from the per-trajectory counters) / the DFMA peak measured in the same "
     "run; HBM fractions are bytes / MEASURED_PEAKS.json `hbm_gbs` (6548.8 GB/s).\n",
     "## Config 2 — Lorenz, 2^20 perturbed initial conditions, DOP54, rtol 1e-11 / atol 1e-14, x in [0,100], no events\n",
     "| arithmetic | ms | trajectories/s | attempts/trajectory | TFLOP/s alg | of fp64 peak |", "|---|---:|---:|---:|---:|---:|"]
for r in rows:
    if r["config"] == 2:
        L.append("| %s | %.0f | %.0f | %.0f | %.2f | %.1f %% |" % (an[r["arith"]], r["ms"], r["value"], r["attempts_per_traj"], r["tflops_alg"], 100 * r["frac_fp64_peak"]))
L.append("\nPer attempt the three-component DOP54 step is ~240 algorithmic flops, while the error norm's divisions, `sqrt(sm/n)`, the "
         "controller's `pow` (~85 fp64 instructions) and its two divisions are ~170 fp64 instructions counted as a handful of flops: in strict "
         "arithmetic a fully busy pipe would read ~30 %.\n")
d = [r for r in rows if r["config"] == 4]
if d:
    L.append("## Config 4 — planar CR3BP, Verner98R, InterpOn (all 6 states), rtol 1e-11, 2^18 orbits, 10 000-point shared grid\n")
    L.append("2^18 orbits on ONE GPU as two shards of 2^17 run back to back (`Yarr` 62.9 GB + dense store 37 GB at 640 slots per shard; the whole "
             "`Yarr` is 125.8 GB); times are sums over both shards. Interpolation bytes: `Yarr` written (N·G·nI·8) and, second figure, plus "
             "every step record read once (SURVEY.md §8d counts both when the coefficients are materialised). %.2f %% of the orbits finish "
             "the integration (the rest give up near a collision, as in the oracle); %.2f %% interpolate (the others have more accepted "
             "steps than the 640 slots this run gives the store and report FLINT_ERROR_INTP_ARRAY).\n" % (
                 100 * d[1]["integrated_fraction"], 100 * d[1]["interpolated_fraction"]))
    L += ["| phase | arithmetic | ms | rate | roofline |", "|---|---|---:|---:|---:|"]
    for r in d:
        if "integrate" in r["what"]:
            L.append("| integrate + dense coefficients (init, step passes, event, dense kernels) | %s | %.1f | %.2f M orbits/s | %.1f %% of fp64 peak |" % (
                an[r["arith"]], r["ms"], r["value"] / 1e6, 100 * r["frac_fp64_peak"]))
        else:
            lay = "`[N][G][nI]` (Fortran `Yarr(nI,G)` per orbit)" if r["what"].endswith("0") else "`[nI][G][N]` (SoA)"
            L.append("| Interpolate, %s | %s | %.1f | %.2f TB/s written, %.2f TB/s with the records; %.1f G points/s | %.1f %% / %.1f %% of HBM peak |" % (
                lay, an[r["arith"]], r["ms"], r["gbs_written"] / 1e3, r["gbs_total"] / 1e3, r["points_per_s"] / 1e9,
                100 * r["frac_hbm_peak_written"], 100 * r["frac_hbm_peak"]))
    L.append("\nThe `[N][G][nI]` kernel is bound by the L1 data pipe (78 %: shared-memory wavefronts of the coefficient reads), not by HBM; see "
             "`profiles/r01_notes.md`.\n")
w = [r for r in rows if r["config"] == 5]
if w:
    L.append("## Config 5 — work-precision sweeps, 2^16 trajectories, strict arithmetic (ms per ensemble / trajectories per second / mean FCalls / fraction of fp64 peak)\n")
    for sysname, title in (("halo", "spatial CR3BP halo orbit (tests/test_WP.f90), atol 1e-80"), ("twobody", "two-body Molniya orbit, 10 periods (tests/test_TBP.f90), atol = rtol*1e-3")):
        L += ["### %s\n" % title, "| method | rtol 1e-6 | 1e-8 | 1e-10 | 1e-12 | 1e-14 |", "|---|---|---|---|---|---|"]
        for r in w:
            if r["system"] == sysname:
                cells = ["%.2f ms / %.1f M/s / %.0f / %.0f %%" % (c["ms"], c["traj_per_s"] / 1e6, c["fcalls"], 100 * c["frac_fp64_peak"])
                         for k, c in sorted(r["sweep"].items(), key=lambda kv: int(kv[0][3:]))]
                L.append("| %s | %s |" % (r["method"], " | ".join(cells)))
        L.append("")
open(dst, "w").write("\n".join(L) + "\n")
