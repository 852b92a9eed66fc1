#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native FLINT ensemble integrator.

Metric (BASELINE.json): CR3BP DOP853 orbits/sec for THE 2^20-orbit ensemble, rtol 1e-12, atol 1e-15,
test_CR3BP event set, at 1/2/4/8 GPUs: STRONG scaling -- the same 1 M orbits
are split over the ranks by the library's block-cyclic partition (flint_b200_shard_*: chunks of 4096
trajectories, chunk c -> rank c mod N).  `--scaling weak` keeps 2^20 orbits per GPU instead.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--arith strict|fma] [--scaling strong|weak]

A "step" is one pass of the hot path (ERK_class%Integrate for every orbit of the ensemble)
over one batch of seeded synthetic initial conditions.
  value  : orbits/s with inputs resident in HBM, timed with CUDA events on the launch stream
  e2e    : the same through the C-ABI call with pinned HOST buffers (H2D + kernel + D2H timed)
  roofline: algorithmic fp64 flops (SURVEY.md §8d formula, from the per-trajectory step counters)
           / kernel time, against the fp64 DFMA peak measured in this run by a register-resident
           microbenchmark (MEASURED_PEAKS.json carries no fp64 figure)
  cpu_baseline: the CPU oracle (oracle/, a C++ restatement of the reference; the Fortran reference
           cannot be compiled here -- no Fortran compiler) on all host cores, on a bounded sample.
N > 1 is launched by torch.distributed.run (one rank per GPU); there is no collective on the data
path -- NCCL is used only for the barrier and the max-over-ranks of the timings.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_ENSEMBLE = 1 << 20         # the metric's ensemble
RTOL = 1e-12
MAX_EVENTS = 2
CPU_SAMPLE = 1 << 18          # cpu_baseline leg: ~10 s of CPU work on 16 cores
REF_STEP_SAMPLE = int(os.environ.get("FLINT_BENCH_REF_SAMPLE", 1 << 16))     # --impl reference: orbits per step (the variable is for the CPU test suite)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arith", default=os.environ.get("FLINT_BENCH_ARITH", "strict"), choices=["strict", "fma"])
    ap.add_argument("--orbits", type=int, default=N_ENSEMBLE, help="orbits of the ensemble (strong: in total; weak: per GPU); default: the metric's 2^20")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, the metric): the one ensemble is split over the GPUs; weak: every GPU gets an ensemble of that size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


def workload_name(n, per_gpu=False):
    return ("CR3BP planar Arenstorf ensemble, %d orbits%s, y0(1)*(1+1e-6*u), 2.5 periods, DOP853 rtol 1e-12 atol 1e-15, "
            "events m=3 mask [F,T,F] EventStepSz [1e-5,0,0] ETol [1e-3,1e-9,1e-9] (tests/test_CR3BP.f90 shape; SURVEY.md 8d config 3a)"
            % (n, " per GPU" if per_gpu else " in total"))


# --------------------------------------------------------------------------- CPU oracle legs
def run_oracle_sample(arith, n_sample, start=0, nthreads=0):
    """Integrate n_sample orbits of the SAME seeded ensemble with the CPU oracle; returns (orbits/s, threads, seconds)."""
    import cases as K
    from flint_b200 import problems as P
    import flint_b200 as F
    case = K.cr3bp_case(F.ERK_DOP853, RTOL, arith, max_events=MAX_EVENTS)
    y0 = P.cr3bp_ensemble(n_sample, start=start)
    if nthreads <= 0:
        # every host core this process may use -- torch.distributed.run exports OMP_NUM_THREADS=1, which would
        # silently turn the CPU arm into a single-thread run
        nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    t0 = time.perf_counter()
    o = case.run_oracle(y0, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return n_sample / dt, int(o["threads"]), dt, o


def reference_kind():
    """Which CPU implementation this box can run: the real reference (oracle/_ref/flint_ref_driver, gfortran -O3 -fopenmp build of
    /root/reference/src + oracle/ref_driver.f90, made by `make -C oracle ref` where a Fortran compiler and the sources exist) or
    the C++ port (oracle/liboracle.so).  The probe result is part of the line."""
    import shutil
    exe = os.path.join(ROOT, "oracle", "_ref", "flint_ref_driver")
    probe = {"gfortran": shutil.which("gfortran"), "flang": shutil.which("flang"), "ifx": shutil.which("ifx"),
             "reference_sources": os.path.isdir("/root/reference/src"), "oracle/_ref/flint_ref_driver": os.path.exists(exe)}
    return ("reference" if os.path.exists(exe) else "port"), exe, probe


def run_reference_driver(exe, n_sample, start, nthreads):
    """oracle/_ref/flint_ref_driver N START THREADS -> 'orbits seconds threads' on stdout (oracle/ref_driver.f90)."""
    env = dict(os.environ, OMP_NUM_THREADS=str(nthreads))
    t0 = time.perf_counter()
    out = subprocess.run([exe, str(n_sample), str(start), str(nthreads)], capture_output=True, text=True, env=env, check=True).stdout
    dt = time.perf_counter() - t0
    f = out.split()
    return (float(f[1]) if len(f) > 1 else dt), (int(f[2]) if len(f) > 2 else nthreads)


def main_reference(a, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host cores: FLINT itself (gfortran -O3,
    OpenMP over trajectories) where oracle/_ref holds its build, else the C++ port of oracle/ (this image has no Fortran
    compiler, so here it is always the port; the probe is reported in the line)."""
    if rank != 0:
        return
    arith = 1 if a.arith == "fma" else 0
    kind, exe, probe = reference_kind()
    if kind == "reference":
        nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        run_reference_driver(exe, 2048, 0, nthreads)
        tot_t, tot_n, threads = 0.0, 0, nthreads
        for s_ in range(a.steps):
            dt, threads = run_reference_driver(exe, REF_STEP_SAMPLE, s_ * REF_STEP_SAMPLE, nthreads)
            tot_t += dt
            tot_n += REF_STEP_SAMPLE
        val = tot_n / tot_t
        emit_line(reference_line(a, val, tot_t, threads, "reference", probe))
        return
    for _ in range(max(a.warmup, 0)):
        run_oracle_sample(arith, 2048)
    tot_t, tot_n, threads = 0.0, 0, 1
    for s in range(a.steps):
        _, threads, dt, _ = run_oracle_sample(arith, REF_STEP_SAMPLE, start=s * REF_STEP_SAMPLE)
        tot_t += dt
        tot_n += REF_STEP_SAMPLE
    val = tot_n / tot_t
    emit_line(reference_line(a, val, tot_t, threads, "port", probe))


def reference_line(a, val, tot_t, threads, kind, probe):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return {"impl": "reference", "metric": "CR3BP DOP853 orbits/sec", "value": val, "unit": "orbits/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot_t / a.steps, "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a.orbits, a.scaling == "weak"), "sample": "%d orbits per step of the same seeded ensemble" % REF_STEP_SAMPLE,
                       "arith": a.arith, "gpus_used": 0, "fortran_probe": probe},
            "cpu_baseline": {"value": val, "unit": "orbits/s", "cores": threads, "kind": kind,
                             "sample": "%d steps x %d orbits, OpenMP schedule(dynamic,64), one ERK object per thread" % (a.steps, REF_STEP_SAMPLE)},
            "e2e": {"value": val, "unit": "orbits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# --------------------------------------------------------------------------- GPU arm
def main_ours(a, rank, world, local_rank):
    import numpy as np
    import torch
    import flint_b200 as F
    from flint_b200 import problems as P

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    arith = F.FLINT_ARITH_FMA if a.arith == "fma" else F.FLINT_ARITH_STRICT
    sys_, erk, ikw, xf = P.cr3bp_config3(method=F.ERK_DOP853, rtol=RTOL, arith=arith)
    if a.scaling == "strong":
        # THE ensemble of a.orbits trajectories, dealt over the ranks by the library's own partition (chunks of 4096, chunk c ->
        # rank c mod world: flint_b200_shard_indices): shards, no exchange
        total_orbits = a.orbits
        index = F.shard_indices(total_orbits, world, rank)
    else:
        # every rank its own a.orbits trajectories of the seeded stream
        total_orbits = a.orbits * world
        index = np.arange(rank * a.orbits, (rank + 1) * a.orbits, dtype=np.int64)
    n = len(index)
    y0_host = P.cr3bp_ensemble(0, index=index)
    y0 = torch.from_numpy(y0_host).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    peak_tf, _ = F.measure_fp64_peak(local_rank)

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step():
        r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, MaxEvents=MAX_EVENTS, **ikw)
        return r

    for _ in range(max(a.warmup, 3)):
        flush.zero_()
        r = one_step()
    # ---------------- timed region: device-resident inputs ----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    launches0 = F.launch_count()
    t_wall0 = time.time()
    ker_ms = []
    for _ in range(a.steps):
        flush.zero_()                      # L2 flush between timed iterations (not inside the event-timed kernel window)
        r = one_step()                     # synchronous; kernel_ms = CUDA events around the launch on its stream
        ker_ms.append(r["kernel_ms"])
    barrier()
    t_wall1 = time.time()
    launches = F.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    dev_ms = float(sum(ker_ms))
    # per-trajectory counters of the last step -> algorithmic flops actually required
    cnt = r["counters"].cpu().numpy().astype(np.int64)
    nev = int(r["nEvents"].sum().item())
    status = r["status"].cpu().numpy()
    flops_step = P.algorithmic_flops(F.ERK_DOP853, F.FLINT_SYS_CR3BP_PLANAR, 6, int(cnt[0].sum()) + nev, int(cnt[1].sum()) + nev, nev)
    flops4 = P.algorithmic_flops(F.ERK_DOP853, F.FLINT_SYS_CR3BP_PLANAR, 4, int(cnt[0].sum()) + nev, int(cnt[1].sum()) + nev, nev)
    # ---------------- e2e: host buffers through the C ABI ----------------
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()  # noqa: E731
    h_y0 = pin((6, n), torch.float64)
    h_y0[:] = y0_host
    outs = {"Yf": pin((6, n), torch.float64), "status": pin((n,), torch.int32), "counters": pin((4, n), torch.int32),
            "Xf": pin((n,), torch.float64), "StepSz": pin((n,), torch.float64), "StiffTest": pin((n,), torch.int32),
            "EventStates": pin((8, MAX_EVENTS, n), torch.float64), "nEvents": pin((n,), torch.int32)}
    e2e_t = []
    rh = erk.IntegrateBatch(0.0, h_y0, xf, StepSz=0.0, MaxEvents=MAX_EVENTS, out=outs, device=local_rank, **ikw)   # warm
    barrier()
    for _ in range(a.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        rh = erk.IntegrateBatch(0.0, h_y0, xf, StepSz=0.0, MaxEvents=MAX_EVENTS, out=outs, device=local_rank, **ikw)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    e2e_s = float(sum(e2e_t))
    nt = total_orbits                                                                   # whole job: every rank copies its own shard
    h2d = 6 * nt * 8 + nt * 8 + nt * 8 + nt * 4                                       # y0, xf, stepsz, stifftest
    d2h = nt * 8 + 6 * nt * 8 + nt * 4 + nt * 8 + 4 * nt * 4 + nt * 4 + nt * 4 + (6 + 2) * MAX_EVENTS * nt * 8   # xf,yf,status,stepsz,counters,stiff,evcount,events
    same = bool(np.array_equal(rh["Yf"], r["Yf"].cpu().numpy()))

    # ---------------- max over ranks ----------------
    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device=dev)
    agg = torch.tensor([flops_step, float((status == 0).sum())], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max = float(t[0]), float(t[1])
    value = total_orbits * a.steps / (dev_ms_max * 1e-3)
    e2e_value = total_orbits * a.steps / e2e_s_max
    achieved_tf = flops_step / (dev_ms / a.steps * 1e-3) / 1e12            # this rank's kernel (the dominant and only kernel of a step)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch_%s" % a.arith)
        except Exception:
            traffic = None

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, thr, dt, o = run_oracle_sample(arith, CPU_SAMPLE)
        # the oracle is also the checker here: the sample is the first CPU_SAMPLE orbits of the timed GPU step
        k = min(CPU_SAMPLE, n)                   # world == 1: this rank's trajectories are the ensemble's, in order
        parity = bool(np.array_equal(o["yf"][:, :k], r["Yf"][:, :k].cpu().numpy()) and
                      np.array_equal(o["counters"][:, :k], r["counters"][:, :k].cpu().numpy()))
        kind, exe, probe = reference_kind()
        ref_leg = None
        if kind == "reference":                  # FLINT itself, timed beside the port on the same sample
            rdt, rthr = run_reference_driver(exe, CPU_SAMPLE, 0, thr)
            ref_leg = {"value": CPU_SAMPLE / rdt, "unit": "orbits/s", "cores": rthr, "kind": "reference", "sample": "the same %d orbits" % CPU_SAMPLE}
        cpu = {"value": v, "unit": "orbits/s", "cores": thr, "kind": "port", "fortran_probe": probe, "reference_leg": ref_leg,
               "sample": "first %d orbits of the same seeded ensemble, %.1f s, OpenMP dynamic; bit-parity with the GPU step: %s" % (CPU_SAMPLE, dt, parity)}
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "CR3BP DOP853 orbits/sec", "value": value, "unit": "orbits/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.orbits, a.scaling == "weak"), "orbits_total": total_orbits, "orbits_per_gpu": n,
                   "parallelism": "shard%d: block-cyclic chunks of %d orbits, independent, no collective" % (world, F.FLINT_SHARD_CHUNK),
                   "arith": a.arith, "l2": "256 MB flush between timed iterations", "timing": "CUDA events on the launch stream, max over ranks",
                   "success_fraction": float(agg[1]) / total_orbits, "host_and_device_results_identical": same},
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "traffic": traffic, "kernel": "erk_step_kernel<DOP853,CR3BPPlanar,%s,park,no-dense> (2 passes; + erk_init_kernel, erk_event_kernel < 1 %% of the step)" % a.arith,
                     "peak_source": "DFMA microbenchmark measured in this run (flint_b200_measure_fp64_peak); MEASURED_PEAKS.json has no fp64 entry",
                     "algorithmic_flop_per_orbit": flops_step / n,
                     # the plane-variant kernel proves two of the six components identically zero and skips their sums: the flops
                     # it EXECUTES are the n = 4 count (the reference algorithm, and `achieved` above, count n = 6)
                     "frac_executed_n4": flops4 / (dev_ms / a.steps * 1e-3) / 1e12 / peak_tf},
        "e2e": {"value": e2e_value, "unit": "orbits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "wall_s_timed_region": t_wall1 - t_wall0,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    emit_line(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print on stdout while the bench runs (e.g. NCCL's version banner) goes to stderr: stdout carries the
    one JSON line and nothing else."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit_line(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if a.impl == "ours" and a.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` by hand: one rank per GPU needs the launcher -- start it instead of silently running one GPU
        if _REAL_STDOUT is not None:
            os.dup2(_REAL_STDOUT, 1)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
                                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
                                   os.path.abspath(__file__)] + sys.argv[1:])
    if a.impl == "reference":
        main_reference(a, rank, world)
    else:
        main_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
