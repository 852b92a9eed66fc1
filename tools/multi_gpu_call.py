"""ONE flint_erk_integrate_batch call with host buffers sharded over the GPUs of the box by the library itself
(flint_exec.n_devices): wall time of the call against the one-device call, results compared bit for bit.
    python tools/multi_gpu_call.py [N] [n_devices ...]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
ndev = torch.cuda.device_count()
counts = [int(v) for v in sys.argv[2:]] or [d for d in (1, 2, 4, 8) if d <= ndev]
pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()  # noqa: E731
y0 = pin((6, N), torch.float64)
y0[:] = P.cr3bp_ensemble(N)
sys_, erk, ikw, xf = P.cr3bp_config3()
ref = None
for D in counts:
    outs = {"Yf": pin((6, N), torch.float64), "status": pin((N,), torch.int32), "counters": pin((4, N), torch.int32),
            "Xf": pin((N,), torch.float64), "StepSz": pin((N,), torch.float64), "StiffTest": pin((N,), torch.int32),
            "EventStates": pin((8, 2, N), torch.float64), "nEvents": pin((N,), torch.int32)}
    ts, ks = [], []
    for rep in range(4):
        t0 = time.perf_counter()
        r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, MaxEvents=2, out=outs, devices=list(range(D)) if D > 1 else None, **ikw)
        ts.append(time.perf_counter() - t0)
        ks.append(r["kernel_ms"])
    res = {k: np.array(r[k]) for k in ("Yf", "Xf", "status", "counters", "nEvents", "EventStates")}
    same = None
    if ref is None:
        ref = res
    else:
        same = all(np.array_equal(res[k], ref[k], equal_nan=True) for k in ref)
    print(json.dumps({"N": N, "devices": D, "call_ms_best": 1e3 * min(ts[1:]), "device_ms_slowest_shard": min(ks[1:]),
                      "orbits_per_s_e2e": N / min(ts[1:]), "identical_to_one_device": same}), flush=True)
