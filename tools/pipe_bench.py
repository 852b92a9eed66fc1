"""Host-buffer Integrate with and without the chunk pipeline (flint_exec.pipeline) on a copy-heavy configuration:
spatial CR3BP halo orbit, DOP54, rtol 1e-6, 2^20 trajectories, pinned host buffers (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P
import cases as K

N = 1 << 20
case = K.halo_case(F.ERK_DOP54, 1e-6)
y0 = P.perturbed_ensemble(P.HALO_Y0, N)
pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()  # noqa: E731
h_y0 = pin((6, N), torch.float64); h_y0[:] = y0
outs = {"Yf": pin((6, N), torch.float64), "status": pin((N,), torch.int32), "counters": pin((4, N), torch.int32),
        "Xf": pin((N,), torch.float64), "StepSz": pin((N,), torch.float64), "StiffTest": pin((N,), torch.int32)}
sys_, erk, st = case.make_erk()
ref = None
for k in (1, 2, 4, 8):
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = erk.IntegrateBatch(case.x0, h_y0, case.xf, StepSz=0.0, out=outs, pipeline=k, **case.int_kwargs())
        best = min(best, time.perf_counter() - t0)
    yf = r["Yf"].copy()
    if ref is None:
        ref = yf
    print("pipeline=%d: %.2f ms per 2^20 trajectories (host buffers), results %s" % (k, best * 1e3, "same" if np.array_equal(yf, ref) else "DIFFERENT"), flush=True)
