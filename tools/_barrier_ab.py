"""Per-attempt block barrier on/off for the smaller hot loops (development aid): run with FLINT_B200_LIB set."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P
import cases as K
runs = [("Lorenz DOP54 1e-11 x<=25", K.lorenz_case(F.ERK_DOP54, 1e-11, 0, xf=25.0), P.lorenz_ensemble(1 << 18)),
        ("two-body DOP853 1e-10", K.twobody_case(F.ERK_DOP853, 1e-10, 0), P.perturbed_ensemble(P.TBP_Y0, 1 << 17)),
        ("halo DOP853 1e-12", K.halo_case(F.ERK_DOP853, 1e-12), P.perturbed_ensemble(P.HALO_Y0, 1 << 18)),
        ("halo Verner98R 1e-12", K.halo_case(F.ERK_VERNER98R, 1e-12), P.perturbed_ensemble(P.HALO_Y0, 1 << 18))]
for name, case, y0 in runs:
    sys_, erk, st = case.make_erk()
    yd = torch.from_numpy(y0).cuda()
    best = 1e9
    for _ in range(3):
        r = erk.IntegrateBatch(case.x0, yd, case.xf, StepSz=0.0, **case.int_kwargs())
        best = min(best, r["kernel_ms"])
    print("%-28s %8.2f ms" % (name, best), flush=True)
