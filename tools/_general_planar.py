"""General (6-component) planar CR3BP kernels: the plane variant switched off (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P
N = 1 << 19
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
for arith in (0, 1):
    for evset, em in ((F.FLINT_EVSET_CR3BP_SAMPLE3, 0), (F.FLINT_EVSET_CR3BP_SAMPLE3, 1), (F.FLINT_EVSET_NONE, 0)):
        sys_, erk, ikw, xf = P.cr3bp_config3(arith=arith, eventset=evset)
        best = 1e30
        for rep in range(2):
            r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, plane_variant=2, event_mode=em, **ikw)
            best = min(best, r["kernel_ms"])
        print("arith %d events %d event_mode %d: %.1f ms" % (arith, evset, em, best), flush=True)
