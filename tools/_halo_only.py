import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P
import cases as K
case = K.halo_case(F.ERK_DOP853, 1e-12); y0 = P.perturbed_ensemble(P.HALO_Y0, 1 << 18)
sys_, erk, st = case.make_erk()
yd = torch.from_numpy(y0).cuda()
best = 1e9
for _ in range(3):
    r = erk.IntegrateBatch(case.x0, yd, case.xf, StepSz=0.0, **case.int_kwargs()); best = min(best, r["kernel_ms"])
print("halo DOP853 %.2f ms" % best)
