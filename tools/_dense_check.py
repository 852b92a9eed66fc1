import sys, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import flint_b200 as F
from flint_b200 import problems as P
N = 1 << 17
for arith in (0, 1, 0):
    for cap in (640, 2500 if N < (1 << 17) else 700):
        sys_, erk, ikw, xf = P.cr3bp_config3(method=F.ERK_VERNER98R, rtol=1e-11, arith=arith, interp_on=True)
        y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
        for rep in range(3):
            r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, dense_cap=cap, **ikw)
            print("arith", arith, "cap", cap, "rep", rep, "kernel ms %.2f" % r["kernel_ms"], "ok frac %.5f" % float((r["status"] == 0).float().mean()), flush=True)
        del erk
        torch.cuda.empty_cache()
