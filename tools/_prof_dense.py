import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flint_b200 as F
from flint_b200 import problems as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
only = sys.argv[2] if len(sys.argv) > 2 else ""
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
for method, name in ((F.ERK_VERNER98R, "V98R"), (F.ERK_DOP853, "DOP853"), (F.ERK_DOP54, "DOP54"), (F.ERK_VERNER65E, "V65E")):
    for dense in (False, True):
        if only and only != "%s%d" % (name, dense): continue
        sys_, erk, ikw, xf = P.cr3bp_config3(method=method, rtol=1e-11, arith=0, interp_on=dense)
        best = 1e9
        for _ in range(2):
            r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, dense_cap=2500 if dense else 0, **ikw)
            best = min(best, r["kernel_ms"])
        cnt = r["counters"].cpu().numpy().astype(np.int64)
        print("%-7s dense=%d: %8.2f ms  %.3f Morbit/s  attempts/orbit %.0f accepted %.0f" % (name, dense, best, N / best / 1e3, cnt[0].mean(), cnt[1].mean()), flush=True)
        del erk
