"""Config-4 dense output on N orbits for ncu: Integrate (dense_mode from argv), then Interpolate layout 0 (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flint_b200 as F
from flint_b200 import problems as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
arith = int(sys.argv[3]) if len(sys.argv) > 3 else 0
layout = int(sys.argv[4]) if len(sys.argv) > 4 else 0
G = 10000
sys_, erk, ikw, xf = P.cr3bp_config3(method=F.ERK_VERNER98R, rtol=1e-11, arith=arith, interp_on=True)
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, dense_cap=1024, dense_mode=mode, **ikw)
print("integrate ms", r["kernel_ms"])
xarr = torch.linspace(0.0, xf, G, dtype=torch.float64).cuda(); xarr[-1] = xf
yarr = torch.empty((N, G, 6) if layout == 0 else (6, G, N), dtype=torch.float64, device="cuda")
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st, _, ist = erk.InterpolateBatch(xarr, N, SaveInterpCoeffs=True, layout=layout, out=yarr); e1.record(); torch.cuda.synchronize()
    print("interpolate ms", e0.elapsed_time(e1), "ok", float((ist == 0).float().mean()))
