"""Config 4 (separate path) on N orbits, one Integrate + one Interpolate, for a launch list (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flint_b200 as F
from flint_b200 import problems as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
method = {"v98r": F.ERK_VERNER98R, "dop853": F.ERK_DOP853, "dop54": F.ERK_DOP54, "v65e": F.ERK_VERNER65E}[sys.argv[3] if len(sys.argv) > 3 else "v98r"]
G = 10000
sys_, erk, ikw, xf = P.cr3bp_config3(method=method, rtol=1e-11, arith=0, interp_on=True)
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
for _ in range(2):
    r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, dense_cap=640, dense_mode=mode, **ikw)
    print("integrate ms", r["kernel_ms"])
xarr = torch.linspace(0.0, xf, G, dtype=torch.float64).cuda(); xarr[-1] = xf
yarr = torch.empty((N, G, 6), dtype=torch.float64, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st, _, ist = erk.InterpolateBatch(xarr, N, SaveInterpCoeffs=True, layout=0, out=yarr); e1.record(); torch.cuda.synchronize()
    print("interpolate ms", e0.elapsed_time(e1))
