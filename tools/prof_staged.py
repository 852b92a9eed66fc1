"""Config-4 Interpolate [N][G][nI] from step logs: staged evaluation against part size, and the fused kernel (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flint_b200 as F
from flint_b200 import problems as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
arith = int(sys.argv[2]) if len(sys.argv) > 2 else 0
G = 10000
sys_, erk, ikw, xf = P.cr3bp_config3(method=F.ERK_VERNER98R, rtol=1e-11, arith=arith, interp_on=True)
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, dense_cap=640, dense_mode=0, **ikw)
print("integrate ms", r["kernel_ms"])
xarr = torch.linspace(0.0, xf, G, dtype=torch.float64).cuda(); xarr[-1] = xf
yarr = torch.empty((N, G, 6), dtype=torch.float64, device="cuda")
for how, part in (("fused", 0), ("staged", 512), ("staged", 2048), ("staged", 8192), ("staged", 32768), ("staged", N)):
    os.environ["FLINT_B200_DENSE_EVAL"] = how
    os.environ["FLINT_B200_DENSE_PART"] = str(part)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st, _, ist = erk.InterpolateBatch(xarr, N, SaveInterpCoeffs=True, layout=0, out=yarr); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(how, part, "interpolate ms", best, "ok", float((ist == 0).float().mean()), flush=True)
