import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flint_b200 as F
from flint_b200 import problems as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
G = 10000
sys_, erk, ikw, xf = P.cr3bp_config3(method=F.ERK_VERNER98R, rtol=1e-11, arith=0, interp_on=True)
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, dense_cap=640, **ikw)
print("integrate ms", r["kernel_ms"])
acc = r["counters"][1].double().sum().item()
rec_bytes = acc * (54 + 1) * 8                      # Verner98R: 9 x 6 coefficients + the step's end abscissa
xarr = torch.linspace(0.0, xf, G, dtype=torch.float64).cuda(); xarr[-1] = xf
for layout in (0, 1):
    yarr = torch.empty((N, G, 6) if layout == 0 else (6, G, N), dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); erk.InterpolateBatch(xarr, N, SaveInterpCoeffs=True, layout=layout, out=yarr); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("layout %d: %.3f ms  written %.0f GB/s  written+records read %.0f GB/s  checksum %r" % (
        layout, best, N * G * 48 / best / 1e6, (N * G * 48 + rec_bytes) / best / 1e6, float(yarr.double().sum().item())), flush=True)
    del yarr
