import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flint_b200 as F
from flint_b200 import problems as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
arith = int(sys.argv[2]) if len(sys.argv) > 2 else 0
evsets = [int(e) for e in sys.argv[3].split(",")] if len(sys.argv) > 3 else [F.FLINT_EVSET_CR3BP_SAMPLE3, F.FLINT_EVSET_NONE]
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
for evset in evsets:
    sys_, erk, ikw, xf = P.cr3bp_config3(arith=arith, eventset=evset)
    r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, **ikw)
    print("evset", evset, "kernel ms", r["kernel_ms"])
