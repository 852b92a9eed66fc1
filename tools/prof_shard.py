"""One north-star call on N orbits with a given slice setting (for ncu; development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import flint_b200 as F
from flint_b200 import problems as P
N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
slices = int(sys.argv[2]) if len(sys.argv) > 2 else 0
arith = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
sys_, erk, ikw, xf = P.cr3bp_config3(arith=arith)
for _ in range(reps):
    r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, round_len=slices, MaxEvents=2, **ikw)
    print("N", N, "slices", slices, "kernel ms", r["kernel_ms"], flush=True)
