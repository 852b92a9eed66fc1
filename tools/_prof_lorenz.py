import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import flint_b200 as F
from flint_b200 import problems as P
import cases as K
case = K.lorenz_case(F.ERK_DOP54, 1e-11, int(sys.argv[1]) if len(sys.argv) > 1 else 0, xf=25.0)
y0 = torch.from_numpy(P.lorenz_ensemble(1 << 18)).cuda()
sys_, erk, st = case.make_erk()
r = erk.IntegrateBatch(case.x0, y0, case.xf, StepSz=0.0, **case.int_kwargs())
print("kernel ms", r["kernel_ms"], "attempts/traj", float(r["counters"][0].double().mean()))
