"""Strong-scaling shard sizes of the north-star step (2^20 orbits / {1,2,4,8} GPUs) against the lane count of the persistent
step kernel: kernel ms per call for each (orbits per GPU, CTA threads).  Development aid (profiles/r02_notes.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P

blocks = [int(b) for b in sys.argv[1].split(",")] if len(sys.argv) > 1 else [512, 448, 384, 320, 256, 192, 128]
sizes = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1 << 17, 1 << 18, 1 << 19]
arith = int(os.environ.get("ARITH", "0"))
for N in sizes:
    y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
    sys_, erk, ikw, xf = P.cr3bp_config3(arith=arith, eventset=F.FLINT_EVSET_CR3BP_SAMPLE3)
    for block in blocks:
        ts = []
        for rep in range(4):
            r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, block_threads=block, MaxEvents=2, **ikw)
            ts.append(r["kernel_ms"])
        best = min(ts[1:])
        print("N=%d block=%d lanes=%d orbits/lane=%.2f: %.2f ms  %.3f Morbit/s  (x%.3f of 2^20-rate 8.64)" % (
            N, block, 148 * block, N / (148.0 * block), best, N / best / 1e3, N / best / 1e3 / 8.64), flush=True)
