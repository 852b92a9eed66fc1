"""Throughput of the north-star step kernel against resident CTAs per SM (1 CTA = 1 warp per sub-partition), with the
ensemble scaled so that every lane gets the same number of orbits (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P
for arith in (0, 1):
    for bps in (1, 2, 3, 4):
        N = bps * (1 << 18)
        y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
        sys_, erk, ikw, xf = P.cr3bp_config3(arith=arith, eventset=F.FLINT_EVSET_CR3BP_SAMPLE3)
        best = 1e30
        for rep in range(2):
            r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, blocks_per_sm=bps, **ikw)
            best = min(best, r["kernel_ms"])
        print("arith %d  %d CTA/SM  N=%d: %.1f ms  %.3f M orbits/s  (%.3f per CTA/SM)" % (arith, bps, N, best, N / best / 1e3, N / best / 1e3 / bps), flush=True)
