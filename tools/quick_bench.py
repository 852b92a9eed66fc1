"""Quick device-side timing of the north-star kernel (not the contract bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P

tf, clk = F.measure_fp64_peak(0)
print("lib", os.environ.get("FLINT_B200_LIB", "default"), "fp64 DFMA peak: %.2f TFLOP/s" % tf, flush=True)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
blocks = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [128]
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
for arith in (0, 1):
    for evset in (F.FLINT_EVSET_CR3BP_SAMPLE3, F.FLINT_EVSET_NONE):
        sys_, erk, ikw, xf = P.cr3bp_config3(arith=arith, eventset=evset)
        for block in blocks:
            best = 1e30
            for rep in range(2):
                r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, block_threads=block, **ikw)
                best = min(best, r["kernel_ms"])
            cnt = r["counters"].cpu().numpy().astype(np.int64)
            fl = P.algorithmic_flops(F.ERK_DOP853, F.FLINT_SYS_CR3BP_PLANAR, 6, cnt[0].sum(), cnt[1].sum(), N if evset else 0)
            print("N=%d arith=%d events=%d block=%d: %.1f ms  %.3f Morbit/s  %.2f TFLOP/s alg (%.1f%% of peak)" % (
                N, arith, evset, block, best, N / best / 1e3, fl / best / 1e9, 100 * fl / best / 1e9 / tf), flush=True)
