"""Time library variants of the north-star kernel against each other (development aid).
   python tools/variant_sweep.py tag1 tag2 ...   (libflint_b200_<tag>.so built by flint_b200.build --tag _<tag>)
Each variant runs in its own process (the library is chosen at import); results are compared bit for bit with the first."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
import flint_b200 as F
from flint_b200 import problems as P
N = int(sys.argv[1])
y0 = torch.from_numpy(P.cr3bp_ensemble(N)).cuda()
out = {}
for arith in (0, 1):
    sys_, erk, ikw, xf = P.cr3bp_config3(arith=arith, eventset=F.FLINT_EVSET_CR3BP_SAMPLE3)
    best = 1e30
    for rep in range(3):
        r = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, **ikw)
        best = min(best, r["kernel_ms"])
    h = hashlib.sha256()
    for k in ("Yf", "counters", "Xf", "status"):
        h.update(r[k].cpu().numpy().tobytes())
    out[arith] = (best, h.hexdigest()[:16])
print("RESULT " + json.dumps(out))
''' % ROOT
N = 1 << 20
ref = None
for tag in sys.argv[1:]:
    env = dict(os.environ)
    if tag != "default":
        env["FLINT_B200_LIB"] = os.path.join(ROOT, "flint_b200", "libflint_b200_%s.so" % tag)
    r = subprocess.run([sys.executable, "-c", CHILD, str(N)], env=env, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    if not line:
        print(tag, "FAILED", r.stderr[-400:], flush=True)
        continue
    res = json.loads(line[0][7:])
    if ref is None:
        ref = res
    print("%-8s strict %.1f ms %s   fma %.1f ms %s" % (tag, res["0"][0], "same" if res["0"][1] == ref["0"][1] else "DIFFERENT",
                                                      res["1"][0], "same" if res["1"][1] == ref["1"][1] else "DIFFERENT"), flush=True)
