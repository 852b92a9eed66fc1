"""Runs in a subprocess with FLINT_B200_LIB = the library built from example_systems.cuh (the library path is fixed at import).
Prints OK lines; any assertion failure makes the parent test fail."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import flint_b200 as F
from flint_b200 import problems as P
import cases as K

USER_LORENZ, USER_VDP = F.FLINT_SYS_USER_BASE + 0, F.FLINT_SYS_USER_BASE + 1
assert F.lib().flint_b200_kernel_count() == 24
# ---- validation of ids and sizes (host side)
for bad in (dict(system_id=USER_LORENZ, n=4, m=0), dict(system_id=USER_VDP, n=2, m=2), dict(system_id=F.FLINT_SYS_USER_BASE + 7, n=3, m=0)):
    try:
        F.DiffEqSys(bad["system_id"], F.FLINT_EVSET_NONE if bad["m"] == 0 else 1, [1.0], n=bad["n"], m=bad["m"])
        raise SystemExit("accepted a bad user system: %r" % bad)
    except ValueError:
        pass
print("OK validation")
if F.lib().flint_b200_device_count() == 0:
    raise SystemExit(0)

# ---- the user-registered Lorenz functor reproduces the oracle's Lorenz bit for bit
for method, arith in ((F.ERK_DOP54, 0), (F.ERK_DOP853, 1), (F.ERK_VERNER98R, 0)):
    case = K.lorenz_case(method, 1e-10, arith, xf=8.0)
    y0 = P.lorenz_ensemble(300)
    o = case.run_oracle(y0)
    sys_ = F.DiffEqSys(USER_LORENZ, F.FLINT_EVSET_NONE, list(P.LORENZ_PARAMS), n=3, m=0)
    erk = F.ERK()
    assert erk.Init(sys_, case.max_steps, Method=method, ATol=[case.atol], RTol=[case.rtol], Arith=arith) == 0
    g = erk.IntegrateBatch(0.0, y0, 8.0, StepSz=0.0, StiffTest=0)
    for a, b, what in ((g["Yf"], o["yf"], "Yf"), (g["counters"], o["counters"], "counters"), (g["status"], o["status"], "status")):
        assert np.array_equal(np.asarray(a), np.asarray(b)), what
print("OK user Lorenz == oracle Lorenz")

# ---- Van der Pol with an event, against scipy
from scipy.integrate import solve_ivp
mu, xf = 1.5, 20.0
N = 64
y0 = np.stack([2.0 + 1e-3 * P.uniform01(np.arange(N), 5), np.zeros(N)])
sys_ = F.DiffEqSys(USER_VDP, 1, [mu], n=2, m=1)
erk = F.ERK()
assert erk.Init(sys_, 100000, Method=F.ERK_DOP853, ATol=[1e-13], RTol=[1e-12], EventsOn=True, EventTol=[1e-11]) == 0
g = erk.IntegrateBatch(0.0, y0, xf, StepSz=0.0, EventStates=True, MaxEvents=8)
assert (g["status"] == 0).all()
up = lambda t, y: y[0]
up.direction = 1
for i in (0, 17, 63):
    r = solve_ivp(lambda t, y: [y[1], mu * (1 - y[0] ** 2) * y[1] - y[0]], (0.0, xf), y0[:, i], method="DOP853", rtol=1e-12, atol=1e-13, events=up)
    assert np.allclose(g["Yf"][:, i], r.y[:, -1], rtol=0, atol=1e-8), (g["Yf"][:, i], r.y[:, -1])
    te = r.t_events[0]
    assert g["nEvents"][i] == len(te), (g["nEvents"][i], len(te))
    assert np.allclose(g["EventStates"][0, :len(te), i], te, rtol=0, atol=1e-8)
    assert (g["EventStates"][3, :len(te), i] == 1.0).all()
print("OK user Van der Pol vs scipy: %d events per trajectory" % g["nEvents"][0])
