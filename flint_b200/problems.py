"""Workload definitions of BASELINE.json's configs -- host-side only.

Constants are the reference test programs' (file:line cited per item).  Initial
conditions of the ensembles come from a counter-based generator (splitmix64 keyed
by seed, trajectory index and component), so the CPU oracle and the GPU run see
bit-identical inputs and any slice of an ensemble can be regenerated anywhere.
"""
from fractions import Fraction

import numpy as np

from . import api as F

SEED = 20240607

# ---- planar CR3BP, Arenstorf orbit: tests/test_CR3BP.f90:33,287,334-337 ----
CR3BP_MU = 0.012277471
ARENSTORF_Y0 = np.array([0.994, 0.0, 0.0, 0.0, -2.00158510637908252240537862224, 0.0])
ARENSTORF_PERIOD = 17.0652165601579625588917206249
CR3BP_NORB = 2.5
# ---- spatial CR3BP halo orbit: tests/test_WP.f90:47-57 ----
HALO_MU = 1.0 / (81.30059 + 1.0)
HALO_Y0 = np.array([1.02202151273581740824714855590570360, 0.0, -0.182096761524240501132977765539282777,
                    0.0, -0.103256341062793815791764364248006121, 0.0])
HALO_PERIOD = 1.5111111111111111111111111111111111111111
# ---- Lorenz: tests/test_Lorenz.f90:36-38,120-123 ----
LORENZ_PARAMS = (10.0, 28.0, 8.0 / 3.0)
LORENZ_Y0 = np.array([1.0, 0.0, 0.0])
# ---- two-body Molniya: tests/test_TBP.f90:36,166-171 ----
TBP_GM = 398600.436233
TBP_Y0 = np.array([0.0, -3096.70185149293, -6183.97070198107, 10.0141943725295, 0.0, 0.0])


def wp_rtol(i):
    """rtol(i) = (1.0_WP/10.0_WP)**i evaluated at compile time (tests/test_WP.f90:146): the exact
    power of the double nearest 0.1, rounded once."""
    return float(Fraction(0.1) ** i)


def tbp_period():
    """tests/test_TBP.f90:177-180 (energy, semi-major axis, period), plain double arithmetic."""
    r = np.sqrt(np.sum(TBP_Y0[:3] ** 2))
    v2 = np.sum(TBP_Y0[3:] ** 2)
    energy = v2 / 2 - TBP_GM / r
    sma = -TBP_GM / (2 * energy)
    return 2.0 * np.pi * np.sqrt(sma ** 3 / TBP_GM)


def uniform01(index, comp=0, seed=SEED):
    """u in [0,1): splitmix64 of seed ^ (index*0x9E3779B97F4A7C15 + comp), top 53 bits.  Vectorised."""
    with np.errstate(over="ignore"):
        idx = np.asarray(index, dtype=np.uint64)
        z = np.uint64(seed) ^ (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(comp))
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def cr3bp_ensemble(N, start=0, spread=1e-6, index=None):
    """Config 3/4 ICs: y0_i = (0.994*(1 + spread*u_i), 0, 0, 0, -2.0015..., 0), SoA (6, N).  index: the trajectory indices
    (any order, e.g. a shard of the ensemble, api.shard_indices) instead of start .. start+N-1."""
    i = np.arange(start, start + N, dtype=np.uint64) if index is None else np.asarray(index, dtype=np.uint64)
    N = len(i)
    y0 = np.repeat(ARENSTORF_Y0[:, None], N, axis=1)
    y0[0] = ARENSTORF_Y0[0] * (1.0 + spread * uniform01(i, 0))
    return np.ascontiguousarray(y0)


def lorenz_ensemble(N, start=0, spread=1e-3):
    """Config 2 ICs: (1 + s*u1, s*u2, s*u3), SoA (3, N)."""
    i = np.arange(start, start + N, dtype=np.uint64)
    y0 = np.empty((3, N))
    y0[0] = 1.0 + spread * uniform01(i, 1)
    y0[1] = spread * uniform01(i, 2)
    y0[2] = spread * uniform01(i, 3)
    return y0


def perturbed_ensemble(base, N, start=0, spread=1e-9):
    """Config 5 ICs: IC(1)*(1 + spread*u_i)."""
    i = np.arange(start, start + N, dtype=np.uint64)
    y0 = np.repeat(np.asarray(base, dtype=np.float64)[:, None], N, axis=1)
    scale = 1.0 + spread * uniform01(i, 0)
    k = int(np.flatnonzero(np.asarray(base) != 0.0)[0])
    y0[k] = base[k] * scale
    return np.ascontiguousarray(y0)


def cr3bp_system(eventset=F.FLINT_EVSET_CR3BP_SAMPLE3):
    return F.DiffEqSys(F.FLINT_SYS_CR3BP_PLANAR, eventset, [CR3BP_MU])


def cr3bp_config3(method=F.ERK_DOP853, rtol=1e-12, arith=F.FLINT_ARITH_STRICT, eventset=F.FLINT_EVSET_CR3BP_SAMPLE3,
                  interp_on=False, max_steps=100000):
    """The north-star configuration: test_CR3BP's Init call
    (tests/test_CR3BP.f90:390-392) at rtol 1e-12.  Returns (sys, erk, integrate kwargs, xf)."""
    sys_ = cr3bp_system(eventset)
    erk = F.ERK()
    kw = dict(Method=method, ATol=[rtol * 1e-3], RTol=[rtol], InterpOn=interp_on, EventsOn=eventset != F.FLINT_EVSET_NONE,
              Arith=arith)
    if eventset in (F.FLINT_EVSET_CR3BP_SAMPLE3, F.FLINT_EVSET_CR3BP_SAMPLE3_TERM, F.FLINT_EVSET_CR3BP_SAMPLE3_RESTART):
        kw.update(EventStepSz=[1e-5, 0.0, 0.0], EventTol=[1e-3, 1e-9, 1e-9])
        ikw = dict(EventMask=[False, True, False], EventStates=True, StiffTest=1, UseConstStepSz=False)
    elif eventset == F.FLINT_EVSET_XY_CROSSING:
        kw.update(EventTol=[1e-9, 1e-9])
        ikw = dict(EventStates=True, StiffTest=1)
    else:
        ikw = dict(StiffTest=1)
    st = erk.Init(sys_, max_steps, **kw)
    assert st == F.FLINT_SUCCESS, F.status_string(st)
    return sys_, erk, ikw, ARENSTORF_PERIOD * CR3BP_NORB


# algorithmic flops: fma = 2, add = mul = div = sqrt = 1
FLOPS_RHS = {F.FLINT_SYS_CR3BP_PLANAR: 31, F.FLINT_SYS_CR3BP_SPATIAL: 40, F.FLINT_SYS_LORENZ: 9, F.FLINT_SYS_TWOBODY: 14}
# (per-component flops per attempt, RHS evaluations per attempt, RHS evaluations added on acceptance)
FLOPS_ATTEMPT = {F.ERK_DOP853: (157, 11, 1), F.ERK_DOP54: (63, 6, 0), F.ERK_VERNER98R: (211, 15, 1), F.ERK_VERNER65E: (85, 8, 0)}
# dense coefficients: (per-component flops, extra RHS evaluations)
FLOPS_DENSE = {F.ERK_DOP853: (172, 3), F.ERK_DOP54: (25, 0), F.ERK_VERNER98R: (313, 5), F.ERK_VERNER65E: (85, 1)}


def algorithmic_flops(method, system_id, n, total_attempts, accepted, dense_evals=0):
    """Flops the algorithm needs for the given step counters (sum over the ensemble), charging
    rejected work and divergence honestly; dense_evals = number of dense-coefficient builds that
    were actually evaluated."""
    per_c, nrhs, nacc = FLOPS_ATTEMPT[method]
    fF = FLOPS_RHS[system_id]
    dc, drhs = FLOPS_DENSE[method]
    return (float(total_attempts) * (per_c * n + nrhs * fF) + float(accepted) * nacc * fF
            + float(dense_evals) * (dc * n + drhs * fF))
