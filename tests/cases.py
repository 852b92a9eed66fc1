"""Shared case definitions: the same configuration is handed to the CPU oracle
(tests/oracle_api.py) and to the product (flint_b200 over the C ABI)."""
import numpy as np

import flint_b200 as F
from flint_b200 import problems as P

import oracle_api as O


class Case:
    """One Init + Integrate configuration."""

    def __init__(self, system_id, sysparams, method, rtol, atol, x0, xf, eventset=F.FLINT_EVSET_NONE, max_steps=100000,
                 arith=F.FLINT_ARITH_STRICT, interp_on=None, interp_states=None, max_step=None, min_step=None,
                 stepsz_params=None, events_on=None, event_stepsz=None, event_options=None, event_tol=None,
                 const_step=None, step_advance=None, int_steps_on=None, event_mask=None, stiff_test=None, stepsz=None,
                 max_events=4):
        self.__dict__.update(locals())
        self.n = {1: 6, 2: 6, 3: 3, 4: 6}[system_id]
        self.m = {0: 0, 1: 3, 2: 2, 3: 1, 4: 3, 5: 3}[eventset]
        if events_on is None:
            self.events_on = eventset != F.FLINT_EVSET_NONE

    # ---------------------------------------------------------------- oracle
    def oracle_init(self, pow_mode=1):
        return O.make_init(self.system_id, self.n, self.method, self.rtol, self.atol, self.max_steps, sysparams=self.sysparams,
                           eventset_id=self.eventset, m=self.m, interp_on=self.interp_on, interp_states=self.interp_states,
                           min_step=self.min_step, max_step=self.max_step, stepsz_params=self.stepsz_params,
                           events_on=self.events_on, event_stepsz=self.event_stepsz, event_options=self.event_options,
                           event_tol=self.event_tol, arith_fma=self.arith, pow_mode=pow_mode)

    def oracle_int(self):
        return O.make_int(const_step=self.const_step, step_advance=self.step_advance, int_steps_on=self.int_steps_on,
                          event_mask=self.event_mask, event_states=self.events_on, stiff_test=self.stiff_test)

    def run_oracle(self, y0, xarr=None, pow_mode=1, nthreads=0):
        return O.integrate_ensemble(self.oracle_init(pow_mode), self.oracle_int(), self.x0, y0, self.xf, stepsz=self.stepsz,
                                    ev_cap=self.max_events if self.events_on else 0, xarr=xarr, nthreads=nthreads,
                                    n_interp=len(self.interp_states) if self.interp_states else None)

    def oracle_scalar(self, pow_mode=1):
        return O.OracleERK(self.oracle_init(pow_mode))

    # ---------------------------------------------------------------- product
    def make_erk(self):
        sys_ = F.DiffEqSys(self.system_id, self.eventset, self.sysparams)
        erk = F.ERK()
        st = erk.Init(sys_, self.max_steps, Method=self.method, ATol=np.atleast_1d(self.atol), RTol=np.atleast_1d(self.rtol),
                      InterpOn=self.interp_on, InterpStates=self.interp_states, MinStepSize=self.min_step,
                      MaxStepSize=self.max_step, StepSzParams=self.stepsz_params,
                      EventsOn=self.events_on if self.eventset != F.FLINT_EVSET_NONE or self.events_on else None,
                      EventStepSz=self.event_stepsz, EventOptions=self.event_options, EventTol=self.event_tol, Arith=self.arith)
        return sys_, erk, st

    def int_kwargs(self):
        return dict(UseConstStepSz=self.const_step, StepAdvance=self.step_advance, IntStepsOn=self.int_steps_on,
                    EventMask=self.event_mask, EventStates=bool(self.events_on), StiffTest=self.stiff_test)

    def run_gpu(self, y0, **kw):
        sys_, erk, st = self.make_erk()
        assert st == F.FLINT_SUCCESS, F.status_string(st)
        r = erk.IntegrateBatch(self.x0, y0, self.xf, StepSz=self.stepsz if self.stepsz is not None else 0.0,
                               MaxEvents=self.max_events, **self.int_kwargs(), **kw)
        r["erk"], r["sys"] = erk, sys_
        return r


def cr3bp_case(method=F.ERK_DOP853, rtol=1e-12, arith=0, eventset=F.FLINT_EVSET_CR3BP_SAMPLE3, norb=P.CR3BP_NORB, **kw):
    """Config 3: test_CR3BP's setup."""
    base = dict(stiff_test=1, const_step=False)
    if eventset in (1, 4, 5):
        base.update(event_stepsz=[1e-5, 0.0, 0.0], event_tol=[1e-3, 1e-9, 1e-9], event_mask=[0, 1, 0])
    elif eventset == 2:
        base.update(event_tol=[1e-9, 1e-9])
    base.update(kw)
    return Case(F.FLINT_SYS_CR3BP_PLANAR, [P.CR3BP_MU], method, rtol, rtol * 1e-3, 0.0, P.ARENSTORF_PERIOD * norb,
                eventset=eventset, arith=arith, **base)


def lorenz_case(method=F.ERK_DOP54, rtol=1e-11, arith=0, xf=10.0, **kw):
    return Case(F.FLINT_SYS_LORENZ, list(P.LORENZ_PARAMS), method, rtol, rtol * 1e-3, 0.0, xf, arith=arith, stiff_test=0, **kw)


def halo_case(method, rtol, arith=0, **kw):
    """tests/test_WP.f90 setup."""
    kw.setdefault("stiff_test", 1)
    return Case(F.FLINT_SYS_CR3BP_SPATIAL, [P.HALO_MU], method, rtol, 1e-80, 0.0, P.HALO_PERIOD, max_steps=99000, arith=arith, **kw)


def twobody_case(method=F.ERK_DOP853, rtol=1e-12, arith=0, nper=10, eventset=F.FLINT_EVSET_NONE, **kw):
    kw.setdefault("stiff_test", 1)
    return Case(F.FLINT_SYS_TWOBODY, [P.TBP_GM], method, rtol, rtol * 1e-3, 0.0, P.tbp_period() * nper, eventset=eventset,
                max_steps=10000, arith=arith, min_step=1e-6, max_step=1e6, **kw)


def assert_bit_equal(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = (a == b) | (np.isnan(a) & np.isnan(b)) if a.dtype.kind == "f" else (a == b)
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError("%s: %d of %d entries differ; first at %s: %r vs %r" % (what, len(bad), a.size, i, a[i], b[i]))
    if a.dtype.kind == "f":   # signed zeros too
        assert (np.signbit(a) == np.signbit(b))[~np.isnan(a)].all(), what + ": signed-zero mismatch"


def compare_batch(g, o, case, what=""):
    """GPU result dict vs oracle result dict: bit-exact on everything the reference returns."""
    assert_bit_equal(g["status"], o["status"], what + " status")
    assert_bit_equal(g["counters"], o["counters"], what + " counters")
    assert_bit_equal(g["Xf"], o["xf"], what + " Xf")
    assert_bit_equal(g["Yf"], o["yf"], what + " Yf")
    if g.get("StepSz") is not None and not case.const_step:
        assert_bit_equal(g["StepSz"], o["stepsz"], what + " StepSz")
    if case.events_on:
        assert_bit_equal(g["nEvents"], o["ev_count"], what + " nEvents")
        assert_bit_equal(g["EventStates"], o["events"], what + " EventStates")
    if case.stiff_test is not None:
        assert_bit_equal(g["StiffTest"], o["stiff"], what + " StiffTest")
