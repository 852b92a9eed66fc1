"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (flint_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(_ORACLE_DIR, "liboracle.so")

# enums (src/FLINT_base.f90:43-61,86-120; src/ERK.f90:56-66)
ERK_DOP853, ERK_DOP54, ERK_VERNER98R, ERK_VERNER65E = 17, 18, 19, 20
SYS_CR3BP_PLANAR, SYS_CR3BP_SPATIAL, SYS_LORENZ, SYS_TWOBODY = 1, 2, 3, 4
EVSET_NONE, EVSET_CR3BP_SAMPLE3, EVSET_XY_CROSSING, EVSET_APSIS = 0, 1, 2, 3
EVSET_CR3BP_SAMPLE3_TERM, EVSET_CR3BP_SAMPLE3_RESTART = 4, 5


class OracleInit(C.Structure):
    _fields_ = [
        ("system_id", C.c_int), ("eventset_id", C.c_int), ("n", C.c_int), ("m", C.c_int),
        ("sysparams", C.c_double * 8),
        ("max_steps", C.c_int), ("method", C.c_int),
        ("natol", C.c_int), ("nrtol", C.c_int), ("atol", C.c_double * 8), ("rtol", C.c_double * 8),
        ("has_interp_on", C.c_int), ("interp_on", C.c_int),
        ("n_interp_states", C.c_int), ("interp_states", C.c_int * 8),
        ("has_min_step", C.c_int), ("min_step", C.c_double),
        ("has_max_step", C.c_int), ("max_step", C.c_double),
        ("has_stepsz_params", C.c_int), ("stepsz_params", C.c_double * 6),
        ("has_events_on", C.c_int), ("events_on", C.c_int),
        ("n_event_stepsz", C.c_int), ("event_stepsz", C.c_double * 4),
        ("n_event_options", C.c_int), ("event_options", C.c_int * 4),
        ("n_event_tol", C.c_int), ("event_tol", C.c_double * 4),
        ("arith_fma", C.c_int), ("pow_mode", C.c_int),
    ]


class OracleInt(C.Structure):
    _fields_ = [
        ("has_const_step", C.c_int), ("const_step", C.c_int), ("has_step_advance", C.c_int), ("step_advance", C.c_int),
        ("has_int_steps_on", C.c_int), ("int_steps_on", C.c_int), ("has_xint", C.c_int),
        ("has_event_mask", C.c_int), ("event_mask", C.c_int * 4),
        ("has_event_states", C.c_int), ("has_stiff_test", C.c_int), ("stiff_test", C.c_int),
        ("nparams", C.c_int), ("params", C.c_double * 8),
    ]


_lib = None


def build():
    subprocess.run(["make", "-C", _ORACLE_DIR, "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(OracleInit)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_status.argtypes = [C.c_void_p]
        L.oracle_status_string.restype = C.c_char_p
        L.oracle_integrate.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.POINTER(C.c_double), C.c_void_p,
                                       C.POINTER(C.c_double), C.POINTER(OracleInt), C.c_void_p, C.c_void_p, C.c_long,
                                       C.POINTER(C.c_long), C.c_void_p, C.c_long, C.POINTER(C.c_long),
                                       C.POINTER(C.c_int), C.c_void_p]
        L.oracle_interpolate.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_int]
        L.oracle_dense_steps.restype = C.c_long
        L.oracle_dense_steps.argtypes = [C.c_void_p]
        L.oracle_dense_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_integrate_ensemble.argtypes = [C.POINTER(OracleInit), C.POINTER(OracleInt), C.c_long, C.c_double,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_long, C.c_void_p, C.c_int]
        L.oracle_rhs.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_pow.restype = C.c_double
        L.oracle_pow.argtypes = [C.c_double, C.c_double, C.c_int]
        L.oracle_pow_vec.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int]
        L.oracle_brent.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def make_init(system_id, n, method, rtol, atol, max_steps, sysparams=(), eventset_id=EVSET_NONE, m=0,
              interp_on=None, interp_states=None, min_step=None, max_step=None, stepsz_params=None,
              events_on=None, event_stepsz=None, event_options=None, event_tol=None, arith_fma=0, pow_mode=0):
    ci = OracleInit()
    ci.system_id, ci.eventset_id, ci.n, ci.m = system_id, eventset_id, n, m
    for i, v in enumerate(sysparams):
        ci.sysparams[i] = v
    ci.max_steps, ci.method = max_steps, method
    rt = np.atleast_1d(np.asarray(rtol, dtype=float))
    at = np.atleast_1d(np.asarray(atol, dtype=float))
    ci.natol, ci.nrtol = len(at), len(rt)
    for i, v in enumerate(at):
        ci.atol[i] = v
    for i, v in enumerate(rt):
        ci.rtol[i] = v
    if interp_on is not None:
        ci.has_interp_on, ci.interp_on = 1, int(interp_on)
    if interp_states is not None:
        ci.n_interp_states = len(interp_states)
        for i, v in enumerate(interp_states):
            ci.interp_states[i] = v
    if min_step is not None:
        ci.has_min_step, ci.min_step = 1, min_step
    if max_step is not None:
        ci.has_max_step, ci.max_step = 1, max_step
    if stepsz_params is not None:
        ci.has_stepsz_params = 1
        for i, v in enumerate(stepsz_params):
            ci.stepsz_params[i] = v
    if events_on is not None:
        ci.has_events_on, ci.events_on = 1, int(events_on)
    ci.n_event_stepsz = ci.n_event_options = ci.n_event_tol = -1
    if event_stepsz is not None:
        ci.n_event_stepsz = len(event_stepsz)
        for i, v in enumerate(event_stepsz):
            ci.event_stepsz[i] = v
    if event_options is not None:
        ci.n_event_options = len(event_options)
        for i, v in enumerate(event_options):
            ci.event_options[i] = v
    if event_tol is not None:
        ci.n_event_tol = len(event_tol)
        for i, v in enumerate(event_tol):
            ci.event_tol[i] = v
    ci.arith_fma, ci.pow_mode = int(arith_fma), int(pow_mode)
    return ci


def make_int(const_step=None, step_advance=None, int_steps_on=None, event_mask=None, event_states=False,
             stiff_test=None, params=None):
    io = OracleInt()
    if const_step is not None:
        io.has_const_step, io.const_step = 1, int(const_step)
    if step_advance is not None:
        io.has_step_advance, io.step_advance = 1, int(step_advance)
    if int_steps_on is not None:
        io.has_int_steps_on, io.int_steps_on, io.has_xint = 1, int(int_steps_on), 1
    if event_mask is not None:
        io.has_event_mask = 1
        for i, v in enumerate(event_mask):
            io.event_mask[i] = int(v)
    io.has_event_states = int(bool(event_states))
    if stiff_test is not None:
        io.has_stiff_test, io.stiff_test = 1, int(stiff_test)
    if params is not None:
        io.nparams = len(params)
        for i, v in enumerate(params):
            io.params[i] = v
    return io


class OracleERK:
    """One ERK_class object of the oracle (scalar Init / Integrate / Interpolate / Info)."""

    def __init__(self, init):
        self.init = init
        self.n = init.n
        self.h = lib().oracle_create(C.byref(init))
        self.counters = np.zeros(4, dtype=np.int32)

    def __del__(self):
        try:
            if self.h:
                lib().oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def status(self):
        return lib().oracle_status(self.h)

    def integrate(self, x0, y0, xf, stepsz=0.0, io=None, xint_cap=0, ev_cap=64):
        io = io or make_int()
        n = self.n
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        yf = np.zeros(n)
        xf_c, h_c = C.c_double(xf), C.c_double(stepsz)
        xint = np.zeros(max(xint_cap, 1))
        yint = np.zeros((max(xint_cap, 1), n))
        ev = np.zeros((ev_cap, n + 2))
        nx, ne, stiff = C.c_long(0), C.c_long(0), C.c_int(0)
        st = lib().oracle_integrate(self.h, x0, _ptr(y0), C.byref(xf_c), _ptr(yf), C.byref(h_c), C.byref(io),
                                    _ptr(xint) if xint_cap else None, _ptr(yint) if xint_cap else None, xint_cap,
                                    C.byref(nx), _ptr(ev), ev_cap, C.byref(ne), C.byref(stiff), _ptr(self.counters))
        return dict(status=st, xf=xf_c.value, yf=yf, stepsz=h_c.value, xint=xint[:min(nx.value, xint_cap)],
                    yint=yint[:min(nx.value, xint_cap)], n_xint=nx.value, events=ev[:min(ne.value, ev_cap)].copy(),
                    n_events=ne.value, stiff=stiff.value, total=int(self.counters[0]), accepted=int(self.counters[1]),
                    rejected=int(self.counters[2]), fcalls=int(self.counters[3]))

    def interpolate(self, xarr, save=None, n_states=None):
        xarr = np.ascontiguousarray(xarr, dtype=np.float64)
        nI = n_states or self.n
        yarr = np.zeros((len(xarr), nI))
        st = lib().oracle_interpolate(self.h, _ptr(xarr), len(xarr), _ptr(yarr), int(save is not None), int(bool(save)))
        return st, yarr

    def dense(self, n_states=None, pstar=None):
        ns = lib().oracle_dense_steps(self.h)
        nI = n_states or self.n
        xint = np.zeros(ns + 1)
        bip = np.zeros((ns, pstar + 1, nI))
        lib().oracle_dense_copy(self.h, _ptr(xint), _ptr(bip))
        return xint, bip


def integrate_ensemble(init, io, x0, y0_soa, xf, stepsz=None, ev_cap=0, xarr=None, nthreads=0, n_interp=None):
    """y0_soa: (n, N).  Returns dict of SoA outputs like the product's batch ABI."""
    n, N = y0_soa.shape
    y0_soa = np.ascontiguousarray(y0_soa, dtype=np.float64)
    xf_a = np.ascontiguousarray(np.broadcast_to(np.asarray(xf, dtype=np.float64), (N,))).copy()
    h_a = np.zeros(N) if stepsz is None else np.ascontiguousarray(np.broadcast_to(np.asarray(stepsz, dtype=np.float64), (N,))).copy()
    yf = np.zeros((n, N))
    counters = np.zeros((4, N), dtype=np.int32)
    status = np.zeros(N, dtype=np.int32)
    ev_count = np.zeros(N, dtype=np.int32) if ev_cap else None
    events = np.full((n + 2, ev_cap, N), np.nan) if ev_cap else None
    stiff = np.full(N, io.stiff_test, dtype=np.int32)
    yarr = None
    G = 0
    if xarr is not None:
        xarr = np.ascontiguousarray(xarr, dtype=np.float64)
        G = len(xarr)
        yarr = np.zeros((N, G, n_interp or n))
    used = lib().oracle_integrate_ensemble(C.byref(init), C.byref(io), N, float(x0), _ptr(y0_soa), _ptr(xf_a), _ptr(yf),
                                           _ptr(h_a), _ptr(counters), _ptr(status), ev_cap, _ptr(ev_count), _ptr(events),
                                           _ptr(stiff), _ptr(xarr), G, _ptr(yarr), nthreads)
    return dict(xf=xf_a, yf=yf, stepsz=h_a, counters=counters, status=status, ev_count=ev_count, events=events,
                stiff=stiff, yarr=yarr, threads=used)


def rhs(system_id, sysparams, y, fma=0, x=0.0):
    sp = np.zeros(8)
    sp[:len(sysparams)] = sysparams
    y = np.ascontiguousarray(y, dtype=np.float64)
    f = np.zeros(len(y))
    lib().oracle_rhs(system_id, _ptr(sp), len(y), fma, x, _ptr(y), _ptr(f))
    return f
