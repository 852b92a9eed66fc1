"""Host-side mirror of the reference's object API over the C ABI (include/flint_b200.h).

    reference (Fortran)                         here
    ------------------------------------------  -----------------------------------------
    type, extends(DiffEqSys) :: sys             DiffEqSys(system_id, eventset_id, sysparams)
    type(ERK_class) :: erk                      ERK()
    call erk%Init(sys, MaxSteps, Method=..)     erk.Init(sys, MaxSteps, Method=..)
    call erk%Integrate(X0, Y0, Xf, Yf, ..)      erk.Integrate(X0, Y0, Xf, ..)          one trajectory
                                                erk.IntegrateBatch(X0, Y0[n,N], Xf, ..) ensemble
    call erk%Interpolate(Xarr, Yarr, Save)      erk.Interpolate(Xarr, Save) / erk.InterpolateBatch(..)
    call erk%Info(status, msg, nSteps=..)       erk.Info()

Argument names, defaults and status codes are the reference's (src/FLINT_base.f90:318-533).
This module is plumbing (ctypes + numpy, torch tensors accepted as device buffers); all
compute happens in libflint_b200.so.  It raises at import time of the library if the
CUDA extension has not been built -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

# ---- enums (values are ABI; src/FLINT_base.f90:43-61,86-120; src/ERK.f90:56-66) ----
FLINT_SUCCESS, FLINT_EVENT_TERM, FLINT_STIFF_PROBLEM, FLINT_ERROR, FLINT_ERROR_PARAMS = 0, 1, 2, 3, 4
FLINT_SYS_USER_BASE = 100
FLINT_ERROR_MAXSTEPS, FLINT_ERROR_MEMALLOC, FLINT_ERROR_MEMDEALLOC, FLINT_ERROR_DIMENSION = 5, 6, 7, 8
FLINT_ERROR_STEPSZ_TOOSMALL, FLINT_ERROR_INTPSTATES, FLINT_ERROR_INTP_ARRAY, FLINT_ERROR_INTP_OFF = 9, 10, 11, 12
FLINT_ERROR_INIT_REQD, FLINT_ERROR_EVENTPARAMS, FLINT_ERROR_EVENTROOT, FLINT_ERROR_CONSTSTEPSZ = 13, 14, 15, 16
FLINT_EVENTACTION_CONTINUE, FLINT_EVENTACTION_TERMINATE, FLINT_EVENTACTION_CHANGESOLY = 0, 1, 2
FLINT_EVENTACTION_RESTARTINT, FLINT_EVENTACTION_MASK = 4, 128
FLINT_EVENTOPTION_ROOTFINDING, FLINT_EVENTOPTION_STEPBEGIN, FLINT_EVENTOPTION_STEPEND = 0, 1, 2
ERK_DOP853, ERK_DOP54, ERK_VERNER98R, ERK_VERNER65E = 17, 18, 19, 20
FLINT_SYS_CR3BP_PLANAR, FLINT_SYS_CR3BP_SPATIAL, FLINT_SYS_LORENZ, FLINT_SYS_TWOBODY = 1, 2, 3, 4
FLINT_EVSET_NONE, FLINT_EVSET_CR3BP_SAMPLE3, FLINT_EVSET_XY_CROSSING, FLINT_EVSET_APSIS = 0, 1, 2, 3
FLINT_EVSET_CR3BP_SAMPLE3_TERM, FLINT_EVSET_CR3BP_SAMPLE3_RESTART = 4, 5
FLINT_ARITH_STRICT, FLINT_ARITH_FMA = 0, 1
FLINT_MEM_HOST, FLINT_MEM_DEVICE = 0, 1

_INIT_BITS = dict(METHOD=1 << 0, INTERP_ON=1 << 1, INTERP_STATES=1 << 2, MIN_STEP=1 << 3, MAX_STEP=1 << 4,
                  STEPSZ_PARAMS=1 << 5, EVENTS_ON=1 << 6, EVENT_STEPSZ=1 << 7, EVENT_OPTIONS=1 << 8,
                  EVENT_TOL=1 << 9, ARITH=1 << 10)
_INT_BITS = dict(CONST_STEPSZ=1 << 0, STEP_ADVANCE=1 << 1, INTSTEPS_ON=1 << 2, XINT_YINT=1 << 3, EVENT_MASK=1 << 4,
                 EVENT_STATES=1 << 5, STIFF_TEST=1 << 6, PARAMS=1 << 7)

_SYS_N = {FLINT_SYS_CR3BP_PLANAR: 6, FLINT_SYS_CR3BP_SPATIAL: 6, FLINT_SYS_LORENZ: 3, FLINT_SYS_TWOBODY: 6}
_EVSET_M = {FLINT_EVSET_NONE: 0, FLINT_EVSET_CR3BP_SAMPLE3: 3, FLINT_EVSET_XY_CROSSING: 2, FLINT_EVSET_APSIS: 1,
            FLINT_EVSET_CR3BP_SAMPLE3_TERM: 3, FLINT_EVSET_CR3BP_SAMPLE3_RESTART: 3}
_PSTAR = {ERK_DOP853: 7, ERK_DOP54: 4, ERK_VERNER98R: 8, ERK_VERNER65E: 5}


class _InitOpts(C.Structure):
    _fields_ = [("present", C.c_uint), ("method", C.c_int),
                ("n_atol", C.c_int), ("atol", C.POINTER(C.c_double)), ("n_rtol", C.c_int), ("rtol", C.POINTER(C.c_double)),
                ("interp_on", C.c_int), ("n_interp_states", C.c_int), ("interp_states", C.POINTER(C.c_int)),
                ("min_step_size", C.c_double), ("max_step_size", C.c_double), ("stepsz_params", C.c_double * 6),
                ("events_on", C.c_int), ("n_event_stepsz", C.c_int), ("event_stepsz", C.POINTER(C.c_double)),
                ("n_event_options", C.c_int), ("event_options", C.POINTER(C.c_int)),
                ("n_event_tol", C.c_int), ("event_tol", C.POINTER(C.c_double)), ("arith", C.c_int)]


class _IntOpts(C.Structure):
    _fields_ = [("present", C.c_uint), ("use_const_stepsz", C.c_int), ("step_advance", C.c_int), ("int_steps_on", C.c_int),
                ("event_mask", C.c_int * 4), ("n_params", C.c_int), ("params", C.POINTER(C.c_double))]


class _StepsOut(C.Structure):
    _fields_ = [("cap", C.c_long), ("xint", C.c_void_p), ("yint", C.c_void_p), ("count", C.c_void_p)]


class _EventsOut(C.Structure):
    _fields_ = [("cap", C.c_long), ("rec", C.c_void_p), ("count", C.c_void_p)]


FLINT_MAX_DEVICES = 16
FLINT_SHARD_CHUNK = 4096


class _Exec(C.Structure):
    _fields_ = [("mem", C.c_int), ("device", C.c_int), ("stream", C.c_void_p), ("async_", C.c_int),
                ("dense_cap", C.c_long), ("block_threads", C.c_int), ("blocks_per_sm", C.c_int), ("event_mode", C.c_int), ("plane_variant", C.c_int), ("pipeline", C.c_int),
                ("round_len", C.c_int), ("n_devices", C.c_int), ("devices", C.c_int * FLINT_MAX_DEVICES), ("shard_chunk", C.c_long), ("dense_mode", C.c_int), ("dense_spill", C.c_int)]


def _exec(mem, device=0, stream=None, async_=0, dense_cap=0, block_threads=0, blocks_per_sm=0, event_mode=0, plane_variant=0,
          pipeline=0, round_len=0, devices=None, shard_chunk=0, dense_mode=0, dense_spill=0):
    ex = _Exec()
    ex.dense_mode, ex.dense_spill = int(dense_mode), int(dense_spill)
    ex.mem, ex.device, ex.stream, ex.async_, ex.dense_cap = int(mem), int(device), stream, int(async_), int(dense_cap)
    ex.block_threads, ex.blocks_per_sm, ex.event_mode, ex.plane_variant = int(block_threads), int(blocks_per_sm), int(event_mode), int(plane_variant)
    ex.pipeline, ex.round_len, ex.shard_chunk = int(pipeline), int(round_len), int(shard_chunk)
    if devices is not None:
        devices = list(devices)
        assert len(devices) <= FLINT_MAX_DEVICES
        ex.n_devices = len(devices)
        for k, d in enumerate(devices):
            ex.devices[k] = int(d)
    return ex


LIB_PATH = os.environ.get("FLINT_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libflint_b200.so")
_lib = None


def lib():
    """Load libflint_b200.so.  Fails loudly: the product has no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("flint_b200: %s not built -- run `python -m flint_b200.build` (nvcc, sm_100a). "
                               "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.flint_sys_create.restype = C.c_void_p
        L.flint_sys_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.flint_sys_destroy.argtypes = [C.c_void_p]
        L.flint_sys_register_source.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.flint_erk_create.restype = C.c_void_p
        L.flint_erk_destroy.argtypes = [C.c_void_p]
        L.flint_erk_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_InitOpts)]
        L.flint_erk_integrate.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.POINTER(C.c_double), C.c_void_p,
                                          C.POINTER(C.c_double), C.POINTER(_IntOpts), C.POINTER(_StepsOut),
                                          C.POINTER(_EventsOut), C.POINTER(C.c_int)]
        L.flint_erk_integrate_batch.argtypes = [C.c_void_p, C.c_long, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.POINTER(_IntOpts), C.c_void_p, C.c_void_p,
                                                C.POINTER(_StepsOut), C.POINTER(_EventsOut), C.c_void_p, C.POINTER(_Exec)]
        L.flint_erk_interpolate.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_int]
        L.flint_erk_interpolate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_void_p,
                                                  C.c_int, C.c_int, C.POINTER(_Exec)]
        L.flint_erk_interpolate_batch_grids.argtypes = L.flint_erk_interpolate_batch.argtypes
        L.flint_erk_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 6 + [C.POINTER(C.c_double), C.c_void_p,
                                                                                C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                                                C.c_void_p]
        L.flint_status_string.restype = C.c_char_p
        L.flint_status_string.argtypes = [C.c_int]
        L.flint_last_error.restype = C.c_char_p
        L.flint_b200_measure_fp64_peak.restype = C.c_double
        L.flint_b200_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.flint_b200_launch_count.restype = C.c_long
        L.flint_b200_last_kernel_ms.restype = C.c_double
        L.flint_b200_selftest_arith.restype = C.c_int
        L.flint_b200_selftest_arith.argtypes = [C.c_long, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
        L.flint_b200_shard_count.restype = C.c_long
        L.flint_b200_shard_count.argtypes = [C.c_long, C.c_int, C.c_int, C.c_long]
        L.flint_b200_shard_global_index.restype = C.c_long
        L.flint_b200_shard_global_index.argtypes = [C.c_long, C.c_int, C.c_int, C.c_long, C.c_long]
        L.flint_b200_shard_indices.restype = C.c_long
        L.flint_b200_shard_indices.argtypes = [C.c_long, C.c_int, C.c_int, C.c_long, C.c_void_p]
        _lib = L
    return _lib


def status_string(code):
    return lib().flint_status_string(int(code)).decode()


def last_error():
    return lib().flint_last_error().decode()


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    """Raw pointer of a numpy array or a torch tensor (contiguous)."""
    if x is None:
        return None
    if _is_torch(x):
        assert x.is_contiguous()
        return x.data_ptr()
    assert x.flags["C_CONTIGUOUS"]
    return x.ctypes.data


def register_system_source(struct_name, source, has_events=False):
    """Run-time registration of a user DiffEqSys (include/flint_b200.h: flint_sys_register_source): `source` is the text of a
    header defining `struct <struct_name> : flint::UserSystem<n, m_max>` (F = accel<FMA>, G = events); NVRTC compiles it against
    the library's kernel headers.  Returns the system id for DiffEqSys(id, ...).  Raises ValueError with the compiler's log."""
    out = C.c_int(0)
    rc = lib().flint_sys_register_source(struct_name.encode(), source.encode(), int(bool(has_events)), C.byref(out))
    if rc != FLINT_SUCCESS:
        raise ValueError("flint_sys_register_source (%s): %s" % (status_string(rc), last_error()))
    return out.value


def precompile_system(system_id, Method=ERK_DOP853, Arith=FLINT_ARITH_STRICT, EventsOn=False):
    """Compile and cache the integrate kernels of a system registered from source, without a GPU (flint_sys_precompile)."""
    rc = lib().flint_sys_precompile(int(system_id), int(Method), int(Arith), int(bool(EventsOn)))
    if rc != FLINT_SUCCESS:
        raise ValueError("flint_sys_precompile (%s): %s" % (status_string(rc), last_error()))


class DiffEqSys:
    """`type, extends(DiffEqSys)` of the reference (src/FLINT_base.f90:220-227): n, m, F, G.
    F and G are device functors compiled into the library, selected by id."""

    def __init__(self, system_id, eventset_id=FLINT_EVSET_NONE, sysparams=(), n=None, m=None):
        """n, m: only for user systems (ids >= FLINT_SYS_USER_BASE), whose sizes the library checks against the functor."""
        self.system_id, self.eventset_id = system_id, eventset_id
        self.n = _SYS_N[system_id] if n is None else int(n)
        self.m = _EVSET_M[eventset_id] if m is None else int(m)
        self.sysparams = np.asarray(list(sysparams), dtype=np.float64)
        self._h = lib().flint_sys_create(system_id, eventset_id, self.n, self.m, _ptr(self.sysparams), len(self.sysparams))
        if not self._h:
            raise ValueError("flint_sys_create: " + last_error())

    def __del__(self):
        try:
            if self._h:
                lib().flint_sys_destroy(self._h)
                self._h = None
        except Exception:
            pass


class ERK:
    """`type(ERK_class)` of the reference (src/ERK.f90:72-117)."""

    def __init__(self):
        self._h = lib().flint_erk_create()
        self.status = FLINT_SUCCESS
        self.sys = None
        self._keep = []
        self.method = ERK_DOP853
        self.nI = 0

    def __del__(self):
        try:
            if self._h:
                lib().flint_erk_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ Init
    def Init(self, DE, MaxSteps, Method=None, ATol=None, RTol=None, InterpOn=None, InterpStates=None, MinStepSize=None,
             MaxStepSize=None, StepSzParams=None, EventsOn=None, EventStepSz=None, EventOptions=None, EventTol=None,
             Arith=None):
        o = _InitOpts()
        keep = []

        def darr(v):
            a = np.ascontiguousarray(np.atleast_1d(np.asarray(v, dtype=np.float64)))
            keep.append(a)
            return len(a), a.ctypes.data_as(C.POINTER(C.c_double))

        def iarr(v):
            a = np.ascontiguousarray(np.atleast_1d(np.asarray(v, dtype=np.int32)))
            keep.append(a)
            return len(a), a.ctypes.data_as(C.POINTER(C.c_int))

        if Method is not None:
            o.present |= _INIT_BITS["METHOD"]
            o.method = int(Method)
            self.method = int(Method)
        if ATol is not None:
            o.n_atol, o.atol = darr(ATol)
        if RTol is not None:
            o.n_rtol, o.rtol = darr(RTol)
        if InterpOn is not None:
            o.present |= _INIT_BITS["INTERP_ON"]
            o.interp_on = int(bool(InterpOn))
        if InterpStates is not None:
            o.present |= _INIT_BITS["INTERP_STATES"]
            o.n_interp_states, o.interp_states = iarr(InterpStates)
        if MinStepSize is not None:
            o.present |= _INIT_BITS["MIN_STEP"]
            o.min_step_size = float(MinStepSize)
        if MaxStepSize is not None:
            o.present |= _INIT_BITS["MAX_STEP"]
            o.max_step_size = float(MaxStepSize)
        if StepSzParams is not None:
            o.present |= _INIT_BITS["STEPSZ_PARAMS"]
            for i, v in enumerate(StepSzParams):
                o.stepsz_params[i] = float(v)
        if EventsOn is not None:
            o.present |= _INIT_BITS["EVENTS_ON"]
            o.events_on = int(bool(EventsOn))
        if EventStepSz is not None:
            o.present |= _INIT_BITS["EVENT_STEPSZ"]
            o.n_event_stepsz, o.event_stepsz = darr(EventStepSz)
        if EventOptions is not None:
            o.present |= _INIT_BITS["EVENT_OPTIONS"]
            o.n_event_options, o.event_options = iarr(EventOptions)
        if EventTol is not None:
            o.present |= _INIT_BITS["EVENT_TOL"]
            o.n_event_tol, o.event_tol = darr(EventTol)
        if Arith is not None:
            o.present |= _INIT_BITS["ARITH"]
            o.arith = int(Arith)
        self.sys = DE
        self.status = lib().flint_erk_init(self._h, DE._h, int(MaxSteps), C.byref(o))
        self.MaxSteps = int(MaxSteps)
        if self.status == FLINT_SUCCESS:
            self.InterpOn = bool(InterpOn) if InterpOn is not None else getattr(self, "InterpOn", False)
            ev = bool(EventsOn) if EventsOn is not None else False
            self.nI = len(np.atleast_1d(InterpStates)) if (InterpStates is not None and not ev) else DE.n
        return self.status

    # ------------------------------------------------------------------ options
    def _int_opts(self, UseConstStepSz, StepAdvance, IntStepsOn, EventMask, params, keep):
        io = _IntOpts()
        if UseConstStepSz is not None:
            io.present |= _INT_BITS["CONST_STEPSZ"]
            io.use_const_stepsz = int(bool(UseConstStepSz))
        if StepAdvance is not None:
            io.present |= _INT_BITS["STEP_ADVANCE"]
            io.step_advance = int(bool(StepAdvance))
        if IntStepsOn is not None:
            io.present |= _INT_BITS["INTSTEPS_ON"]
            io.int_steps_on = int(bool(IntStepsOn))
        if EventMask is not None:
            io.present |= _INT_BITS["EVENT_MASK"]
            for i, v in enumerate(EventMask):
                io.event_mask[i] = int(bool(v))
        if params is not None:
            a = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
            keep.append(a)
            io.present |= _INT_BITS["PARAMS"]
            io.n_params, io.params = len(a), a.ctypes.data_as(C.POINTER(C.c_double))
        return io

    # ------------------------------------------------------------------ Integrate (one trajectory)
    def Integrate(self, X0, Y0, Xf, StepSz=0.0, UseConstStepSz=None, StepAdvance=None, IntStepsOn=None,
                  EventMask=None, EventStates=False, StiffTest=None, params=None, MaxEvents=64):
        """Returns a dict: status, Xf, Yf, StepSz, Xint, Yint, EventStates[(nev, n+2)], StiffTest."""
        n = self.sys.n
        keep = []
        io = self._int_opts(UseConstStepSz, StepAdvance, IntStepsOn, EventMask, params, keep)
        y0 = np.ascontiguousarray(np.asarray(Y0, dtype=np.float64))
        yf = np.zeros(n)
        xf, h = C.c_double(float(Xf)), C.c_double(float(StepSz))
        steps = ev = None
        scount, ecount = C.c_int(0), C.c_int(0)
        if IntStepsOn:
            cap = self.MaxSteps + 1
            xint, yint = np.zeros(cap), np.zeros((cap, n))
            steps = _StepsOut(cap, xint.ctypes.data, yint.ctypes.data, C.addressof(scount))
        if EventStates:
            rec = np.zeros((MaxEvents, n + 2))
            ev = _EventsOut(MaxEvents, rec.ctypes.data, C.addressof(ecount))
        stiff = C.c_int(int(StiffTest)) if StiffTest is not None else None
        st = lib().flint_erk_integrate(self._h, float(X0), _ptr(y0), C.byref(xf), _ptr(yf), C.byref(h), C.byref(io),
                                       C.byref(steps) if steps else None, C.byref(ev) if ev else None,
                                       C.byref(stiff) if stiff is not None else None)
        self.status = st
        out = dict(status=st, Xf=xf.value, Yf=yf, StepSz=h.value, StiffTest=stiff.value if stiff is not None else None)
        if IntStepsOn:
            k = min(scount.value, steps.cap)
            out["Xint"], out["Yint"] = xint[:k].copy(), yint[:k].copy()
        if EventStates:
            out["EventStates"] = rec[:min(ecount.value, MaxEvents)].copy()
            out["nEvents"] = ecount.value
        return out

    # ------------------------------------------------------------------ Integrate (ensemble)
    def IntegrateBatch(self, X0, Y0, Xf, StepSz=None, UseConstStepSz=None, StepAdvance=None, IntStepsOn=None,
                       EventMask=None, EventStates=False, StiffTest=None, params=None, MaxEvents=4, StepsCap=None,
                       out=None, device=0, stream=None, dense_cap=0, block_threads=0, blocks_per_sm=0, want_counters=True,
                       event_mode=0, plane_variant=0, async_=False, pipeline=0, round_len=0, devices=None, shard_chunk=0, dense_mode=0,
                       dense_spill=0):
        """Y0: (n, N) structure-of-arrays, numpy (host buffers: the call stages H2D/D2H itself) or a
        torch CUDA tensor (device buffers: nothing is copied).  Xf scalar or (N,).  Returns a dict of
        SoA outputs (numpy or torch, matching the input kind).  async_ (device buffers only): return after enqueueing
        on the stream; the caller synchronises before reading the outputs.  devices=[d0, d1, ..] (host buffers): the ensemble
        is sharded over these GPUs inside the call (block-cyclic chunks of shard_chunk trajectories); round_len: see flint_exec."""
        n = self.sys.n
        keep = []
        io = self._int_opts(UseConstStepSz, StepAdvance, IntStepsOn, EventMask, params, keep)
        on_dev = _is_torch(Y0)
        N = int(Y0.shape[1])
        assert Y0.shape[0] == n
        if on_dev:
            import torch
            dev = Y0.device
            device = dev.index if dev.index is not None else 0
            f64 = dict(dtype=torch.float64, device=dev)
            i32 = dict(dtype=torch.int32, device=dev)
            zeros = lambda shape, kw: torch.zeros(shape, **kw)  # noqa: E731
            full = lambda shape, v, kw: torch.full(shape, v, **kw)  # noqa: E731
            xf = Xf.clone() if _is_torch(Xf) else full((N,), float(Xf), f64)
            x0v = X0 if _is_torch(X0) else None
            h = None if StepSz is None else (StepSz.clone() if _is_torch(StepSz) else full((N,), float(StepSz), f64))
            stiff = None if StiffTest is None else (StiffTest.clone() if _is_torch(StiffTest) else full((N,), int(StiffTest), i32))
            if stream is None:
                stream = torch.cuda.current_stream(dev).cuda_stream
        else:
            f64, i32 = dict(dtype=np.float64), dict(dtype=np.int32)
            zeros = lambda shape, kw: np.zeros(shape, **kw)  # noqa: E731
            full = lambda shape, v, kw: np.full(shape, v, **kw)  # noqa: E731
            Y0 = np.ascontiguousarray(Y0, dtype=np.float64)
            x0v = np.ascontiguousarray(X0, dtype=np.float64) if isinstance(X0, np.ndarray) else None

            def inout(key, value, dtype):      # in/out array: the caller's (e.g. pinned) buffer filled in place, or a fresh one
                if value is None:
                    return None
                buf = (out or {}).get(key)
                if buf is None:
                    buf = np.empty((N,), dtype=dtype)
                buf[...] = value
                return buf
            xf = inout("Xf", Xf, np.float64)
            h = inout("StepSz", StepSz, np.float64)
            stiff = inout("StiffTest", StiffTest, np.int32)
        out = out or {}
        yf = out.get("Yf") if out.get("Yf") is not None else zeros((n, N), f64)
        status = out.get("status") if out.get("status") is not None else zeros((N,), i32)
        counters = (out.get("counters") if out.get("counters") is not None else zeros((4, N), i32)) if want_counters else None
        steps = ev = None
        res = {}
        if IntStepsOn:
            cap = int(StepsCap or (self.MaxSteps + 1))
            xint, yint, scount = zeros((cap, N), f64), zeros((n, cap, N), f64), zeros((N,), i32)
            steps = _StepsOut(cap, _ptr(xint), _ptr(yint), _ptr(scount))
            res.update(Xint=xint, Yint=yint, nXint=scount)
        if EventStates:
            rec = out.get("EventStates") if out.get("EventStates") is not None else full((n + 2, MaxEvents, N), float("nan"), f64)
            ecount = out.get("nEvents") if out.get("nEvents") is not None else zeros((N,), i32)
            ev = _EventsOut(MaxEvents, _ptr(rec), _ptr(ecount))
            res.update(EventStates=rec, nEvents=ecount)
        ex = _exec(FLINT_MEM_DEVICE if on_dev else FLINT_MEM_HOST, device, stream, int(bool(async_) and on_dev), dense_cap,
                   block_threads, blocks_per_sm, event_mode, plane_variant, pipeline, round_len, devices, shard_chunk, dense_mode, dense_spill)
        x0s = float(X0) if x0v is None else 0.0
        st = lib().flint_erk_integrate_batch(self._h, N, x0s, _ptr(x0v), _ptr(Y0), _ptr(xf), _ptr(yf), _ptr(h), C.byref(io),
                                             _ptr(counters), _ptr(status), C.byref(steps) if steps else None,
                                             C.byref(ev) if ev else None, _ptr(stiff), C.byref(ex))
        self.status = st
        if st == FLINT_ERROR:
            raise RuntimeError("flint_erk_integrate_batch: " + last_error())
        res.update(rc=st, Xf=xf, Yf=yf, StepSz=h, status=status, counters=counters, StiffTest=stiff,
                   kernel_ms=lib().flint_b200_last_kernel_ms())
        return res

    # ------------------------------------------------------------------ Interpolate
    def Interpolate(self, Xarr, SaveInterpCoeffs=None):
        xarr = np.ascontiguousarray(np.asarray(Xarr, dtype=np.float64))
        yarr = np.zeros((len(xarr), self.nI))
        st = lib().flint_erk_interpolate(self._h, _ptr(xarr), len(xarr), _ptr(yarr), int(SaveInterpCoeffs is not None),
                                         int(bool(SaveInterpCoeffs)))
        self.status = st
        return st, yarr

    def InterpolateBatch(self, Xarr, N, SaveInterpCoeffs=None, layout=0, out=None, device=0, stream=None):
        """Xarr: (G,) one grid for the ensemble, or (N, G) one grid per trajectory.  layout 0: Yarr (N, G, nI), each trajectory a
        Fortran Yarr(nI, G); layout 1: (nI, G, N)."""
        on_dev = _is_torch(Xarr)
        per_traj = len(Xarr.shape) == 2
        G = int(Xarr.shape[-1])
        shape = (N, G, self.nI) if layout == 0 else (self.nI, G, N)
        if on_dev:
            import torch
            dev = Xarr.device
            device = dev.index if dev.index is not None else 0
            yarr = out if out is not None else torch.empty(shape, dtype=torch.float64, device=dev)
            status = torch.zeros((N,), dtype=torch.int32, device=dev)
            if stream is None:
                stream = torch.cuda.current_stream(dev).cuda_stream
        else:
            Xarr = np.ascontiguousarray(np.asarray(Xarr, dtype=np.float64))
            yarr = out if out is not None else np.empty(shape)
            status = np.zeros((N,), dtype=np.int32)
        ex = _exec(FLINT_MEM_DEVICE if on_dev else FLINT_MEM_HOST, device, stream)
        fn = lib().flint_erk_interpolate_batch_grids if per_traj else lib().flint_erk_interpolate_batch
        st = fn(self._h, _ptr(Xarr), G, _ptr(yarr), int(layout), _ptr(status),
                int(SaveInterpCoeffs is not None), int(bool(SaveInterpCoeffs)), C.byref(ex))
        self.status = st
        if st == FLINT_ERROR:
            raise RuntimeError("flint_erk_interpolate_batch: " + last_error())
        return st, yarr, status

    # ------------------------------------------------------------------ Info
    def Info(self):
        n = self.sys.n if self.sys else 0
        ints = [C.c_int(0) for _ in range(6)]
        x0, hf, xf = C.c_double(0), C.c_double(0), C.c_double(0)
        y0, yf = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        st = lib().flint_erk_info(self._h, *[C.byref(i) for i in ints], C.byref(x0), _ptr(y0), C.byref(hf), C.byref(xf), _ptr(yf))
        return dict(LastStatus=ints[0].value, StatusMsg=status_string(ints[0].value), nSteps=ints[1].value,
                    nAccept=ints[2].value, nReject=ints[3].value, nFCalls=ints[4].value, InterpReady=bool(ints[5].value),
                    X0=x0.value, Y0=y0[:n], hf=hf.value, Xf=xf.value, Yf=yf[:n], status=st)


def measure_fp64_peak(device=0):
    """DFMA microbenchmark: (TFLOP/s, reported SM clock MHz)."""
    clk = C.c_double(0)
    tf = lib().flint_b200_measure_fp64_peak(int(device), C.byref(clk))
    return tf, clk.value


def selftest_arith(n=1 << 26, seed=12345):
    """(mismatches, divisions checked, square roots checked) of the div/sqrt helper self-test."""
    out = (C.c_ulonglong * 3)()
    st = lib().flint_b200_selftest_arith(int(n), int(seed), out)
    if st != FLINT_SUCCESS:
        raise RuntimeError("selftest_arith failed: " + last_error())
    return int(out[0]), int(out[1]), int(out[2])


def launch_count():
    return int(lib().flint_b200_launch_count())


def shard_count(N, n_shards, shard, chunk=FLINT_SHARD_CHUNK):
    """Trajectories of `shard` in the block-cyclic deal of [0, N) over n_shards (devices of a call / ranks of a job)."""
    return int(lib().flint_b200_shard_count(int(N), int(n_shards), int(shard), int(chunk)))


def shard_indices(N, n_shards, shard, chunk=FLINT_SHARD_CHUNK):
    """Global trajectory indices of `shard`, in the shard's local order (int64 array)."""
    cnt = shard_count(N, n_shards, shard, chunk)
    out = np.empty(cnt, dtype=np.int64)
    got = lib().flint_b200_shard_indices(int(N), int(n_shards), int(shard), int(chunk), out.ctypes.data)
    assert got == cnt
    return out
