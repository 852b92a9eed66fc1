"""CPU-side tests of the product: the C-ABI library loads and exports every symbol include/flint_b200.h
declares, Init validation mirrors the reference's error codes, and compute entry points FAIL LOUDLY
without a CUDA device (there is no CPU fallback).  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import flint_b200 as F
from flint_b200 import api, problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "flint_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(flint_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = api.lib()
    syms = declared_symbols()
    assert len(syms) >= 16 and "flint_erk_integrate_batch" in syms
    for s in syms:
        assert hasattr(lib, s), "libflint_b200.so does not export %s" % s
    assert lib.flint_b200_kernel_count() == 48      # 4 methods x (2 systems x 4 + 2 systems x 2) (arithmetic, events) entries


def test_enum_values_are_the_references():
    assert (F.ERK_DOP853, F.ERK_DOP54, F.ERK_VERNER98R, F.ERK_VERNER65E) == (17, 18, 19, 20)       # src/ERK.f90:56-66
    assert (F.FLINT_EVENTACTION_CHANGESOLY, F.FLINT_EVENTACTION_RESTARTINT, F.FLINT_EVENTACTION_MASK) == (2, 4, 128)
    hdr = open(os.path.join(ROOT, "include", "flint_b200.h")).read()
    for name, val in (("FLINT_ERROR_CONSTSTEPSZ", 16), ("FLINT_ERROR_EVENTROOT", 15), ("FLINT_EVENTOPTION_STEPEND", 2)):
        assert re.search(r"%s = %d\b" % (name, val), hdr)


def test_status_strings_verbatim():
    ref = ["Successfully Completed", "Termination Event Occurred", "Problem appears to be stiff", "Unknown error",
           "Incorrect user-provided parameters", "Maximum integration steps reached", "Memory allocation error",
           "Memory deallocation error", "Incorrect dimensions", "Step size reduced to very small value",
           "Incorrect interpolation states", "Incorrect interpolation array provided", "Interpolation is not enabled",
           "Init is required", "Incorrect event-related parameters ", "Event root finding failed",
           "Incorrect step size for the non-adaptive option"]              # src/FLINT_base.f90:64-82
    assert [F.status_string(i) for i in range(17)] == ref


def test_init_validation_matches_reference_codes():
    s = F.DiffEqSys(F.FLINT_SYS_CR3BP_PLANAR, F.FLINT_EVSET_CR3BP_SAMPLE3, [P.CR3BP_MU])
    e = F.ERK()
    ok = dict(ATol=[1e-14], RTol=[1e-11])
    assert e.Init(s, 1000, Method=F.ERK_DOP853, **ok) == F.FLINT_SUCCESS
    assert e.Init(s, 0, **ok) == F.FLINT_ERROR_PARAMS                                        # src/ERKInit.f90:68-73
    assert e.Init(s, 10, ATol=[1e-14, 1e-14], RTol=[1e-11]) == F.FLINT_ERROR_PARAMS          # :76-87
    assert e.Init(s, 10, ATol=[1e-14] * 6, RTol=[1e-11] * 6) == F.FLINT_SUCCESS
    assert e.Init(s, 10, EventsOn=True, EventStepSz=[1e-5, 0], **ok) == F.FLINT_ERROR_DIMENSION       # :147-158
    assert e.Init(s, 10, EventsOn=True, EventOptions=[0, 1, 3], **ok) == F.FLINT_ERROR_EVENTPARAMS    # :161-176
    assert e.Init(s, 10, EventsOn=True, EventTol=[1e-9], **ok) == F.FLINT_ERROR_DIMENSION             # :179-188
    assert e.Init(s, 10, InterpOn=True, InterpStates=[1, 7], **ok) == F.FLINT_ERROR_INTPSTATES        # :296-304
    assert e.Init(s, 10, InterpOn=True, InterpStates=[2, 4], **ok) == F.FLINT_SUCCESS
    noev = F.DiffEqSys(F.FLINT_SYS_LORENZ, F.FLINT_EVSET_NONE, P.LORENZ_PARAMS)
    assert F.ERK().Init(noev, 10, EventsOn=True, **ok) == F.FLINT_ERROR_EVENTPARAMS          # :141 (G not associated)
    # faithful quirk: with InterpOn left on from an earlier Init the InterpStates block resets the status (:304-308)
    assert e.Init(noev, 10, EventsOn=True, **ok) == F.FLINT_SUCCESS
    with pytest.raises(ValueError):
        F.DiffEqSys(F.FLINT_SYS_LORENZ, F.FLINT_EVSET_APSIS, P.LORENZ_PARAMS)               # event set of another system


def test_integrate_argument_checks_and_loud_failure_without_gpu():
    s = F.DiffEqSys(F.FLINT_SYS_TWOBODY, F.FLINT_EVSET_NONE, [P.TBP_GM])
    e = F.ERK()
    assert e.Init(s, 100, Method=F.ERK_DOP54, ATol=[1e-12], RTol=[1e-9], InterpOn=True) == 0
    y0 = P.perturbed_ensemble(P.TBP_Y0, 8)
    r = e.IntegrateBatch(0.0, y0, 10.0, IntStepsOn=True)                # IntStepsOn with InterpOn: src/ERKIntegrate.f90:120-127
    assert r["rc"] == F.FLINT_ERROR_PARAMS
    e2 = F.ERK()
    assert e2.Interpolate([0.0, 1.0])[0] in (F.FLINT_ERROR_INTP_OFF, F.FLINT_ERROR_INIT_REQD)
    if api.lib().flint_b200_device_count() == 0:
        assert e.Init(s, 100, Method=F.ERK_DOP54, ATol=[1e-12], RTol=[1e-9]) == 0
        with pytest.raises(RuntimeError, match="cuda|CUDA"):
            e.IntegrateBatch(0.0, y0, 10.0)                             # no device, no fallback: loud
        r = e.Integrate(0.0, P.TBP_Y0, 10.0)
        assert r["status"] == F.FLINT_ERROR and "CUDA" in F.last_error().upper() or "cuda" in F.last_error()


def test_product_never_touches_the_oracle():
    """The product (flint_b200/, include/) must not include, import or link anything under oracle/."""
    bad = []
    for root in (os.path.join(ROOT, "flint_b200"), os.path.join(ROOT, "include")):
        for dp, _, fns in os.walk(root):
            if "_build" in dp:
                continue
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dp, fn)).read()
                    if re.search(r"oracle_api|liboracle|oracle/|import oracle|from oracle", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_seeded_ensembles_are_reproducible_and_sliceable():
    a = P.cr3bp_ensemble(1000)
    b = P.cr3bp_ensemble(100, start=500)
    assert np.array_equal(a[:, 500:600], b)
    u = P.uniform01(np.arange(200000))
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    assert (a[0] >= 0.994).all() and (a[0] <= 0.994 * (1 + 1e-6)).all() and (a[4] == P.ARENSTORF_Y0[4]).all()


def test_generated_sources_are_current():
    """tools/gen_tableaus.py output matches what is committed (only where the reference tree is present)."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not present on this machine")
    import subprocess, sys, tempfile, shutil
    tmp = tempfile.mkdtemp()
    try:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_tableaus.py"), "--root", tmp], check=True,
                       capture_output=True)
        for rel in ("oracle/oracle_tableaus.h",):
            assert open(os.path.join(tmp, rel)).read() == open(os.path.join(ROOT, rel)).read(), rel
    finally:
        shutil.rmtree(tmp)


def _build_c_example(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cr3bp_abi")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "cr3bp_abi.c"), "-o", exe,
                    "-L", os.path.join(root, "flint_b200"), "-lflint_b200", "-Wl,-rpath," + os.path.join(root, "flint_b200"), "-lm"],
                   check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    return dict(l.split(" ", 1) for l in out.strip().splitlines())


def test_c_program_links_against_the_abi_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/cr3bp_abi.c makes the calls of the reference's tests/test_CR3BP.f90 from plain C (what the Fortran shim
    binds).  Init is host-side validation and succeeds anywhere; Integrate needs a CUDA device and says so."""
    r = _build_c_example(tmp_path)
    assert r["init_status"].startswith("0 ")
    import flint_b200 as F
    if F.lib().flint_b200_device_count() == 0:
        assert r["integrate_status"].startswith("3 ") and "detail" in r


@pytest.mark.gpu
def test_c_program_reproduces_the_reference_known_answers(tmp_path):
    """Same program on a GPU: the event record of tests/results_CR3BP.txt:11 (x = 0.399136, y1 = 0.748351, event 2) and
    step counts within 1.5 % of :5-8 (the golden used randomised ICs and ifort fast-math)."""
    r = _build_c_example(tmp_path)
    assert r["integrate_status"].startswith("0 ")
    ev = r["event1"].split()
    assert abs(float(ev[1]) - 0.399136) < 2e-6 and abs(float(ev[3]) - 0.748351) < 2e-6 and ev[5] == "2"
    st = r["steps"].split()
    assert abs(int(st[2]) - 680) <= 10 and int(r["events"]) == 1
    assert float(r["jacobi_drift"]) < 1e-9


def _run_user_system_check():
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "flint_b200", "libflint_b200_user.so")
    if not os.path.exists(lib):
        pytest.skip("libflint_b200_user.so not built (python __graft_entry__.py builds it)")
    env = dict(os.environ, FLINT_B200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "user_systems", "check_user.py")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_user_registered_systems_host_side():
    """Build-time registration of user F/G functors: tests/user_systems/example_systems.cuh compiled into its
    own library; ids, n and m are validated against the registered functor."""
    assert "OK validation" in _run_user_system_check()


@pytest.mark.gpu
def test_user_registered_systems_on_gpu():
    """A user-registered Lorenz functor equals the oracle's Lorenz bit for bit; a user Van der Pol oscillator with an upward
    zero-crossing event agrees with scipy's DOP853 (final state and event times to 1e-8)."""
    out = _run_user_system_check()
    assert "OK user Lorenz == oracle Lorenz" in out and "OK user Van der Pol vs scipy" in out


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): stdout is exactly one JSON line with the contract's
    keys; everything else goes to stderr."""
    import json
    env = dict(os.environ, FLINT_BENCH_REF_SAMPLE="1024")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "CR3BP DOP853 orbits/sec" and d["unit"] == "orbits/s" and d["value"] > 0
    for key in ("n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def _c_struct_fields(name):
    """Field names of `typedef struct { ... } name;` in include/flint_b200.h, in order."""
    txt = open(os.path.join(ROOT, "include", "flint_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    body = re.search(r"typedef struct \{([^{}]*)\} %s;" % name, txt).group(1)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            m = re.search(r"(\w+)\s*(\[[^\]]*\])?\s*$", part.strip())
            names.append(m.group(1))
    return names


def _fortran_type_fields(name):
    txt = open(os.path.join(ROOT, "fortran", "flint_gpu.f90")).read()
    body = re.search(r"type, bind\(C\)(?:, public)? :: %s\n(.*?)end type" % name, txt, flags=re.S).group(1)
    names = []
    for line in body.splitlines():
        line = line.split("!")[0]
        if "::" not in line:
            continue
        for part in re.split(r",(?![^()]*\))", line.split("::")[1]):
            names.append(re.match(r"\s*(\w+)", part).group(1))
    return names


def test_abi_structs_agree_between_c_header_ctypes_and_fortran():
    """The three descriptions of the ABI structs (include/flint_b200.h, flint_b200/api.py, fortran/flint_gpu.f90) list the same
    fields in the same order -- the Fortran module cannot be compiled in this image, so at least its layout is pinned."""
    strip = lambda names: [n.rstrip("_") for n in names]  # noqa: E731
    for cname, ctype in (("flint_exec", api._Exec), ("flint_init_opts", api._InitOpts), ("flint_int_opts", api._IntOpts),
                         ("flint_steps_out", api._StepsOut), ("flint_events_out", api._EventsOut)):
        c = _c_struct_fields(cname)
        assert strip([f[0] for f in ctype._fields_]) == c, cname
        assert _fortran_type_fields(cname) == c, cname
    # size check of the largest one against the C compiler's own layout
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "flint_b200.h"\nint main(void){printf("%zu %zu %zu", sizeof(flint_exec), sizeof(flint_init_opts), sizeof(flint_int_opts));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "s.c"), "-o", os.path.join(td, "s")], check=True)
        sizes = [int(v) for v in subprocess.run([os.path.join(td, "s")], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(api._Exec), C.sizeof(api._InitOpts), C.sizeof(api._IntOpts)]


RT_SOURCE = os.path.join(ROOT, "tests", "user_systems", "runtime_systems.cuh")


def test_runtime_registration_from_source_host_side(tmp_path, monkeypatch):
    """flint_sys_register_source (csrc/rtc.cu): a user functor's SOURCE is compiled by NVRTC against the kernel
    headers embedded in the library.  Registration needs no GPU: the functor's n / m_max / id come back from the compiled probe,
    flint_sys_create validates against them, a source that does not compile returns the compiler's log, an id cannot be taken
    twice, and flint_sys_precompile builds + caches the integrate kernels' cubin."""
    monkeypatch.setenv("FLINT_B200_RTC_CACHE", str(tmp_path))
    src = open(RT_SOURCE).read()
    assert F.register_system_source("RtLorenz", src, False) == F.FLINT_SYS_USER_BASE + 10
    assert F.register_system_source("RtVanDerPol", src, True) == F.FLINT_SYS_USER_BASE + 11
    assert F.register_system_source("RtLorenz", src, False) == F.FLINT_SYS_USER_BASE + 10           # idempotent
    F.DiffEqSys(F.FLINT_SYS_USER_BASE + 10, F.FLINT_EVSET_NONE, [10.0, 28.0, 8.0 / 3.0], n=3, m=0)
    F.DiffEqSys(F.FLINT_SYS_USER_BASE + 11, 1, [1.5], n=2, m=1)
    for bad in (dict(system_id=F.FLINT_SYS_USER_BASE + 10, n=4, m=0), dict(system_id=F.FLINT_SYS_USER_BASE + 10, n=3, m=1),
                dict(system_id=F.FLINT_SYS_USER_BASE + 11, n=2, m=2), dict(system_id=F.FLINT_SYS_USER_BASE + 12, n=3, m=0)):
        with pytest.raises(ValueError):
            F.DiffEqSys(bad["system_id"], F.FLINT_EVSET_NONE if bad["m"] == 0 else 1, [1.0], n=bad["n"], m=bad["m"])
    with pytest.raises(ValueError, match="different source"):
        F.register_system_source("RtLorenz", src + "\n/* changed */\n", False)
    broken = src.replace("RtLorenz", "RtBroken").replace("USER_BASE + 10", "USER_BASE + 13").replace("sigma * (y[1] - y[0]);", "sigma * (y[1] - y[0])")
    with pytest.raises(ValueError, match="expected a \";\""):
        F.register_system_source("RtBroken", broken, False)
    with pytest.raises(ValueError):
        F.register_system_source("NoSuchStruct", src, False)
    with pytest.raises(ValueError, match="no event function"):
        F.precompile_system(F.FLINT_SYS_USER_BASE + 10, F.ERK_DOP853, 0, True)
    F.precompile_system(F.FLINT_SYS_USER_BASE + 11, F.ERK_DOP54, F.FLINT_ARITH_FMA, True)
    blobs = [f for f in os.listdir(tmp_path) if f.startswith("rt_") and f.endswith(".bin")]
    assert len(blobs) == 1 and os.path.getsize(os.path.join(tmp_path, blobs[0])) > 100000
    F.precompile_system(F.FLINT_SYS_USER_BASE + 11, F.ERK_DOP54, F.FLINT_ARITH_FMA, True)            # served from the cache
    assert len(os.listdir(tmp_path)) == 1
