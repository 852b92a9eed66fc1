"""Cross-language check of the oracle against the REAL reference, where it could be built: oracle/_ref/flint_ref_driver is
FLINT's own Fortran (gfortran -O3 -fopenmp, `make -C oracle ref`) behind this repository's OpenMP driver
(oracle/ref_driver.f90).  The build image has no Fortran compiler, so these tests skip there (DESIGN.md "Oracle");
on a machine with gfortran and the reference tree they pin libm `pow`, libgfortran `norm2` and gfortran's evaluation
order -- the four intrinsics the restatement cannot prove."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "flint_ref_driver")
REC = np.dtype([("i", "<i4", 5), ("d", "<f8", 15)])


def test_reference_probe_is_reported():
    """bench.py names what it could run: the probe of compilers / sources / the built driver is part of its JSON line."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    kind, exe, probe = bench.reference_kind()
    assert kind in ("reference", "port") and exe == EXE
    assert set(probe) >= {"gfortran", "reference_sources", "oracle/_ref/flint_ref_driver"}
    assert (kind == "reference") == os.path.exists(EXE)


@pytest.mark.skipif(not os.path.exists(EXE), reason="the reference could not be built here (no Fortran compiler): oracle/_ref is empty")
def test_oracle_equals_the_fortran_reference_on_config3():
    import flint_b200 as F
    from flint_b200 import problems as P
    import cases as K
    n, start = 256, 4096
    dump = os.path.join(ROOT, "oracle", "_ref", "dump.bin")
    subprocess.run([EXE, str(n), str(start), "4", dump], check=True, capture_output=True)
    r = np.fromfile(dump, dtype=REC)
    assert len(r) == n
    case = K.cr3bp_case(F.ERK_DOP853, 1e-12, 0, max_events=2)
    o = case.run_oracle(P.cr3bp_ensemble(n, start=start), pow_mode=0)          # libm pow, as gfortran's `**`
    same_counts = (r["i"][:, 1:5].T == o["counters"]).all(axis=0)
    # the north star's own bar: step counts identical on >= 99.9 % of trajectories, states within 10 * rtol * |y|
    assert same_counts.mean() >= 0.999, same_counts.mean()
    ok = same_counts & (o["status"] == 0)
    assert np.abs(r["d"][ok, 1:7].T - o["yf"][:, ok]).max() <= 1e-11 * np.abs(o["yf"][:, ok]).max()
    assert np.abs(r["d"][ok, 7] - o["events"][0, 0, ok]).max() <= 1e-9           # event abscissa within the event tolerance
