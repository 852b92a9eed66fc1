#!/bin/bash
# Config-4 Interpolate ([N][G][nI], 2^15 orbits x 10^4 points): thread-per-item kernel (default and tagged builds) against the
# step-major kernel.  python -m flint_b200.build --tag _iX -D FLINT_ITEMS_P=.. -D FLINT_ITEMS_THREADS=.. -D FLINT_ITEMS_CTAS=..
cd "$(dirname "$0")/.."
for arith in 0 1; do
for lib in flint_b200/libflint_b200.so flint_b200/libflint_b200_i*.so; do
  [ -f "$lib" ] || continue
  echo "== items $lib arith $arith"
  FLINT_B200_INTERP_KERNEL=items FLINT_B200_LIB=$PWD/$lib timeout 200 python tools/prof_interp.py ${1:-32768} 1 $arith 0 2>&1 | tail -3
done
echo "== steps arith $arith"
FLINT_B200_INTERP_KERNEL=steps timeout 200 python tools/prof_interp.py ${1:-32768} 1 $arith 0 2>&1 | tail -3
done
