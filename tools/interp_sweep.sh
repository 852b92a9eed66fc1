#!/bin/bash
# Config-4 Interpolate ([N][G][nI], 2^15 orbits) under every tagged build of the persistent tile kernel present in flint_b200/
# (python -m flint_b200.build --tag _tX -D FLINT_TILE_STEPS=.. -D FLINT_TILE_CONSUMERS=.. -D FLINT_TILE_STAGES=.. -D FLINT_TILE_CTAS=..).
cd "$(dirname "$0")/.."
for lib in flint_b200/libflint_b200.so flint_b200/libflint_b200_t*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"
  FLINT_B200_INTERP_KERNEL=tiles FLINT_B200_LIB=$PWD/$lib timeout 200 python tools/prof_interp.py ${1:-32768} 1 ${2:-0} 0 2>&1 | tail -2
done
echo "== one tile per CTA (round-2 first kernel)"
FLINT_B200_INTERP_KERNEL=steps timeout 200 python tools/prof_interp.py ${1:-32768} 1 ${2:-0} 0 2>&1 | tail -2
