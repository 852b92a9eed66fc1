for t in t128_256 t64_256 t64_128 t32_64; do
  echo "$t: $(FLINT_B200_LIB=$PWD/flint_b200/libflint_b200_$t.so python tools/prof_interp.py 32768 1 | tail -1)"
done
