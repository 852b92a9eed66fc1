for rep in 1 2; do
for lib in "" _b1 _b2; do
  echo "lib=$lib $(FLINT_B200_LIB=$PWD/flint_b200/libflint_b200$lib.so python tools/prof_shard.py 131072 64 0 4 | tail -1)"
done; done
