for lib in "" _sb1 _sb4; do
  for s in 1 2 3 4 6 8 14; do
    echo "lib=$lib"; FLINT_B200_LIB=$PWD/flint_b200/libflint_b200$lib.so python tools/prof_shard.py 131072 $s 0 3 | tail -1
  done
done
echo FMA
for s in 1 2 4 8 14; do python tools/prof_shard.py 131072 $s 1 3 | tail -1; done
