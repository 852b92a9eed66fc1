python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "gang or small_shard" 2>&1 | tail -5
for s in 1 32 64 128; do python tools/prof_shard.py 131072 $s 0 3 | tail -1; done
for s in 1 64; do python tools/prof_shard.py 262144 $s 0 3 | tail -1; done
for s in 1 64 128; do python tools/prof_shard.py 524288 $s 0 3 | tail -1; done
for s in 1 128; do python tools/prof_shard.py 1048576 $s 0 3 | tail -1; done
