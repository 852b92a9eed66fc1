for s in 1 48 64 96; do python tools/prof_shard.py 131072 $s 0 4 | tail -1; done
for s in 64 128; do python tools/prof_shard.py 262144 $s 0 3 | tail -1; done
for s in 128; do python tools/prof_shard.py 524288 $s 0 3 | tail -1; done
