# scratch: auto-plan latency at the smaller circuit sizes and for an emulated rank of 8 / 4 (one GPU)
for n in 100000 200000 400000 800000; do echo "n=$n"; timeout 120 python tools/probe_shard.py $n 1 0 2>&1 | tail -1; done
timeout 200 python tools/probe_shard.py 3200000 8,4,2 0 2>&1 | tail -3
