for st in 0 15000 30000 60000; do echo "stagger=$st"; B200_BA_STAGGER_NS=$st python tools/probe_pre.py 3200002 16 0 2>&1 | grep -v precompute; done
B200_BA_STAGGER_NS=30000 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ba_launches.csv python tools/probe_pre.py 3200002 16 0 > /dev/null 2>&1
