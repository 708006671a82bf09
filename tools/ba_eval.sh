B200_BATCH_AFFINE=3 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k msm 2>&1 | tail -2
for g in "" g2; do for ba in 0 auto; do echo "BA=$ba $g"; B200_BATCH_AFFINE=$ba python tools/probe_pre.py 3200002 16 0 $g 2>&1 | grep -v precompute; done; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ba_launches.csv python tools/probe_pre.py 3200002 16 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ba_launches_g2.csv python tools/probe_pre.py 3200002 16 0 g2 > /dev/null 2>&1
