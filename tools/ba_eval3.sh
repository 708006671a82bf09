echo G2; for k in 16 32 64; do for r in 3 4 5 6; do echo "K=$k r=$r"; B200_BA_K=$k B200_BATCH_AFFINE=$r python tools/probe_pre.py 3200002 16 0 g2 2>&1 | grep -v precompute; done; done
