#!/bin/bash
# k_sad_fs probes on the bench workload (development): warp-cycle split of -DFS_PROFILE builds
O=gpurun_out/fs_sweep.log; : > $O
for L in $EXTRA_LIBS; do echo "== $L" >> $O; B2ME_LIB=/root/repo/h264_b200/$L B2ME_FS_PROFILE=1 timeout 120 python tools/fs_probe.py 2>&1 | cut -c1-400 >> $O; done
cat $O
