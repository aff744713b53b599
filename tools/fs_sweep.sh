#!/bin/bash
# k_sad_fs probes on the bench workload (development): task repetition (pure task-code time) and the warp-cycle split
O=gpurun_out/fs_sweep.log; : > $O
run() { echo "== LIB=$1 VAR=$2 CLAIM=$3 REP=$4" >> $O
  B2ME_LIB=$1 B2ME_FS_VAR=$2 B2ME_FS_CLAIM=$3 B2ME_FS_REP=$4 timeout 120 python tools/fs_probe.py 2>&1 | cut -c1-400 >> $O; }
for REP in 0 1 2; do run "" 4x3 3 $REP; done
for L in $EXTRA_LIBS; do B2ME_FS_PROFILE=1 run /root/repo/h264_b200/$L 4x3 3 0; done
cat $O
