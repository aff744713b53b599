#!/bin/bash
# k_sad_fs variant sweep on the bench workload (development): library build x claim size
O=gpurun_out/fs_sweep.log; : > $O
run() { echo "== LIB=$1 VAR=$2 CLAIM=$3" >> $O
  B2ME_LIB=$1 B2ME_FS_VAR=$2 B2ME_FS_CLAIM=$3 timeout 120 python tools/fs_probe.py 2>&1 | grep -v "^\[" | cut -c1-200 >> $O; }
for CL in 3 4 6; do run "" 4x3 $CL; done
run "" 7x2 3
for L in $EXTRA_LIBS; do for CL in 3 6; do run /root/repo/h264_b200/$L 4x3 $CL; done; done
cat $O
