#!/bin/bash
# k_sad_fs probes on the bench workload (development): task-body builds x CTA shapes
O=gpurun_out/fs_sweep.log; : > $O
run() { echo "== LIB=$1 VAR=$2" >> $O
  B2ME_LIB=$1 B2ME_FS_VAR=$2 timeout 120 python tools/fs_probe.py 2>&1 | grep -v "^\[" | cut -c1-300 >> $O; }
for L in "" $EXTRA_LIBS; do for V in 4x3 3x3; do run "${L:+/root/repo/h264_b200/$L}" $V; done; done
cat $O
