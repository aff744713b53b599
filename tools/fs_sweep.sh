#!/bin/bash
# k_sad_fs probes on the bench workload (development)
O=gpurun_out/fs_sweep.log; : > $O
run() { echo "== SLEEP=$1 CLAIM=$2" >> $O
  B2ME_FS_SLEEP=$1 B2ME_FS_CLAIM=$2 timeout 120 python tools/fs_probe.py 2>&1 | grep -v "^\[" | cut -c1-300 >> $O; }
for S in 0 1 0 1; do run $S 3; done
cat $O
