#!/bin/bash
# k_sad_fs variant sweep (development): variant libraries x CTA shapes (workers x CTAs per SM), pan workload + robustness cases
# usage: bash tools/fs_shapes.sh "lib:shape lib:shape ..."   (libs relative to h264_b200/)
O=gpurun_out/fs_shapes.log; : > $O
for ls in $1; do
  L=${ls%%:*}; V=${ls##*:}
  echo "== lib=$L var=$V" >> $O
  B2ME_LIB=/root/repo/h264_b200/$L B2ME_FS_VAR=$V timeout 120 python tools/fs_probe.py 2>&1 | grep -E "k_sad_fs|checksum" | cut -c1-200 >> $O
  B2ME_LIB=/root/repo/h264_b200/$L B2ME_FS_VAR=$V timeout 200 python tools/fs_robust.py 2>&1 | cut -c1-120 >> $O
done
cat $O
