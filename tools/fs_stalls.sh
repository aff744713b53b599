#!/bin/bash
# k_sad_fs: issue / stall metrics of variant libraries (development; one k_sad_fs launch each under ncu)
# usage: bash tools/fs_stalls.sh "lib lib ..."
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
for k in no_instruction math_pipe_throttle not_selected short_scoreboard long_scoreboard wait branch_resolving dispatch_stall mio_throttle barrier sleeping membar; do M=$M,smsp__average_warps_issue_stalled_${k}_per_issue_active.ratio; done
O=gpurun_out/fs_stalls.log; : > $O
for L in $1; do
  echo "== $L" >> $O
  B2ME_LIB=/root/repo/h264_b200/$L ITERS=1 timeout 300 ncu --metrics $M --clock-control none -k regex:k_sad_fs -c 1 --csv python tools/prof_fs.py 2>&1 | grep -E '"k_sad_fs|^"[0-9]' | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' >> $O
done
cat $O
