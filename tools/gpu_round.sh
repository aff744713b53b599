#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu --set full of the main kernels.
# usage (under gpurun): bash tools/gpu_round.sh TAG
TAG=${1:-rX}
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err && cat $O/${TAG}_bench.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_ncu_bench.log 2>&1
B2ME_BANDS=1 ITERS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_sad_fs -c 1 -f -o $O/${TAG}_sad_fs python tools/prof_fs.py > $O/${TAG}_ncu1.log 2>&1
B2ME_BANDS=1 ITERS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_subpel_refine|k_search_plane' -c 5 -f -o $O/${TAG}_other python tools/prof_fs.py > $O/${TAG}_ncu2.log 2>&1
ND=16384 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_frac_pool|k_fp_' -c 5 -f -o $O/${TAG}_pool python tools/prof_pool.py > $O/${TAG}_ncu3.log 2>&1
timeout 300 python tools/pool_bench.py > $O/${TAG}_pool_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_half|k_quarter|k_apply_wp|k_mc_luma|k_tq|k_distortion|k_bipred|k_cand_dist|k_frac_domain|k_frac_range|k_frac_window|k_frac_decide|k_frac_predict' -c 48 -f -o $O/${TAG}_rest python tools/prof_rest.py > $O/${TAG}_ncu4.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_mc_mb|k_tq16x16|k_tq_chroma|k_dbk_prep|k_deblock|k_epzs|k_sad_table|k_bicand|k_bid_cost|k_expand_pred|k_select_refs|k_gather_best' -c 48 -f -o $O/${TAG}_rest2 python tools/prof_rest2.py > $O/${TAG}_ncu5.log 2>&1
tail -n 2 $O/${TAG}_ncu1.log; tail -n 2 $O/${TAG}_ncu2.log
# summaries are made on the box; only the k_sad_fs report itself travels back (gpurun merges at most 64 MiB)
for n in sad_fs other pool rest rest2; do [ -f $O/${TAG}_$n.ncu-rep ] && python tools/ncu_summary.py $O/${TAG}_$n.ncu-rep $O/${TAG}_$n.ncu.txt > /dev/null 2>&1; done
rm -f $O/${TAG}_other.ncu-rep $O/${TAG}_pool.ncu-rep $O/${TAG}_rest.ncu-rep $O/${TAG}_rest2.ncu-rep $O/${TAG}_sad_fs.ncu-rep
ls -la $O | tail -20
