O=gpurun_out/s5_v10.log; : > $O
for L in libb2me.so libb2me_st4.so; do echo "== $L" >> $O; B2ME_LIB=/root/repo/h264_b200/$L timeout 300 python -m pytest tests/test_gpu_pool.py -x -q 2>&1 | tail -2 >> $O; B2ME_LIB=/root/repo/h264_b200/$L timeout 200 python tools/pool_bench.py 2>&1 | cut -c1-100 >> $O; done
cat $O
