O=gpurun_out/s5_v9.log; : > $O
timeout 300 python -m pytest tests/test_gpu_pool.py -x -q 2>&1 | tail -2 >> $O; timeout 200 python tools/pool_bench.py 2>&1 | cut -c1-330 >> $O
cat $O
