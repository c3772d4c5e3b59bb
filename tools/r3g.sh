set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3g_tests.log 2>&1
tail -4 gpurun_out/r3g_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r3g_bench.log 2>&1
tail -c 600 gpurun_out/r3g_bench.log
timeout 300 python tools/big_tiles.py 7 > gpurun_out/r3g_big_best.log 2>&1; cat gpurun_out/r3g_big_best.log
timeout 300 python tools/big_tiles.py 8 > gpurun_out/r3g_big_ftl.log 2>&1; cat gpurun_out/r3g_big_ftl.log
