set -x
python -m pytest tests -m gpu -x -q -k "best_tiles_in_parts or rle_by_chunks or large_tiles or rle or golden or smoke" > gpurun_out/r3b_tests.log 2>&1
tail -15 gpurun_out/r3b_tests.log
python tools/big_tiles.py 7 > gpurun_out/r3b_big_best.log 2>&1; cat gpurun_out/r3b_big_best.log
B="python bench.py --no-cpu-baseline --no-e2e --no-others"
$B --workload c2best --steps 3 --warmup 3 > gpurun_out/r3b_c2best.log 2>&1
$B --workload c3best --steps 3 --warmup 3 > gpurun_out/r3b_c3best.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_kernel|rle_kernel|decode_kernel" -c 12 --csv --log-file gpurun_out/r3b_launch_c2best.csv $B --workload c2best --steps 1 --warmup 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_kernel|rle_kernel|decode_kernel" -c 12 --csv --log-file gpurun_out/r3b_launch_c3best.csv $B --workload c3best --steps 1 --warmup 1 > /dev/null 2>&1
grep -h "rle_kernel\|encode_kernel" gpurun_out/r3b_launch_c2best.csv gpurun_out/r3b_launch_c3best.csv | cut -c1-200
