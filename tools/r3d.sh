set -x
B="python bench.py --no-cpu-baseline --no-e2e --no-others"
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 2 -c 1 -o gpurun_out/r3d_full_dec $B --workload c2 --steps 1 --warmup 2 > gpurun_out/r3d_ncu.log 2>&1
ncu -i gpurun_out/r3d_full_dec.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r3d_src_dec.csv 2>/dev/null
ncu -i gpurun_out/r3d_full_dec.ncu-rep --page raw --csv > gpurun_out/r3d_raw_dec.csv 2>/dev/null
rm -f gpurun_out/r3d_full_dec.ncu-rep
timeout 300 $B --workload c2 --tiles 1024 --steps 3 --warmup 3 > gpurun_out/r3d_c2_1024.log 2>&1
timeout 300 $B --workload c2 --tiles 148 --steps 3 --warmup 3 > gpurun_out/r3d_c2_148.log 2>&1
grep -o '"decode_ms": [0-9.]*' gpurun_out/r3d_c2_1024.log gpurun_out/r3d_c2_148.log
timeout 600 python -m pytest tests -m gpu -x -q -k "cli or bandmix" 2>&1 | tail -3
