timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="python bench.py --no-cpu-baseline --no-e2e --no-others"
for t in 4096 1024; do echo -n "c2 $t: "; timeout 300 $B --workload c2 --tiles $t --steps 5 --warmup 3 2>/dev/null | grep -o '"decode_ms": [0-9.]*'; done
echo -n "c2best: "; timeout 300 $B --workload c2best --steps 3 --warmup 3 2>/dev/null | grep -o '"decode_ms": [0-9.]*'
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 2 -c 1 -o gpurun_out/r3j_full_dec $B --workload c2 --steps 1 --warmup 2 > gpurun_out/r3j_ncu.log 2>&1
ncu -i gpurun_out/r3j_full_dec.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r3j_src_dec.csv 2>/dev/null
ncu -i gpurun_out/r3j_full_dec.ncu-rep --page raw --csv > gpurun_out/r3j_raw_dec.csv 2>/dev/null
rm -f gpurun_out/r3j_full_dec.ncu-rep
