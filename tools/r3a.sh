set -x
B="python bench.py --no-cpu-baseline --no-e2e --no-others"
$B --workload c2best --steps 3 --warmup 3 > gpurun_out/r3a_c2best.log 2>&1
$B --workload c3best --steps 3 --warmup 3 > gpurun_out/r3a_c3best.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r3a_launch_c2best.csv $B --workload c2best --steps 1 --warmup 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r3a_launch_c3best.csv $B --workload c3best --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 1 -c 1 -o gpurun_out/r3a_full_c2best $B --workload c2best --steps 1 --warmup 1 > gpurun_out/r3a_ncu_c2best.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 1 -c 1 -o gpurun_out/r3a_full_c3best $B --workload c3best --steps 1 --warmup 1 > gpurun_out/r3a_ncu_c3best.log 2>&1
E="python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-others"
$E > gpurun_out/r3a_e2e_default.log 2>&1
$E --e2e-dec-chunk 128 --e2e-dec-depth 16 > gpurun_out/r3a_e2e_d128x16.log 2>&1
$E --e2e-dec-chunk 512 --e2e-dec-depth 6 > gpurun_out/r3a_e2e_d512x6.log 2>&1
$E --e2e-enc-chunk 256 --e2e-enc-depth 8 > gpurun_out/r3a_e2e_e256x8.log 2>&1
$E --e2e-enc-chunk 64 --e2e-enc-depth 16 --e2e-dec-chunk 128 --e2e-dec-depth 12 > gpurun_out/r3a_e2e_small.log 2>&1
for k in c2best c3best; do
  ncu -i gpurun_out/r3a_full_$k.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/r3a_src_$k.csv 2>/dev/null
  ncu -i gpurun_out/r3a_full_$k.ncu-rep --page raw --csv > gpurun_out/r3a_raw_$k.csv 2>/dev/null
  rm -f gpurun_out/r3a_full_$k.ncu-rep
done
ls -la gpurun_out | tail -20
