# Round 2 evidence: bench lines, launch list, raw ncu pages of the two kernels, sweep. Run under gpurun; copies go to profiles/.
set -x
O=gpurun_out
CMD="python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
python bench.py --steps 10 --warmup 3 > $O/r02_bench_final.json 2> $O/r02_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_final_reference.json 2> $O/r02_bench_ref.err
$CMD > $O/r02_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_kernel|decode_kernel|parse_kernel|finish_kernel|seal_kernel|rle_|stitch|derle" -c 200 --csv --log-file $O/r02_launches_c2_4096.csv $CMD > $O/r02_ncu_launch.log 2>&1
for k in encode_kernel decode_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o $O/r02_full_$k python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-others > $O/r02_ncu_$k.log 2>&1
  ncu -i $O/r02_full_$k.ncu-rep --page raw --csv > $O/r02_ncu_raw_$k.csv 2>/dev/null
  ncu -i $O/r02_full_$k.ncu-rep --page source --csv --print-source sass,cuda > $O/r02_ncu_src_$k.csv 2>/dev/null
  rm -f $O/r02_full_$k.ncu-rep
done
python tools/sweep.py > $O/r02_sweep.md 2> $O/r02_sweep.err
ls -la $O | tail -20
