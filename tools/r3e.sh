set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "not pipe and not cli and not packaging and not dropin" > gpurun_out/r3e_tests.log 2>&1
tail -4 gpurun_out/r3e_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --no-others"
for t in 4096 2048 1024; do timeout 300 $B --workload c2 --tiles $t --steps 5 --warmup 3 > gpurun_out/r3e_c2_$t.log 2>&1; done
grep -o '"decode_ms": [0-9.]*' gpurun_out/r3e_c2_*.log
timeout 300 $B --workload c3base --steps 3 --warmup 3 > gpurun_out/r3e_c3base.log 2>&1
timeout 300 $B --workload c2best --steps 3 --warmup 3 > gpurun_out/r3e_c2best.log 2>&1
grep -o '"decode_ms": [0-9.]*' gpurun_out/r3e_c3base.log gpurun_out/r3e_c2best.log
