set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3c_tests.log 2>&1
tail -8 gpurun_out/r3c_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --no-others"
timeout 300 $B --workload c2 --steps 5 --warmup 3 > gpurun_out/r3c_c2.log 2>&1
timeout 300 $B --workload c3base --steps 3 --warmup 3 > gpurun_out/r3c_c3base.log 2>&1
timeout 300 $B --workload c2best --steps 3 --warmup 3 > gpurun_out/r3c_c2best.log 2>&1
timeout 300 $B --workload c3best --steps 3 --warmup 3 > gpurun_out/r3c_c3best.log 2>&1
timeout 300 $B --workload c4i32 --steps 3 --warmup 3 > gpurun_out/r3c_c4i32.log 2>&1
for f in c2 c3base c2best c3best c4i32; do python -c "
import json,sys
for l in open('gpurun_out/r3c_$f.log'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', 'enc %.2f dec %.2f value %.1f'%(d['encode_ms'],d['decode_ms'],d['value']))
"; done
timeout 300 python tools/big_tiles.py 7 > gpurun_out/r3c_big_best.log 2>&1; cat gpurun_out/r3c_big_best.log
timeout 300 python tools/big_tiles.py 8 > gpurun_out/r3c_big_ftl.log 2>&1; cat gpurun_out/r3c_big_ftl.log
timeout 120 python tools/api_latency.py > gpurun_out/r3c_api.log 2>&1; cat gpurun_out/r3c_api.log
