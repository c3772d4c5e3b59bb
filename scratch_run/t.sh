timeout 900 python -m pytest tests -m gpu -x -q -k "not pipe and not cli and not packaging and not dropin" 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-e2e --no-others"
for w in c3base c3best c2best c4i32; do echo -n "$w: "; timeout 300 $B --workload $w --steps 3 --warmup 3 2>/dev/null | grep -o '"decode_ms": [0-9.]*'; done
for t in 4096 1024; do echo -n "c2 $t: "; timeout 300 $B --workload c2 --tiles $t --steps 5 --warmup 3 2>/dev/null | grep -o '"decode_ms": [0-9.]*'; done
