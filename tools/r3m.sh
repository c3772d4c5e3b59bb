timeout 900 python -m pytest tests -m gpu -x -q -k "not pipe and not cli and not packaging and not dropin" 2>&1 | tail -3
B="python bench.py --no-cpu-baseline --no-e2e --no-others"
for w in c2best c3best; do echo -n "$w: "; timeout 300 $B --workload $w --steps 3 --warmup 3 2>/dev/null | grep -o '"encode_ms": [0-9.]*'; done
timeout 300 python tools/big_tiles.py 7 | head -2
