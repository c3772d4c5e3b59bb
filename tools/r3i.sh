timeout 900 python -m pytest tests -m gpu -x -q -k "rle or best or parts or golden or size_only" 2>&1 | tail -4
timeout 300 python tools/big_tiles.py 7
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_kernel|rle_|stitch_kernel|best_resolve" -c 10 --csv --log-file gpurun_out/r3i_launch_big.csv python tools/big_one.py 7 4096 4096 3 0 19 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r3i_launch_big.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
for r in rows[1:]: print(r[ki][:70], r[gi], r[bi], float(r[vi].replace(',',''))/1e6,'ms')
PY
