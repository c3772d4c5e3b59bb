ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"encode_kernel|rle_kernel|stitch_kernel|best_resolve" -c 20 --csv --log-file gpurun_out/r3h_launch_big.csv python tools/big_one.py 7 4096 4096 3 0 19 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r3h_launch_big.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
for r in rows[1:]: print(r[ki][:70], r[gi], r[bi], float(r[vi].replace(',',''))/1e6,'ms')
PY
