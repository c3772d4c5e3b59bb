B="python bench.py --no-cpu-baseline --no-e2e --no-others --workload c2 --steps 4 --warmup 3"
for ub in new old; do for sh in 0 1; do for t in 1024 4096; do
  if [ $ub = old ]; then export QB3CU_DBG_OLDUB=1; else unset QB3CU_DBG_OLDUB; fi
  export QB3CU_DBG_SHARE=$sh
  echo -n "ub=$ub share=$sh tiles=$t: "; timeout 300 $B --tiles $t 2>/dev/null | grep -o '"decode_ms": [0-9.]*'
done; done; done
