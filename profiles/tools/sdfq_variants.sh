#!/bin/bash
# grid SDF query kernel: CTAs/SM (direction kernel : value kernel) x waves experiments "D:V:WAVES ..."
for v in $1; do
  IFS=: read D V Wv <<< "$v"
  DSDF_EXTRA_FLAGS="-DSDFQ_MINB_DIR=$D -DSDFQ_MINB_VAL=$V -DSDFQ_WAVES=$Wv" python -m diffsdfsim_b200.build --force > /dev/null 2>&1
  grep -A3 "sdf_query_grid_kernel" diffsdfsim_b200/csrc/_build/dsdf_pointops.o.ptxas.log | grep -E "registers|spill" | cut -c1-90 | tr '\n' ' '
  echo "== dir $D val $V waves $Wv"
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --strong-total 0 --worlds 256 --sim-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1])['sdf_query']; print('  frac', round(d['frac'],3), 'ms', round(d['avg_launch_ms'],3), 'value-only frac', round(d['value_only']['frac'],3))"
done
