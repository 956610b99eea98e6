#!/bin/bash
# contact kernel: CTAs/SM (register cap) x candidate capacity experiments: "T:M:CAPK ..."
for v in $1; do
  IFS=: read T M C <<< "$v"
  DSDF_CONTACT_MINBLOCKS=$M DSDF_EXTRA_FLAGS="-DDSDF_CONTACT_THREADS=$T" python -m diffsdfsim_b200.build --force > /dev/null 2>&1
  grep -A2 "contacts_kernel" diffsdfsim_b200/csrc/_build/dsdf_contacts.o.ptxas.log | grep -E "registers|spill" | head -2
  echo "== threads $T minblocks $M capK $C"
  CAPK=$C python profiles/tools/kbench.py 4096 12 2>&1 | grep -E "contacts_detect|single|contacts:"
done
