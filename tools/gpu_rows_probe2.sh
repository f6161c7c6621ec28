#!/bin/bash
# per-role cycle accounting of the row kernel (printf of the -DUNCL_PROBES build) per probe mode; UNCL_RW_NORING=1 = slot variant
export PYTHONPATH=. UNCL_LIB=$PWD/uncltmo_b200/libuncltmo_b200_probes.so
for pr in ${MODES:-0 2 4}; do
  echo "== UNCL_RW_PROBE=$pr"
  UNCL_RW_PROBE=$((pr + 8)) timeout -s KILL 120 python tools/rows_bench.py 240 1 2>&1 | grep -v "^sum" | awk '/rows probe/ {print; next} {print $1, $8, $9, $10, $11, $12}' | awk '!seen[$0]++' | tail -40
done
