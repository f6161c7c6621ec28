#!/bin/bash
# where the row kernel's time goes: probe switches of the -DUNCL_PROBES build (results are wrong while set)
export PYTHONPATH=. UNCL_LIB=$PWD/uncltmo_b200/libuncltmo_b200_probes.so
for pr in 0 1 2 4 3 6 7; do
  echo "== UNCL_RW_PROBE=$pr (1 no TMA traffic, 2 no epilogue work, 4 one MMA per row)"
  UNCL_RW_PROBE=$pr timeout -s KILL 120 python tools/rows_bench.py 240 5 2>&1 | grep -v "^sum" | awk '{print $1, $8, $9, $10, $11, $12}'
done
