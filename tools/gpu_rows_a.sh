#!/bin/bash
# first GPU pass of the row kernel: one small case under a hard timeout (a hung kernel must not hold the box), the test file, per-layer timings
mkdir -p gpurun_out
timeout -s KILL 180 python -m pytest tests/test_gpu_conv_rows.py -x -q -k "test_row_kernel_matches_one_tap_and_fp32 and 32-20-0-3" > gpurun_out/rows_first.log 2>&1; echo "first rc=$?"
tail -5 gpurun_out/rows_first.log
timeout -s KILL 600 python -m pytest tests/test_gpu_conv_rows.py -q --timeout 120 > gpurun_out/rows_tests.log 2>&1; echo "tests rc=$?"
tail -25 gpurun_out/rows_tests.log
timeout -s KILL 300 python tools/rows_bench.py 240 5 > gpurun_out/rows_bench.log 2>&1; echo "bench rc=$?"
cat gpurun_out/rows_bench.log | tail -12
