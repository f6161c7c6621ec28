#!/bin/bash
# row kernel iteration: tests under a hard timeout, per-layer timings at 240 tiles, short bench line
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_conv_rows.py -q -x --timeout 120 > gpurun_out/rows_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/rows_tests.log
PYTHONPATH=. timeout -s KILL 300 python tools/rows_bench.py 240 5 2>&1 | tee gpurun_out/rows_bench.log | tail -8
if [ "$1" = "bench" ]; then
timeout -s KILL 500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/rows_bench_line.json 2> gpurun_out/rows_bench_line.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/rows_bench_line.json')); print(d['summary']); print(d['roofline']['step_breakdown_ms']); print(d['roofline']['frac'], d['sustained']['value'])"
fi
