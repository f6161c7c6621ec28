#!/bin/bash
# end-of-round evidence with the row kernel: full bench line (N = 1), compute-sanitizer over the frame path
mkdir -p gpurun_out
timeout -s KILL 420 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench_1gpu.json 2> gpurun_out/r2e_bench_1gpu.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_1gpu.json')); print(d['summary']); print(d['roofline']['frac'], d['roofline']['traffic'], d['e2e'], d['clocks'])"
timeout -s KILL 300 compute-sanitizer --tool memcheck python tools/sanitize_run.py frame > gpurun_out/r2e_sanitizer_memcheck_frame.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2e_sanitizer_memcheck_frame.txt
timeout -s KILL 400 compute-sanitizer --tool racecheck python tools/sanitize_run.py frame > gpurun_out/r2e_sanitizer_racecheck_frame.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2e_sanitizer_racecheck_frame.txt
