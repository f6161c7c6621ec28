#!/bin/bash
# first GPU pass of round 2: full gpu test suite, the new parity tests with their printed deviations, sanitizer logs, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_train_parity.py --deselect "tests/test_gpu_frame.py::test_bf16_full_frame_matches_oracle" > gpurun_out/pytest_old.log 2>&1; echo "old suite rc=$?" 
tail -5 gpurun_out/pytest_old.log
python -m pytest tests/test_gpu_train_parity.py tests/test_gpu_frame.py -m gpu -q -s -k "parity or bf16_full_frame or mixed or reference_call" > gpurun_out/pytest_new.log 2>&1; echo "new tests rc=$?"
grep -E "mixed step|bf16 frame|mixed gradients|passed|failed|Error|assert" gpurun_out/pytest_new.log | head -60
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitize_run.py > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize_run.py frame > gpurun_out/sanitizer_racecheck_frame.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_frame.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_a.json
