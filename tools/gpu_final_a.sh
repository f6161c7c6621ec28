#!/bin/bash
# end-of-round pass A (1 GPU): whole gpu test suite, smoke, sanitizers over the frame path and a training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "all gpu tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_all.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep smoke gpurun_out/smoke.log | cut -c1-200
timeout 1200 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitize_run.py > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize_run.py frame > gpurun_out/sanitizer_racecheck_frame.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_frame.txt
