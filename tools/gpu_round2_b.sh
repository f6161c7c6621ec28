#!/bin/bash
# second GPU pass: new bf16 training path - parity tests with printed deviations, full suite, training bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train_parity.py tests/test_gpu_train_step.py tests/test_gpu_backward.py -m gpu -q -s -x > gpurun_out/pytest_train.log 2>&1; echo "train tests rc=$?"
grep -E "mixed step|mixed gradients|passed|failed|Error|error|assert" gpurun_out/pytest_train.log | cut -c1-1200 | head -40
python -m pytest tests -m gpu -q --deselect tests/test_gpu_train_parity.py --deselect tests/test_gpu_train_step.py --deselect tests/test_gpu_backward.py > gpurun_out/pytest_rest.log 2>&1; echo "rest rc=$?"; tail -3 gpurun_out/pytest_rest.log
python tools/train_bench.py bf16 > gpurun_out/train_bench.log 2>&1; echo "train_bench rc=$?"; tail -40 gpurun_out/train_bench.log
