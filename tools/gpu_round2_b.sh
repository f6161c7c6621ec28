#!/bin/bash
# GPU pass: bf16 training path - kernel unit tests, parity tests with printed deviations, full suite, training bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train_kernels.py -m gpu -q > gpurun_out/pytest_kernels.log 2>&1; echo "kernel tests rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/pytest_kernels.log | head -30
python -m pytest tests/test_gpu_train_parity.py tests/test_gpu_train_step.py tests/test_gpu_backward.py -m gpu -q -s > gpurun_out/pytest_train.log 2>&1; echo "train tests rc=$?"
grep -E "mixed step|mixed gradients|passed|failed|^FAILED|^ERROR|Error" gpurun_out/pytest_train.log | cut -c1-1500 | head -40
python -m pytest tests -m gpu -q --deselect tests/test_gpu_train_parity.py --deselect tests/test_gpu_train_step.py --deselect tests/test_gpu_backward.py --deselect tests/test_gpu_train_kernels.py > gpurun_out/pytest_rest.log 2>&1; echo "rest rc=$?"; tail -3 gpurun_out/pytest_rest.log
python tools/train_bench.py bf16 > gpurun_out/train_bench.log 2>&1; echo "train_bench rc=$?"; head -22 gpurun_out/train_bench.log | cut -c1-200
