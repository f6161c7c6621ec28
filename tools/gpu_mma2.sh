#!/bin/bash
# GPU pass: two MMA-issuing warps - correctness of the tensor-core conv tests, then per-layer times with one / two issuers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_generator.py tests/test_gpu_pw_tc.py tests/test_gpu_train_kernels.py -m gpu -x -q > gpurun_out/pytest_mma2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_mma2.log
export UNCL_LIB=$PWD/uncltmo_b200/libuncltmo_b200_probes.so
echo "== one issuing warp"; UNCL_MMA_WARPS=1 timeout 120 python tools/profile_conv.py all 5 2>&1 | tee gpurun_out/layers_mma1.txt
echo "== two issuing warps"; timeout 120 python tools/profile_conv.py all 5 2>&1 | tee gpurun_out/layers_mma2.txt
