#!/bin/bash
# GPU pass: merged kernel with resident weights + row-aligned tiles: correctness, then per-layer times with each switched off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_generator.py tests/test_gpu_train_kernels.py -m gpu -x -q > gpurun_out/pytest_mg2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_mg2.log
export UNCL_LIB=$PWD/uncltmo_b200/libuncltmo_b200_probes.so
export PROFILE_DBG=1
L=d0_1,u1_0,u1_1,u2_0,u3_0
echo "== default"; timeout 120 python tools/profile_conv.py $L 5 2>&1
echo "== flat tiles"; UNCL_MG_FLAT=1 timeout 120 python tools/profile_conv.py $L 5 2>&1
echo "== weights per stage"; UNCL_MG_NO_WRES=1 timeout 120 python tools/profile_conv.py $L 5 2>&1
