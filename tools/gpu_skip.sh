#!/bin/bash
# GPU pass: fused skip operators - unit tests, generator parity, per-layer times
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_generator.py -m gpu -x -q -k "fused_skip" > gpurun_out/pytest_skip.log 2>&1; echo "skip tests rc=$?"; tail -15 gpurun_out/pytest_skip.log
timeout 600 python -m pytest tests/test_gpu_generator.py tests/test_gpu_frame.py -m gpu -q > gpurun_out/pytest_skip2.log 2>&1; echo "generator+frame tests rc=$?"; tail -5 gpurun_out/pytest_skip2.log
export UNCL_LIB=$PWD/uncltmo_b200/libuncltmo_b200_probes.so
export PROFILE_DBG=1
timeout 120 python tools/profile_conv.py inc1,inc1n,d0_1,d0_1n,u2_0,u2_0f,u3_0,u3_0f 5 2>&1 | tee gpurun_out/layers_skip.txt
