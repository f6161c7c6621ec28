#!/bin/bash
# profiling pass of the round's final state (1 GPU): launch list of the bench command, ncu full capture of the tensor-core
# conv kernels over one frame (summarised on the box), cooperative frame-stage kernels, per-layer clocks
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 100 -c 25 -o /tmp/r2b_conv_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
python tools/ncu_extract.py /tmp/r2b_conv_full.ncu-rep > gpurun_out/r2b_conv_tc_ncu_full.csv; echo "extract rc=$?"
ncu --set full --clock-control none -k regex:"normalise_tiles|blend_select|post_select" -s 6 -c 3 -o /tmp/r2b_frame_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_frame.log 2>&1; echo "frame capture rc=$?"
python tools/ncu_extract.py /tmp/r2b_frame_full.ncu-rep > gpurun_out/r2b_frame_stage_ncu_full.csv; echo "extract rc=$?"
UNCL_LIB=$PWD/uncltmo_b200/libuncltmo_b200_probes.so PROFILE_DBG=1 timeout 120 python tools/profile_conv.py all 5 > gpurun_out/r2b_conv_layers_us.txt 2>&1; echo "layers rc=$?"
UNCL_LIB=$PWD/uncltmo_b200/libuncltmo_b200_probes.so timeout 120 python tools/conv_roles.py > gpurun_out/r2b_conv_tc_roles.txt 2>&1; echo "roles rc=$?"
ls -la gpurun_out/r2b_*
