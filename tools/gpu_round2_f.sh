#!/bin/bash
# profiling pass (1 GPU): launch lists, ncu full capture of the conv kernels over one frame and of the weight-gradient
# kernels of one training step (summarised on the box: the reports are too large to bring back)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_entry.py -m gpu -q > gpurun_out/pytest_entry.log 2>&1; echo "entry tests rc=$?"; tail -3 gpurun_out/pytest_entry.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 100 -c 25 -o /tmp/r2_conv_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
python tools/ncu_extract.py /tmp/r2_conv_full.ncu-rep > gpurun_out/r2_conv_tc_ncu_full.csv; echo "extract rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step.csv \
    python tools/train_step_profile.py > gpurun_out/ncu_train.log 2>&1; echo "train launch list rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad -c 25 -o /tmp/r2_wgrad_full -f \
    python tools/train_step_profile.py > gpurun_out/ncu_wgrad.log 2>&1; echo "wgrad capture rc=$?"
python tools/ncu_extract.py /tmp/r2_wgrad_full.ncu-rep > gpurun_out/r2_wgrad_ncu_full.csv; echo "extract rc=$?"
ls -la gpurun_out/*.csv
