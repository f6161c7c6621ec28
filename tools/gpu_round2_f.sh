#!/bin/bash
# profiling pass (1 GPU): launch lists, ncu full capture of the conv kernels over one frame and of the weight-gradient
# kernels of one training step, racecheck of the training step
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 100 -c 25 -o gpurun_out/r2_conv_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step.csv \
    python tools/train_step_profile.py > gpurun_out/ncu_train.log 2>&1; echo "train launch list rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:wgrad -c 25 -o gpurun_out/r2_wgrad_full -f \
    python tools/train_step_profile.py > gpurun_out/ncu_wgrad.log 2>&1; echo "wgrad capture rc=$?"
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize_run.py train > gpurun_out/sanitizer_racecheck_train.txt 2>&1; echo "racecheck train rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_train.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 python tools/sanitize_run.py > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_memcheck.txt
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
