#!/bin/bash
# final-state captures of round 2 (r2f_*: row kernel incl. its front mode, 27 tensor-core launches per generator call): launch list of the bench command, one ncu --set full pass over
# the tensor-core kernels of one generator call (4 frames = 240 tiles), per-layer timings, per-role cycle accounting
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2f_launches_bench.csv \
    python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 54 -c 27 -o /tmp/r2f_conv_full -f \
    python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
python tools/ncu_extract.py /tmp/r2f_conv_full.ncu-rep > gpurun_out/r2f_conv_tc_ncu_full.csv; echo "extract rc=$?"
wc -l gpurun_out/r2f_launches_bench.csv gpurun_out/r2f_conv_tc_ncu_full.csv
PYTHONPATH=. timeout -s KILL 200 python tools/rows_bench.py 240 5 > gpurun_out/r2f_row_kernel_layers.txt 2>&1
PYTHONPATH=. timeout -s KILL 100 python tools/rows_bench.py 60 5 >> gpurun_out/r2f_row_kernel_layers.txt 2>&1
tail -5 gpurun_out/r2f_row_kernel_layers.txt
timeout -s KILL 300 compute-sanitizer --tool memcheck python tools/sanitize_run.py frame > gpurun_out/r2f_sanitizer_memcheck_frame.txt 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2f_sanitizer_memcheck_frame.txt
timeout -s KILL 400 compute-sanitizer --tool racecheck python tools/sanitize_run.py frame > gpurun_out/r2f_sanitizer_racecheck_frame.txt 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r2f_sanitizer_racecheck_frame.txt
