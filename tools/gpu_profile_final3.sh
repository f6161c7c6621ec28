#!/bin/bash
# final-state captures of round 2 with the row kernel (r2e_*): launch list of the bench command, one ncu --set full pass over
# the tensor-core kernels of one generator call (4 frames = 240 tiles), per-layer timings, per-role cycle accounting
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2e_launches_bench.csv \
    python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 56 -c 28 -o /tmp/r2e_conv_full -f \
    python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
python tools/ncu_extract.py /tmp/r2e_conv_full.ncu-rep > gpurun_out/r2e_conv_tc_ncu_full.csv; echo "extract rc=$?"
wc -l gpurun_out/r2e_launches_bench.csv gpurun_out/r2e_conv_tc_ncu_full.csv
PYTHONPATH=. timeout -s KILL 200 python tools/rows_bench.py 240 5 > gpurun_out/r2e_row_kernel_layers.txt 2>&1
PYTHONPATH=. timeout -s KILL 100 python tools/rows_bench.py 60 5 >> gpurun_out/r2e_row_kernel_layers.txt 2>&1
MODES="0 2 4" bash tools/gpu_rows_probe2.sh > gpurun_out/r2e_row_kernel_roles.txt 2>&1
tail -5 gpurun_out/r2e_row_kernel_layers.txt
