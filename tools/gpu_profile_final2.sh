#!/bin/bash
# final-state captures with the bench's default command (4 frames = 240 tiles per generator call)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2d_launches_bench.csv \
    python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none -k regex:conv3x3_tc -s 54 -c 27 -o /tmp/r2d_conv_full -f \
    python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
python tools/ncu_extract.py /tmp/r2d_conv_full.ncu-rep > gpurun_out/r2d_conv_tc_ncu_full.csv; echo "extract rc=$?"
wc -l gpurun_out/r2d_launches_bench.csv gpurun_out/r2d_conv_tc_ncu_full.csv
