#!/bin/bash
# end-of-round pass B (1 GPU): frame / generator tests with the fused skip default, the full bench line, ncu capture of the
# conv kernels for the DRAM traffic figure
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_generator.py tests/test_gpu_frame.py tests/test_gpu_entry.py -m gpu -q > gpurun_out/pytest_fs.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_fs.log | cut -c1-300
ncu --set full --clock-control none -k regex:conv3x3_tc -s 100 -c 27 -o /tmp/r2c_conv_full -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
python tools/ncu_extract.py /tmp/r2c_conv_full.ncu-rep > gpurun_out/r2c_conv_tc_ncu_full.csv; echo "extract rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2c_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --sustain-s 0 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print(json.dumps(d["summary"]))
print("roofline", {k:d["roofline"][k] for k in ("achieved","peak","frac","frac_of_sustained_peak","avg_launch_ms","share_of_step","launches_per_step")})
print("hbm", json.dumps(d["roofline_hbm"]["kernels"]))
print("sustained", json.dumps(d["sustained"]))
print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"]["eager_launch_path"]["value"], d["train"]["gpu_launches"], d["train"]["e2e"])
print("e2e", d["e2e"], "launches", d["gpu_launches"])
PY
