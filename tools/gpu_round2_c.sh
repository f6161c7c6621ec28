#!/bin/bash
# GPU pass: all gpu tests + full bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "all gpu tests rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_all.log | cut -c1-300
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c.json'))
print(json.dumps(d["summary"]))
print("roofline", {k:d["roofline"][k] for k in ("achieved","peak","frac","frac_of_sustained_peak","avg_launch_ms","share_of_step")})
print("hbm", json.dumps(d["roofline_hbm"]["kernels"]))
print("sustained", json.dumps(d["sustained"]))
print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"]["eager_launch_path"]["value"], d["train"]["gpu_launches"], d["train"]["e2e"])
print("cpu", d["cpu_baseline"], d["train"].get("cpu_baseline"))
PY
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep smoke gpurun_out/smoke.log | cut -c1-200
