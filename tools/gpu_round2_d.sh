#!/bin/bash
# GPU pass: video training path - unit tests, parity, video bench numbers
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_parity.py tests/test_gpu_train_step.py -m gpu -q -s > gpurun_out/pytest_video.log 2>&1; echo "tests rc=$?"
grep -E "mixed step vid|mixed gradients|passed|failed|^FAILED|^ERROR|Error" gpurun_out/pytest_video.log | cut -c1-1500 | head -40
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-s 0 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_d.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/bench_d.json')); print(json.dumps(d['summary'])); print(d['video']['train']['ms_per_step'], d['video']['train']['eager_launch_path'])"
