#!/bin/bash
# 2-GPU pass: data-parallel parity + the bench line at N = 2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check rc=$?"; grep -E "^dp|Error|error" gpurun_out/dp_check.log | cut -c1-400 | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/bench_2gpu.json')); print(json.dumps(d['summary'])); print(d['train'].get('weak_16_images_per_gpu'))"
