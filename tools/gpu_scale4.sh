#!/bin/bash
# N = 8 pass: the bench line on all GPUs of the box (weak scaling of independent frames; strong / weak training series)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench8 rc=$?"; tail -3 gpurun_out/bench_4gpu.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/bench_4gpu.json')); print(json.dumps(d['summary'])); print(d['train'].get('weak_16_images_per_gpu')); print(d.get('video',{}).get('train',{}).get('value'))"
