#!/bin/bash
# ring vs slot accumulators of the one-chunk layers, whole frame path (probes build for both arms: same instrumentation overhead)
export UNCL_LIB=$PWD/uncltmo_b200/libuncltmo_b200_probes.so
for arm in ring slot ring slot; do
  if [ $arm = slot ]; then export UNCL_RW_NORING=1; else unset UNCL_RW_NORING; fi
  timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --sustain-s 3 > gpurun_out/ab_$arm.json 2> gpurun_out/ab_$arm.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_$arm.json')); s=d['summary']; r=d['roofline']['step_breakdown_ms']
print('$arm', s['fps'], s['e2e_fps'], s['sustained_fps'], s['conv_frac_of_burst_peak'], {k:v for k,v in r.items() if 'conv3x3' in k}, d['sustained']['clocks']['sm_mhz'])"
done
