#!/bin/bash
# GPU pass: the SM clock each conv launch actually runs at (clock64 vs globaltimer inside the MMA role), per probe mode
export UNCL_LIB=$PWD/uncltmo_b200/libuncltmo_b200_probes.so
export PROFILE_DBG=1
for f in 0 1 3 7; do echo "== flags $f"; UNCL_PROBE_FLAGS=$f timeout 120 python tools/profile_conv.py inc1,d0_1,d1_1,u0_0,u1_0,u2_0,u3_0,u3_1 5 2>&1; done | tee gpurun_out/clock_modes.txt
