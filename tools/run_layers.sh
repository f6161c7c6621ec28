for f in 0 1 3 7; do echo "== flags $f"; UNCL_PROBE_FLAGS=$f python tools/profile_conv.py d0_1,u1_0,u1_1,u2_0,u3_0 5 2>&1; done
echo "== merged for shallow layers too (UNCL_MERGED_MIN_CI=32)"
for f in 0 3 7; do echo "== flags $f"; UNCL_MERGED_MIN_CI=32 UNCL_PROBE_FLAGS=$f python tools/profile_conv.py inc1,d0_0,u2_1,u3_1 5 2>&1; done
