# end-of-round evidence run (one B200): tests, smoke, bench, bandwidth table, per-layer conv times
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -8
timeout 280 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
timeout 120 python tools/hbm_kernels.py > gpurun_out/hbm_kernels.txt 2>&1
timeout 100 python tools/profile_conv.py all 5 > gpurun_out/layers_final.txt 2>&1
timeout 60 python tools/convt_bench.py > gpurun_out/convt_bench.txt 2>&1
