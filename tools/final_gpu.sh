# end-of-round evidence run (one B200): tests, smoke, bench, bandwidth table, role accounting, ncu full capture of one frame's conv launches
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -8
timeout 280 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
timeout 120 python tools/hbm_kernels.py > gpurun_out/hbm_kernels.txt 2>&1
timeout 120 python tools/conv_roles.py > gpurun_out/conv_roles.txt 2>&1
timeout 60 python tools/convt_bench.py > gpurun_out/convt_bench.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 100 -c 25 -o gpurun_out/conv_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/conv_full.ncu-rep
