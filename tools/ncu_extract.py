"""Per-launch summary of an `ncu --set full` report (run here, where ncu can read the .ncu-rep):
python tools/ncu_extract.py gpurun_out/conv_full.ncu-rep > profiles/<name>.csv"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct2"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_insts"),
        ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "tma_load_bytes"),
        ("launch__grid_size", "grid"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__cycles_elapsed.avg.per_second", "sm_clock"), ("sm__cycles_elapsed.max", "sm_cycles"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lsu_smem_wavefronts"),
        ("smsp__inst_executed_pipe_uniform.sum", "uniform_insts")]
have = [(m, a) for m, a in want if m in col]
out = csv.writer(sys.stdout)
out.writerow(["id", "kernel"] + ["%s [%s]" % (a, units[col[m]]) for m, a in have])
for r in data:
    out.writerow([r[col["ID"]], r[col["Kernel Name"]][:60]] + [r[col[m]] for m, _ in have])
