"""Summarise an ncu --set full report (.ncu-rep) per kernel launch: duration, DRAM traffic, pipe utilisation, occupancy."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %")]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("kernel:", r[idx["Kernel Name"]][:140])
    for key, name in want:
        if key in idx:
            print(f"  {name:24s} {r[idx[key]]:>16s} {units[idx[key]]}")
