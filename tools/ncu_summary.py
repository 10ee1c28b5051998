"""Summarise selected raw metrics of an .ncu-rep as markdown.  usage: python tools/ncu_summary.py X.ncu-rep > out.md"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_utchmma.sum" ,
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__inst_executed_op_global_red.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = [hdr.index(w) for w in want if w in hdr]
print(f"# {rep.split('/')[-1]}\n")
for r in data:
    print("## " + r[hdr.index("Kernel Name")][:110].replace("<unnamed>::", ""))
    print("| metric | value | unit |\n|---|---:|---|")
    for i in idx[1:]:
        if r[i] not in ("", "n/a"):
            print(f"| `{hdr[i]}` | {r[i]} | {units[i]} |")
    print()
