"""`ncu --set full` report → the handful of metrics the roofline uses (markdown).
usage: python scripts/ncu_summary.py gpurun_out/prof_wv.ncu-rep > profiles/r01_ncu_wv.md"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
path = sys.argv[1]
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# ncu --set full summary: {path}\n")
name_i = hdr.index("Kernel Name")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print(f"## {r[name_i][:100]}\n")
    print("| metric | value | unit |\n|---|---:|---|")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | {r[i]} | {units[i]} |")
    print()
