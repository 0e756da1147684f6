"""Turn `ncu -i X.ncu-rep --page raw --csv` into the short per-kernel summary committed under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "<command that was profiled>" > profiles/NAME.txt"""
import csv, io, subprocess, sys
rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
print(f"# ncu --set full --clock-control none --import-source on ; command: {cmd}")
print(f"# report: {rep} (kept in gpurun_out/, not committed)")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"\n== {name[:110]}")
    for w in WANT + stalls:
        if w in hdr:
            i = hdr.index(w)
            v = r[i]
            if w in stalls:
                try:
                    if float(v) < 0.05:
                        continue
                except ValueError:
                    pass
            print(f"{w:92s} {units[i]:18s} {v}")
