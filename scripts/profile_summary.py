"""Turns an .ncu-rep capture (brought back in gpurun_out/) into the text summary committed under profiles/:
the 'details' page sections that matter for a roofline argument plus the raw per-launch metrics.
Usage: python scripts/profile_summary.py gpurun_out/x.ncu-rep profiles/x.txt"""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
keep_sections = ("GPU Speed Of Light Throughput", "Launch Statistics", "Occupancy", "Memory Workload Analysis", "Compute Workload Analysis",
                 "Scheduler Statistics", "Warp State Statistics", "Instruction Statistics")
lines, on = [], False
for l in det.splitlines():
    if l.strip().startswith("Section: "):
        on = any(k in l for k in keep_sections)
    if l.startswith("  ") and not l.startswith("    ") and "(" in l and "Context" in l:
        on = True
    if on and not l.strip().startswith(("OPT", "INF")) and len(l) < 200:
        lines.append(l.rstrip())
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
wanted = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
          "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
          "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]}\n\n")
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        f.write(f"## launch: {d.get('Kernel Name', '?')}  grid {d.get('launch__grid_size', '?')} x block {d.get('launch__block_size', '?')}\n")
        for k in wanted:
            if k in d:
                f.write(f"{k:85s} {d[k]:>18s} {units[hdr.index(k)]}\n")
        f.write("\n")
    f.write("\n".join(lines) + "\n")
print("wrote", out)
