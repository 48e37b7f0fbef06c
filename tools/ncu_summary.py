"""Summarise .ncu-rep files (raw page) into profiles/ncu_summary.md."""
import csv, subprocess, sys, os, io
WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max"]
out = ["# ncu --set full captures (one kernel each, `--clock-control none`, `-lineinfo`)\n"]
for rep in sys.argv[1:]:
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        out.append(f"## {os.path.basename(rep)}: unreadable\n"); continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    out.append(f"## {os.path.basename(rep)} — `{kname[:90]}`\n\n| metric | value | unit |\n|---|---|---|")
    for i, h in enumerate(hdr):
        if h in WANT or any(h.startswith(w) for w in ("sm__pipe_tensor", "sm__inst_executed_pipe_uniform.sum")):
            out.append(f"| {h} | {vals[i]} | {units[i]} |")
    out.append("")
open("profiles/ncu_summary.md", "w").write("\n".join(out))
print("\n".join(out)[:3000])
