"""markdown summary of .ncu-rep files (first launch of each): python scripts/ncu_summary.py title out.md rep1 rep2 ..."""
import sys, csv, subprocess
WANT = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum"]
title, out = sys.argv[1], sys.argv[2]
lines = ["# " + title, ""]
for rep in sys.argv[3:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(raw.splitlines()) if len(r) > 10]
    if len(rows) < 3:
        continue
    hdr, units, d = rows[0], rows[1], rows[2]
    name = d[hdr.index("Kernel Name")].split("(")[0]
    lines += ["## " + name, "", "`%s`" % rep.split("/")[-1], "", "| metric | value |", "|---|---|"]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append("| %s | %s %s |" % (w, d[i], units[i]))
    # top source lines by warp-stall samples
    src = subprocess.run([sys.executable, "scripts/ncu_lines.py", rep, "8"], capture_output=True, text=True).stdout
    lines += ["", "Top source lines by warp-stall samples:", "", "```", src.strip(), "```", ""]
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
