"""print a fixed list of metrics of the first launch in an .ncu-rep (reads `ncu --page raw --csv`)"""
import sys, csv, subprocess
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("==", rep, "launches:", len(data))
    d = data[0]
    ki = hdr.index("Kernel Name")
    print("kernel:", d[ki])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w} = {d[i]} {units[i]}")
