"""aggregate warp-stall samples of an .ncu-rep by CUDA source line: python scripts/ncu_lines.py rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg = collections.Counter(); text = {}; cur_file = ""
stall = collections.Counter()
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    si = hdr.index("# Samples")
    try: n = int(r[si])
    except ValueError: continue
    if r[0] == "": continue          # SASS rows repeat the samples of their source line
    key = (cur_file, r[0])
    agg[key] += n; text[key] = r[1].strip()[:100]
    for i, h in enumerate(hdr):
        if h.startswith("stall_"):
            try: stall[h] += int(r[i])
            except ValueError: pass
tot = sum(agg.values())
print("total samples", tot)
print("stalls:", ", ".join("%s %.0f%%" % (k, 100 * v / max(tot, 1)) for k, v in stall.most_common(6)))
for key, n in agg.most_common(top):
    print("%5d %5.1f%%  %s:%s  %s" % (n, 100 * n / tot, key[0], key[1], text[key]))
