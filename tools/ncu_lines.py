#!/usr/bin/env python
"""Per-source-line hot spots of an ncu capture taken with --import-source on (kernels built -lineinfo).
   python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter())
text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name", "Kernel Name", "File Name"): continue
    if hdr is None or cur is None or len(r) < len(hdr): continue
    if r[2] != "-": continue    # sass rows carry an address; keep the per-line aggregate rows
    d = dict(zip(hdr, r))
    key = (cur, int(r[0]))
    text[key] = r[1]
    agg[key]["samples"] += int(d["# Samples"] or 0)
    agg[key]["inst"] += int(d["Instructions Executed"] or 0)
    for k in ("stall_long_sb", "stall_wait", "stall_short_sb", "stall_no_inst", "stall_math", "stall_branch_resolving", "stall_selected", "stall_not_selected", "stall_lg", "stall_mio", "stall_dispatch"):
        agg[key][k] += int(d.get(k) or 0)
tot = sum(v["samples"] for v in agg.values()); toti = sum(v["inst"] for v in agg.values())
print(f"total samples {tot}, warp instructions {toti}")
byfile = collections.Counter()
for (f, l), v in agg.items(): byfile[f] += v["samples"]
print(byfile.most_common())
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s}  long  wait short noins math  sel  | source")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    s = max(v["samples"], 1)
    print(f"{key[0][:20]+':'+str(key[1]):28s} {100*v['samples']/tot:6.2f} {100*v['inst']/toti:6.2f}  "
          f"{100*v['stall_long_sb']/s:4.0f} {100*v['stall_wait']/s:4.0f} {100*v['stall_short_sb']/s:4.0f} {100*v['stall_no_inst']/s:4.0f} "
          f"{100*v['stall_math']/s:4.0f} {100*v['stall_selected']/s:4.0f}  | {text[key].strip()[:90]}")
