#!/usr/bin/env python
"""Copy the round record that tools/gpu_round.sh left in gpurun_out/ into profiles/ under a tag and refresh
profiles/optimize_kernel_traffic.json (DRAM bytes and executed fp64 flops of one optimize_kernel launch, read by bench.py).
   python tools/update_profiles.py r01_v12"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "bench.json"), os.path.join(P, f"{tag}_bench.json"))
shutil.copy(os.path.join(G, "bench_reference.json"), os.path.join(P, f"{tag}_bench_reference.json"))
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches.csv"))
for name in ("memcheck.log", "racecheck.log"):
    if os.path.exists(os.path.join(G, name)):
        with open(os.path.join(G, name)) as fh:
            tail = fh.read().strip().split("\n")[-6:]
        with open(os.path.join(P, f"{tag}_compute_sanitizer.txt"), "a" if name == "racecheck.log" else "w") as fh:
            fh.write(f"== compute-sanitizer {name.split('.')[0]} (tools/gpu_sanity.sh), last lines\n" + "\n".join(tail) + "\n")
rep = os.path.join(G, "prof_optimize.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_optimize_kernel_raw.csv"), "w").write(raw)
open(os.path.join(P, f"{tag}_optimize_kernel_summary.txt"), "w").write(
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout)
open(os.path.join(P, f"{tag}_optimize_kernel_hot_lines.txt"), "w").write(
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "60"], capture_output=True, text=True).stdout)
rows = list(csv.reader(raw.splitlines()))
d = dict(zip(rows[0], rows[2]))
f = lambda k: float(d[k].replace(",", ""))
unit = dict(zip(rows[0], rows[1]))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = f("dram__bytes_read.sum") * scale[unit["dram__bytes_read.sum"]]
wr = f("dram__bytes_write.sum") * scale[unit["dram__bytes_write.sum"]]
cyc = f("sm__cycles_elapsed.avg")
per = lambda op: f(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")
flops = (2 * per("dfma") + per("dadd") + per("dmul")) * cyc
bench = json.loads(open(os.path.join(G, "bench.json")).read().strip().split("\n")[-1])
cfg = bench["config"]
evals = bench["mean_evals_per_traj"] * cfg["batch_per_gpu"]
sys.path.insert(0, ROOT)
import bench as _bench
out = {"kernel_source_sha256_16": _bench.sass_fingerprint(), "kernel": d["Kernel Name"], "batch": cfg["batch_per_gpu"], "pieces": cfg["pieces"], "K": cfg["K"], "S": cfg["S"],
       "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
       "fp64_flops_per_launch": flops, "fp64_flops_per_eval": flops / evals,
       "fp64_note": "thread-level, predicated-on: (2*dfma + dadd + dmul) inst/cycle x elapsed cycles = "
                    f"(2*{per('dfma'):.2f} + {per('dadd'):.2f} + {per('dmul'):.2f}) x {cyc:.0f}; {evals:.0f} evaluations in the launch",
       "source": f"profiles/{tag}_optimize_kernel_raw.csv (ncu --set full --clock-control none, one launch of "
                 "`python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e`)"}
json.dump(out, open(os.path.join(P, "optimize_kernel_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
print({k: bench[k] for k in ("value", "ms_per_step")}, bench["e2e"]["value"], bench.get("roofline_fp64"), bench["roofline"]["frac"])
