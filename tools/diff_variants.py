#!/usr/bin/env python
"""Where do two builds of libmincob.so start to differ?  Runs `optimize` with growing iteration caps under
each library (one subprocess per library: MINCOB_LIBRARY), then reports, per cap, how many problems have a
different x / status / evals.  Usage: python tools/diff_variants.py variants/a.so variants/b.so [B]"""
import os, subprocess, sys, tempfile
import numpy as np
CAPS = [1, 2, 3, 4, 6, 8, 9, 10, 12, 16, 24, 40, 80, 200, 1000]
if len(sys.argv) >= 2 and sys.argv[1] == "--child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from allocnet_b200 import api, synth
    from allocnet_b200.params import default_params
    B = int(sys.argv[3]); N = int(sys.argv[4])
    pb = synth.make_problems(B, N=N, K=16, S=3)
    out = {}
    mb = api.MincoBatch(default_params(3), device=0)
    mb.set_problems(pb)
    f, g = mb.evaluate(pb.x0()); out["f0"] = f; out["g0"] = g
    for cap in CAPS:
        mb.set_params(default_params(3, max_iterations=cap))
        r = mb.optimize(pb.x0())
        out[f"x{cap}"] = r["x"]; out[f"s{cap}"] = r["status"]; out[f"e{cap}"] = r["evals"]; out[f"f{cap}"] = r["f"]
    np.savez(sys.argv[2], **out)
    sys.exit(0)
a, b = sys.argv[1], sys.argv[2]
B = sys.argv[3] if len(sys.argv) > 3 else "4096"
N = sys.argv[4] if len(sys.argv) > 4 else "8"
tmp = tempfile.mkdtemp()
res = []
for i, so in enumerate((a, b)):
    o = os.path.join(tmp, f"r{i}.npz")
    subprocess.run([sys.executable, __file__, "--child", o, B, N], check=True, env=dict(os.environ, MINCOB_LIBRARY=os.path.abspath(so)))
    res.append(np.load(o))
ra, rb = res
print("evaluate: max |df|/|f|", float(np.max(np.abs(ra["f0"] - rb["f0"]) / np.abs(ra["f0"]))), "max |dg|", float(np.max(np.abs(ra["g0"] - rb["g0"]))),
      "bit-equal f:", bool((ra["f0"] == rb["f0"]).all()), "g:", bool((ra["g0"] == rb["g0"]).all()))
for cap in CAPS:
    dx = np.abs(ra[f"x{cap}"] - rb[f"x{cap}"]).max(axis=1)
    nd = int((dx > 0).sum())
    print(f"cap {cap:5d}: problems with different x {nd:6d}  max|dx| {dx.max():.3e}  status differs {int((ra[f's{cap}'] != rb[f's{cap}']).sum())}"
          f"  evals differ {int((ra[f'e{cap}'] != rb[f'e{cap}']).sum())}  mean evals {ra[f'e{cap}'].mean():.2f} / {rb[f'e{cap}'].mean():.2f}")
