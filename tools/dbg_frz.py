import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from allocnet_b200 import api, synth
from allocnet_b200 import params as P
from allocnet_b200.params import default_params
def run(mb, pb, nofrz):
    if nofrz: os.environ["MINCOB_NO_FRZ"] = "1"
    else: os.environ.pop("MINCOB_NO_FRZ", None)
    return mb.optimize(pb.x0())
S, N, K, B = 3, 5, 16, 2000
pb = synth.make_problems(B, N=N, K=K, S=S)
for its in (1, 2, 3, 60):
    mb = api.MincoBatch(default_params(S), device=0)
    prm = default_params(S, flags=P.FLAG_FREEZE_TIMES, mapping=P.MAP_THROUGHPUT, max_iterations=its)
    mb.set_params(prm); mb.set_problems(pb)
    a1 = run(mb, pb, False); b1 = run(mb, pb, True)
    bx = np.nonzero((a1["x"] != b1["x"]).any(axis=1))[0]; bf = np.nonzero(a1["f"] != b1["f"])[0]
    print(os.environ.get("MINCOB_LIBRARY", "default").split("/")[-1], "its", its, "x mismatch", len(bx), "f mismatch", len(bf), "first", bx[:6], bf[:6],
          "max rel dx", float(np.max(np.abs(a1["x"] - b1["x"]) / (np.abs(b1["x"]) + 1e-300))), "evals", a1["evals"][bx[:4]] if len(bx) else "")
    mb.close()
