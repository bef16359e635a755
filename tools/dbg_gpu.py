import sys, numpy as np
sys.path.insert(0, '.')
from allocnet_b200 import api, synth
from allocnet_b200.params import default_params
z = np.load("tests/golden/s3_n16_k16.npz")
pb = synth.make_problems(16, N=16, K=16, S=3)
def show(tag, r):
    print(tag, "f", np.round(r["f"][:5], 3), "st", r["status"][:5], "it", r["iters"][:5], "ev", r["evals"][:5])
print("golden f", np.round(z["opt5000_f"][:5], 3), z["opt5000_iters"][:5])
mb = api.MincoBatch(default_params(3, max_iterations=5000), device=0); mb.set_problems(pb)
show("fresh 5000", mb.optimize(z["x0"]))
mb.set_params(default_params(3, max_iterations=3)); show("cap 3", mb.optimize(z["x0"]))
mb.set_params(default_params(3, max_iterations=5000)); show("after cap3, 5000", mb.optimize(z["x0"]))
pb2 = synth.make_problems(128, N=16, K=16, S=3); mb.set_problems(pb2)
show("B=128", mb.optimize(pb2.x0()))
mb.set_problems(pb); show("B=16 again", mb.optimize(z["x0"]))
for B in (1, 2, 8, 9, 16, 17):
    pbb = synth.make_problems(B, N=16, K=16, S=3); mb.set_problems(pbb); show(f"B={B}", mb.optimize(pbb.x0()))
