import sys, numpy as np
sys.path.insert(0, '.')
from allocnet_b200 import api, synth
from allocnet_b200.params import default_params
which = sys.argv[1]
if which == "eval4":
    prm = default_params(4); pb = synth.make_problems(8, N=8, K=16, S=4)
    mb = api.MincoBatch(prm); mb.set_problems(pb); f, g = mb.evaluate(pb.x0()); print(f[:4])
elif which == "opt3":
    prm = default_params(3, max_iterations=int(sys.argv[2]) if len(sys.argv) > 2 else 1000); pb = synth.make_problems(8, N=8, K=16, S=3)
    mb = api.MincoBatch(prm); mb.set_problems(pb); r = mb.optimize(pb.x0()); print(r["f"][:4], r["status"], r["evals"])
elif which == "nearopt":
    from oracle.pyoracle import Oracle
    orc = Oracle()
    prm = default_params(3); pb = synth.make_problems(512, N=8, K=16, S=3)
    mb = api.MincoBatch(prm); mb.set_problems(pb); res = mb.optimize(pb.x0())
    f, g = mb.evaluate(res["x"]); fo, go = orc.cost_batch(prm, pb, res["x"], nthreads=8)
    ad = np.abs(g-go).max(axis=1); gn = np.abs(go).max(axis=1)
    print("abs diff pct", np.percentile(ad,[50,90,99,100]))
    print("gnorm pct", np.percentile(gn,[0,50,90,99,100]))
    print("rel pct", np.percentile(ad/gn,[50,90,99,100]))
    # oracle vs itself under perturbation of x by 1 ulp-ish (conditioning of the hinge gradient)
    xp = res["x"]*(1+1e-15*np.sign(np.random.default_rng(0).normal(size=res["x"].shape)))
    fo2, go2 = orc.cost_batch(prm, pb, xp, nthreads=8)
    ad2 = np.abs(go2-go).max(axis=1)
    print("oracle self-sensitivity to 1e-15 rel perturbation of x: abs diff pct", np.percentile(ad2,[50,90,99,100]))
    print("status", np.unique(res["status"], return_counts=True), "evals mean", res["evals"].mean(), "iters", res["iters"].mean())
