"""Export the reference's trained time-allocation network to a fixture that can travel to the GPU box.

Runs only where the read-only reference tree is present:

    python tests/golden/make_timealloc_fixture.py     # writes tests/golden/timealloc_seq5.npz

SURVEY.md section 2 row 13: the four TorchScript nets under `src/planner/models/` are USED, not rebuilt.  The planner
loads `seq5_tokenthresh0_35*.pt` (learning_planner.hpp:58-79) and calls it once per plan (:174-179).  This script
  1. loads `/root/reference/src/planner/models/seq5_tokenthresh0_35_cpu.pt` with torch.jit.load (UNMODIFIED) and
     stores its state_dict (311 656 fp32 parameters: the reference's trained artefact, data not source code),
  2. runs THAT TorchScript model, sample by sample as the planner does, on seeded planner-shaped inputs and
     stores inputs and outputs (golden vectors for allocnet_b200/timealloc.py::forward_batched on CPU and cuda),
  3. does the same for inputs packed from the synthetic corridor generator (what bench.py --config 4 feeds it).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_MODEL = "/root/reference/src/planner/models/seq5_tokenthresh0_35_cpu.pt"


def planner_shaped_inputs(B, L=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    state = torch.randn(B, 9, 2, generator=g)
    hp = torch.randn(B, 50, 4, L, generator=g)
    for b in range(B):                       # zero-padded tail segments and rows, as the planner produces
        seg = 1 + b % L
        hp[b, :, :, seg:] = 0.0
        hp[b, 10 + (3 * b) % 40:, :, :] = 0.0
    return state, hp


def main():
    if not os.path.exists(REF_MODEL):
        raise SystemExit("reference tree absent")
    from allocnet_b200 import synth, timealloc
    ref = torch.jit.load(REF_MODEL, map_location="cpu")
    sd = {k: v.detach().cpu().numpy() for k, v in ref.state_dict().items()}
    out = {"w/" + k: v for k, v in sd.items()}

    def run(state, hp):
        y = np.zeros((state.shape[0], 5), dtype=np.float32)
        with torch.no_grad():
            for b in range(state.shape[0]):
                y[b] = ref(state[b:b + 1], hp[b:b + 1]).detach().numpy().reshape(-1)[:5]
        return y
    st, hp = planner_shaped_inputs(64, seed=3)
    out["rand_state"], out["rand_hpolys"], out["rand_times"] = st.numpy(), hp.numpy(), run(st, hp)
    pb = synth.make_problems(96, N=5, K=16, S=3)
    st, hp = timealloc.pack_inputs(pb.head, pb.tail, pb.hpolys, pb.hrows, 0, 5)
    out["synth_times"] = run(st, hp)          # inputs are regenerated from the seeded generator by the test
    path = os.path.join(HERE, "timealloc_seq5.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", sum(v.size for v in sd.values()), "parameters")
    print("stop pattern on synthetic corridors: nonzero counts", np.bincount((out["synth_times"] != 0).sum(axis=1), minlength=6))


if __name__ == "__main__":
    main()
