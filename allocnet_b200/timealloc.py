"""Batched inference of AllocNet's time-allocation network (the warm start of BASELINE.json configs[3]).

The reference ships the net as TorchScript (`src/planner/models/seq5_tokenthresh0_35*.pt`) whose forward is
batch-1 only (`.item()` on the stop token) and capped at `ModelMaxSeg` = 5 segments
(`planner/learning_planner.hpp:174-179, 287-291`; architecture `network/utils/learning/
minsnap_network_conv_lstm.py:37-88, 114-187`).  This module re-implements that forward
from the model's `state_dict` for B problems at once, in plain PyTorch on whatever device the inputs live
(fp32 like the reference; it is a pre-processing step, not the optimizer's inner loop):

    state (B,9,2)  --Conv1d(9->8,k3,p1)-ReLU-MaxPool1d(2)-Flatten-Linear(8->6)-->            se (B,6)
    hpolys (B,50,4,L) --Conv2d(50->16,k3,p1)-ReLU-MaxPool2d(2)-MaxPool2d(2)-Flatten-Linear--> he (B,32)
    u = [se, he] (B,38);  h = c = 0;  for k < L:  one LSTM cell step with the SAME input u;
        t_k = Linear(256->1)(h);  stop_k = sigmoid(Linear(256->1)(h));  record t_k, then stop if stop_k > 0.5
    output (B,L): durations, zero after the stop step.

Weights: `load_weights(path)` reads a TorchScript file of the reference (any of its four .pt models);
`load_weights_npz()` reads the state_dict exported from `seq5_tokenthresh0_35_cpu.pt` by
`tests/golden/make_timealloc_fixture.py` (tests/golden/timealloc_seq5.npz: the reference's trained parameters, used
not rebuilt, SURVEY.md section 2 row 13 -- that file is what travels to a GPU box that has no reference tree).  Input layouts are the planner's:
state channels [px,vx,ax,py,vy,ay,pz,vz,az] x {start, goal}; polytope rows [n, b] with n.p <= b, unit
normals, zero padded to 50 rows and L segments.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn.functional as F

MAX_ROWS = 50


def load_weights(path: str, device="cpu") -> dict:
    """state_dict of the reference's TorchScript model (any of the four .pt files)."""
    m = torch.jit.load(path, map_location="cpu")
    return {k: v.detach().to(device=device, dtype=torch.float32) for k, v in m.state_dict().items()}


FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "timealloc_seq5.npz")


def load_weights_npz(path: str = FIXTURE, device="cpu") -> dict:
    """state_dict exported by tests/golden/make_timealloc_fixture.py (keys `w/<name>`)."""
    z = np.load(path)
    return {k[2:]: torch.from_numpy(z[k]).to(device=device, dtype=torch.float32) for k in z.files if k.startswith("w/")}


def random_weights(seq_len: int = 5, hidden: int = 256, seed: int = 0, device="cpu") -> dict:
    """Same names and shapes as the reference's state_dict, random values (tests, benches without the model)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.1).to(device)
    return {
        "state_input_module.0.weight": r(8, 9, 3), "state_input_module.0.bias": r(8),
        "state_input_module.4.weight": r(6, 8), "state_input_module.4.bias": r(6),
        "hpoly_input_module.0.weight": r(16, 50, 3, 3), "hpoly_input_module.0.bias": r(16),
        "hpoly_input_module.5.weight": r(32, 16 if seq_len == 5 else 32), "hpoly_input_module.5.bias": r(32),
        "output_module.weight_ih_l0": r(4 * hidden, 38), "output_module.weight_hh_l0": r(4 * hidden, hidden),
        "output_module.bias_ih_l0": r(4 * hidden), "output_module.bias_hh_l0": r(4 * hidden),
        "tfs_output_layer.weight": r(1, hidden), "tfs_output_layer.bias": r(1),
        "stop_token_output_layer.0.weight": r(1, hidden), "stop_token_output_layer.0.bias": r(1),
    }


@torch.no_grad()
def forward_batched(w: dict, state: torch.Tensor, hpolys: torch.Tensor, stop_threshold: float = 0.5) -> torch.Tensor:
    """(B,9,2), (B,50,4,L) -> (B,L) durations; rows after a problem's stop step are zero."""
    state = state.float(); hpolys = hpolys.float()
    B, L = state.shape[0], hpolys.shape[3]
    se = F.max_pool1d(F.relu(F.conv1d(state, w["state_input_module.0.weight"], w["state_input_module.0.bias"], padding=1)), 2)
    se = F.linear(se.flatten(1), w["state_input_module.4.weight"], w["state_input_module.4.bias"])
    he = F.relu(F.conv2d(hpolys, w["hpoly_input_module.0.weight"], w["hpoly_input_module.0.bias"], padding=1))
    he = F.max_pool2d(F.max_pool2d(he, 2), 2)
    he = F.linear(he.flatten(1), w["hpoly_input_module.5.weight"], w["hpoly_input_module.5.bias"])
    u = torch.cat([se, he], dim=1)                                       # state first (reference :135)
    H = w["output_module.weight_hh_l0"].shape[1]
    gin = F.linear(u, w["output_module.weight_ih_l0"], w["output_module.bias_ih_l0"]) + w["output_module.bias_hh_l0"]
    h = torch.zeros(B, H, device=u.device); c = torch.zeros(B, H, device=u.device)
    out = torch.zeros(B, L, device=u.device)
    alive = torch.ones(B, dtype=torch.bool, device=u.device)
    for k in range(L):
        gates = gin + F.linear(h, w["output_module.weight_hh_l0"])
        i, f, g, o = gates.chunk(4, dim=1)                               # PyTorch LSTM gate order
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        t = F.linear(h, w["tfs_output_layer.weight"], w["tfs_output_layer.bias"]).squeeze(1)
        stop = torch.sigmoid(F.linear(h, w["stop_token_output_layer.0.weight"], w["stop_token_output_layer.0.bias"])).squeeze(1)
        out[:, k] = torch.where(alive, t, torch.zeros_like(t))           # record t_k, THEN stop
        alive = alive & ~(stop > stop_threshold)
    return out


def pack_inputs(head, tail, hpolys, hrows, seg_from: int = 0, seg_count: int | None = None, L: int = 5):
    """Library layouts (include/mincob.h: head/tail [B][S][3] rows P,V,A; hpolys [B][N][K][4] with n.p + d <= 0)
    -> the planner's tensors for segments seg_from .. seg_from+seg_count-1 (learning_planner.hpp:147-168)."""
    head = np.asarray(head); tail = np.asarray(tail); hpolys = np.asarray(hpolys); hrows = np.asarray(hrows)
    B, N, K = hpolys.shape[0], hpolys.shape[1], hpolys.shape[2]
    seg_count = min(L, N - seg_from) if seg_count is None else seg_count
    state = np.zeros((B, 9, 2), dtype=np.float32)
    for a in range(3):
        for d in range(3):
            state[:, 3 * a + d, 0] = head[:, d, a]
            state[:, 3 * a + d, 1] = tail[:, d, a]
    hp = np.zeros((B, MAX_ROWS, 4, L), dtype=np.float32)
    kk = min(K, MAX_ROWS)
    for s in range(seg_count):
        rows = hpolys[:, seg_from + s, :kk].copy()                       # [B][kk][4]
        nrm = np.linalg.norm(rows[:, :, :3], axis=2, keepdims=True)
        rows = np.where(nrm > 0, rows / np.maximum(nrm, 1e-300), 0.0)
        rows[:, :, 3] *= -1.0                                            # n.p + d <= 0  ->  n.p <= b
        live = np.arange(kk)[None, :] < np.minimum(hrows[:, seg_from + s], kk)[:, None]
        hp[:, :kk, :, s] = np.where(live[:, :, None], rows, 0.0)
    return torch.from_numpy(state), torch.from_numpy(hp)


def warm_start_durations(w: dict, pb, fallback_T0=None, L: int = 5, device="cpu") -> np.ndarray:
    """Initial durations [B][N] for a ProblemBatch from the net; problems (or windows) where the net's answer is
    unusable (duration < 1e-10 on a used segment: the planner rejects those, learning_planner.hpp:181-189)
    keep the fallback (the generator's trapezoid rule).  N > L is handled on consecutive windows of <= L pieces
    whose inner cut states are (waypoint, 0, 0) -- a build-side choice, the reference rejects seg > ModelMaxSeg."""
    B, N = pb.B, pb.N
    T0 = np.array(pb.T0 if fallback_T0 is None else fallback_T0, dtype=np.float64, copy=True)
    wpts = np.concatenate([pb.head[:, :1, :], pb.q0, pb.tail[:, :1, :]], axis=1)   # [B][N+1][3]
    for s0 in range(0, N, L):
        cnt = min(L, N - s0)
        head = np.zeros_like(pb.head[:, :3]); tail = np.zeros_like(pb.tail[:, :3])
        head[:, 0] = wpts[:, s0]; tail[:, 0] = wpts[:, s0 + cnt]
        if s0 == 0:
            head = pb.head[:, :3]
        if s0 + cnt == N:
            tail = pb.tail[:, :3]
        st, hp = pack_inputs(head, tail, pb.hpolys, pb.hrows, s0, cnt, L)
        t = forward_batched(w, st.to(device), hp.to(device)).cpu().numpy().astype(np.float64)[:, :cnt]
        ok = (t >= 1e-10).all(axis=1)
        T0[ok, s0:s0 + cnt] = t[ok]
    return T0


@torch.no_grad()
def pack_inputs_torch(head, tail, hpolys, hrows, seg_from: int = 0, seg_count: int | None = None, L: int = 5):
    """pack_inputs on torch tensors of any device (fp64 library layouts in, fp32 planner tensors out)."""
    B, N, K = hpolys.shape[0], hpolys.shape[1], hpolys.shape[2]
    seg_count = min(L, N - seg_from) if seg_count is None else seg_count
    dev = hpolys.device
    # state[:, 3*a + d, 0] = head[:, d, a]
    state = torch.stack([head[:, :3].transpose(1, 2).reshape(B, 9), tail[:, :3].transpose(1, 2).reshape(B, 9)], dim=2).float()
    hp = torch.zeros(B, MAX_ROWS, 4, L, dtype=torch.float32, device=dev)
    kk = min(K, MAX_ROWS)
    rows = hpolys[:, seg_from:seg_from + seg_count, :kk].clone()                  # [B][s][kk][4]
    nrm = rows[..., :3].norm(dim=3, keepdim=True)
    rows = torch.where(nrm > 0, rows / nrm.clamp_min(1e-300), torch.zeros_like(rows))
    rows[..., 3] *= -1.0
    live = torch.arange(kk, device=dev)[None, None, :] < hrows[:, seg_from:seg_from + seg_count].clamp(max=kk)[:, :, None]
    rows = torch.where(live[..., None], rows, torch.zeros_like(rows))
    hp[:, :kk, :, :seg_count] = rows.permute(0, 2, 3, 1).float()
    return state, hp


@torch.no_grad()
def warm_start_durations_torch(w: dict, head, tail, hpolys, hrows, q0, T0, L: int = 5, chunk: int = 8192):
    """warm_start_durations for torch tensors already on the device (bench.py --config 4): returns (T [B][N] fp64,
    accepted [B][ceil(N/L)] bool -- windows whose net answer the planner would accept, learning_planner.hpp:181-189)."""
    B, N = T0.shape
    T = T0.clone()
    nwin = (N + L - 1) // L
    acc = torch.zeros(B, nwin, dtype=torch.bool, device=T0.device)
    wpts = torch.cat([head[:, :1, :], q0, tail[:, :1, :]], dim=1)                # [B][N+1][3]
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        for wi, s0 in enumerate(range(0, N, L)):
            cnt = min(L, N - s0)
            hd = torch.zeros(hi - lo, 3, 3, dtype=head.dtype, device=head.device); tl = torch.zeros_like(hd)
            hd[:, 0] = wpts[lo:hi, s0]; tl[:, 0] = wpts[lo:hi, s0 + cnt]
            if s0 == 0:
                hd = head[lo:hi, :3]
            if s0 + cnt == N:
                tl = tail[lo:hi, :3]
            st, hp = pack_inputs_torch(hd, tl, hpolys[lo:hi], hrows[lo:hi], s0, cnt, L)
            t = forward_batched(w, st, hp).double()[:, :cnt]
            ok = (t >= 1e-10).all(dim=1)
            T[lo:hi, s0:s0 + cnt] = torch.where(ok[:, None], t, T[lo:hi, s0:s0 + cnt])
            acc[lo:hi, wi] = ok
    return T, acc
