"""Batched time-allocation inference (allocnet_b200/timealloc.py; SURVEY.md section 8f "next" #1).
(1) with random weights: the functional batched forward equals a per-sample forward built from torch.nn
    modules in the reference's layer order (Conv/Pool/Linear stems, nn.LSTM fed a (1,38) tensor as the
    reference does, stop-token break), so the gate order, pooling and masking are right;
(2) when the reference tree is present (this container): equals the reference's own TorchScript model
    `seq5_tokenthresh0_35_cpu.pt`, run per sample, on random planner-shaped inputs."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from allocnet_b200 import synth, timealloc

REF_MODEL = "/root/reference/src/planner/models/seq5_tokenthresh0_35_cpu.pt"


def _inputs(B, L=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    state = torch.randn(B, 9, 2, generator=g)
    hp = torch.randn(B, 50, 4, L, generator=g)
    for b in range(B):                       # zero-padded tail segments and rows, as the planner produces
        seg = 1 + b % L
        hp[b, :, :, seg:] = 0.0
        hp[b, 10 + (3 * b) % 40:, :, :] = 0.0
    return state, hp


def _module_forward(w, state, hp, L):
    """Per-sample forward out of torch.nn modules, mirroring minsnap_network_conv_lstm.py:55-88,114-187."""
    hs = nn.Sequential(nn.Conv2d(50, 16, 3, 1, 1), nn.ReLU(), nn.MaxPool2d(2, 2), nn.MaxPool2d(2, 2), nn.Flatten(),
                       nn.Linear(w["hpoly_input_module.5.weight"].shape[1], 32))
    ss = nn.Sequential(nn.Conv1d(9, 8, 3, 1, 1), nn.ReLU(), nn.MaxPool1d(2, 2), nn.Flatten(), nn.Linear(8, 6))
    lstm = nn.LSTM(38, 256, 1)
    tf = nn.Linear(256, 1); st = nn.Linear(256, 1)
    hs.load_state_dict({k.split("hpoly_input_module.")[1]: v for k, v in w.items() if k.startswith("hpoly_input_module.")})
    ss.load_state_dict({k.split("state_input_module.")[1]: v for k, v in w.items() if k.startswith("state_input_module.")})
    lstm.load_state_dict({k.split("output_module.")[1]: v for k, v in w.items() if k.startswith("output_module.")})
    tf.load_state_dict({"weight": w["tfs_output_layer.weight"], "bias": w["tfs_output_layer.bias"]})
    st.load_state_dict({"weight": w["stop_token_output_layer.0.weight"], "bias": w["stop_token_output_layer.0.bias"]})
    out = torch.zeros(state.shape[0], L)
    with torch.no_grad():
        for b in range(state.shape[0]):
            u = torch.cat([ss(state[b:b + 1]), hs(hp[b:b + 1])], dim=1)
            h = torch.zeros(1, 256); c = torch.zeros(1, 256)
            for k in range(L):
                o, (h, c) = lstm(u, (h, c))          # (1,38): unbatched length-1 sequence, as in the reference
                out[b, k] = tf(o)[0, 0]
                if torch.sigmoid(st(o))[0, 0] > 0.5:
                    break
    return out


def test_batched_forward_equals_module_forward_random_weights():
    for seed in (0, 1):
        w = timealloc.random_weights(seed=seed)
        w["stop_token_output_layer.0.weight"] *= 400.0
        state, hp = _inputs(24, seed=seed)
        mixed = False
        for bias in np.linspace(-40.0, 40.0, 33):           # find a bias for which the stop token fires at different steps
            w["stop_token_output_layer.0.bias"] = torch.tensor([float(bias)])
            want = _module_forward(w, state, hp, 5)
            if (want == 0).any() and (want[:, 1:] != 0).any():
                mixed = True
                break
        assert mixed
        got = timealloc.forward_batched(w, state, hp)
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=2e-6)


@pytest.mark.skipif(not os.path.exists(REF_MODEL), reason="reference model not present on this box")
def test_batched_forward_equals_reference_torchscript():
    ref = torch.jit.load(REF_MODEL, map_location="cpu")
    w = timealloc.load_weights(REF_MODEL)
    state, hp = _inputs(48, seed=3)
    got = timealloc.forward_batched(w, state, hp).numpy()
    want = np.zeros_like(got)
    for b in range(state.shape[0]):
        want[b] = ref(state[b:b + 1], hp[b:b + 1]).detach().numpy().reshape(-1)[:5]
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-6)
    assert (want == 0).any()


def test_pack_inputs_layout_and_warm_start_shapes():
    pb = synth.make_problems(6, N=8, K=16, S=3)
    st, hp = timealloc.pack_inputs(pb.head, pb.tail, pb.hpolys, pb.hrows, 0, 5)
    assert st.shape == (6, 9, 2) and hp.shape == (6, 50, 4, 5)
    np.testing.assert_allclose(st[:, [0, 3, 6], 0].numpy(), pb.head[:, 0].astype(np.float32))   # px, py, pz of the start
    np.testing.assert_allclose(st[:, [1, 4, 7], 0].numpy(), pb.head[:, 1].astype(np.float32))   # vx, vy, vz
    np.testing.assert_allclose(hp[:, :16, 3, 0].numpy(), -pb.hpolys[:, 0, :, 3].astype(np.float32), rtol=1e-6)  # b = -d
    assert float(hp[:, 16:].abs().max()) == 0.0
    T0 = timealloc.warm_start_durations(timealloc.random_weights(seed=5), pb)
    assert T0.shape == (6, 8) and (T0 > 0).all()


# ---- the reference's trained weights, exported to tests/golden/timealloc_seq5.npz (make_timealloc_fixture.py) -------

def test_fixture_weights_reproduce_reference_outputs_cpu():
    """forward_batched with the exported state_dict == what the reference's TorchScript model answered, sample by
    sample, when the fixture was made (random planner-shaped inputs and inputs packed from the corridor generator)."""
    z = np.load(timealloc.FIXTURE)
    w = timealloc.load_weights_npz()
    assert sum(v.numel() for v in w.values()) == 311656
    got = timealloc.forward_batched(w, torch.from_numpy(z["rand_state"]), torch.from_numpy(z["rand_hpolys"])).numpy()
    np.testing.assert_allclose(got, z["rand_times"], rtol=0, atol=5e-6)
    assert (z["rand_times"] == 0).any() and (z["rand_times"][:, 1] != 0).any()
    pb = synth.make_problems(96, N=5, K=16, S=3)
    st, hp = timealloc.pack_inputs(pb.head, pb.tail, pb.hpolys, pb.hrows, 0, 5)
    got = timealloc.forward_batched(w, st, hp).numpy()
    np.testing.assert_allclose(got, z["synth_times"], rtol=0, atol=5e-6)
    # the torch packing used on the device equals the numpy one
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    st2, hp2 = timealloc.pack_inputs_torch(t(pb.head), t(pb.tail), t(pb.hpolys), t(pb.hrows), 0, 5)
    np.testing.assert_array_equal(st2.numpy(), st.numpy()); np.testing.assert_array_equal(hp2.numpy(), hp.numpy())
    if os.path.exists(REF_MODEL):                                          # the fixture is the model's state_dict
        w2 = timealloc.load_weights(REF_MODEL)
        for k in w2:
            assert torch.equal(w[k], w2[k])


def test_windowed_warm_start_torch_equals_numpy():
    w = timealloc.load_weights_npz()
    pb = synth.make_problems(40, N=16, K=16, S=3)
    want = timealloc.warm_start_durations(w, pb)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    got, acc = timealloc.warm_start_durations_torch(w, t(pb.head), t(pb.tail), t(pb.hpolys), t(pb.hrows), t(pb.q0), t(pb.T0), chunk=16)
    np.testing.assert_allclose(got.numpy(), want, rtol=0, atol=1e-6)      # fp32 net, batch chunking changes the GEMM blocking
    assert acc.shape == (40, 4)


@pytest.mark.gpu
def test_fixture_weights_reproduce_reference_outputs_cuda():
    """Same golden vectors, net on cuda:0 (cuDNN/cuBLAS fp32: 2e-5)."""
    z = np.load(timealloc.FIXTURE)
    dev = torch.device("cuda:0")
    w = timealloc.load_weights_npz(device=dev)
    got = timealloc.forward_batched(w, torch.from_numpy(z["rand_state"]).to(dev), torch.from_numpy(z["rand_hpolys"]).to(dev))
    np.testing.assert_allclose(got.cpu().numpy(), z["rand_times"], rtol=0, atol=2e-5)
    pb = synth.make_problems(96, N=5, K=16, S=3)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    st, hp = timealloc.pack_inputs_torch(t(pb.head), t(pb.tail), t(pb.hpolys), t(pb.hrows), 0, 5)
    got = timealloc.forward_batched(w, st, hp)
    # the stop decision (sigmoid > 0.5) may flip for a sample whose token sits at the threshold: compare where both agree
    gn, zn = got.cpu().numpy(), z["synth_times"]
    same = ((gn != 0) == (zn != 0)).all(axis=1)
    assert same.mean() >= 0.97
    np.testing.assert_allclose(gn[same], zn[same], rtol=0, atol=2e-5)
