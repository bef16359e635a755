"""Plain PyTorch fp64 reference of MINCO for the differentiable-layer test (tests/test_gpu_parity.py::
test_autograd_layer_against_torch_reference); pinned to the CPU oracle by tests/test_oracle_minco.py::
test_torch_reference_equals_oracle."""
import math

import torch


def torch_dense_minco(S, head, tail, q, T):
    """Plain PyTorch fp64 statement of MINCO for ONE trajectory, differentiable by
    torch.autograd: dense 2S*N x 2S*N system (rows: head conditions, per junction the waypoint + continuity of
    derivatives 0..2S-2, tail conditions), torch.linalg.solve, energy = sum_i int |p^(S)|^2."""
    N = T.shape[0]; D = 2 * S

    def beta(t, d):
        row = [torch.zeros((), dtype=torch.float64, device=T.device)] * D
        for k in range(d, D):
            row[k] = (math.factorial(k) / math.factorial(k - d)) * t ** (k - d)
        return torch.stack(row)
    zero = torch.zeros((), dtype=torch.float64, device=T.device)
    A_rows, b_rows = [], []

    def put(piece_rows, rhs):
        full = torch.zeros(N * D, dtype=torch.float64, device=T.device)
        for i, r in piece_rows:
            full = full + torch.nn.functional.pad(r, (i * D, (N - 1 - i) * D))
        A_rows.append(full); b_rows.append(rhs)
    for d in range(S):
        put([(0, beta(zero, d))], head[d])
    for i in range(N - 1):
        put([(i, beta(T[i], 0))], q[i])
        for d in range(D - 1):
            put([(i, beta(T[i], d)), (i + 1, -beta(zero, d))], torch.zeros(3, dtype=torch.float64, device=T.device))
    for d in range(S):
        put([(N - 1, beta(T[N - 1], d))], tail[d])
    A = torch.stack(A_rows); b = torch.stack(b_rows)
    c = torch.linalg.solve(A, b)                              # [D*N][3], row D*i+k = c_k of piece i
    E = torch.zeros((), dtype=torch.float64, device=T.device)
    for i in range(N):
        for k in range(S, D):
            for l in range(S, D):
                w = (math.factorial(k) / math.factorial(k - S)) * (math.factorial(l) / math.factorial(l - S)) / (k + l - 2 * S + 1)
                E = E + w * T[i] ** (k + l - 2 * S + 1) * (c[D * i + k] * c[D * i + l]).sum()
    return E, c
