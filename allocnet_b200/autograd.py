"""Differentiable MINCO layer on the device (SURVEY.md section 8f "next" #4).

The reference trains its time-allocation net through `OsqpLayer` (`network/utils/learning/layers.py:35-151`): one
OSQP solve per sample on the CPU, then a dense `torch.linalg.solve` of the KKT Jacobian inside a gradient hook
(`:120-151`) to get d(solution)/d(times).  With MINCO the map (waypoints q, durations T) -> coefficients c is a linear
solve whose adjoint is `propogateGrad`, so the same derivative is two kernel launches for a whole batch:

    forward   mincob_minco_forward_device    setParameters + getCoeffs + getEnergy + partial gradients
    backward  mincob_minco_propagate_device  propogateGrad:  (dL/dc, dL/dT|partial) -> (dL/dq, dL/dT)

`minco_layer(mb, head, tail, q, T)` returns `(energy [B], coeffs [B][2S*N][3])`, both differentiable with respect to
`q` [B][N-1][3] and `T` [B][N] (fp64 CUDA tensors; head/tail [B][S][3] are constants).  Any loss built on the energy
and/or the coefficients (sampled positions, corridor hinges, time regularisation ...) back-propagates through it.
The kernels are the ones the optimizer uses (`minco_kernel`, csrc/kernels_inst.cu); nothing here runs on the CPU.
"""
from __future__ import annotations

import torch

from .api import MincoBatch


# cudaStreamLegacy ((cudaStream_t)0x1): the explicit handle of the legacy default stream.  torch reports its default
# stream as 0, which the C-ABI reads as "the handle's own (non-blocking) stream" -- kernels there are NOT ordered with
# torch's default-stream work, so q / T could be read before they are written and the outputs before they are filled.
_CUDA_STREAM_LEGACY = 0x1


def _bind_to_torch_stream(mb: MincoBatch, device) -> None:
    s = torch.cuda.current_stream(device).cuda_stream
    mb.set_stream(s if s else _CUDA_STREAM_LEGACY)


class _MincoFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mb: MincoBatch, head, tail, q, T):
        if not (T.is_cuda and T.dtype == torch.float64):
            raise TypeError("minco_layer wants fp64 CUDA tensors")
        B, N = T.shape
        S = mb.S
        head = head.contiguous(); tail = tail.contiguous(); T = T.contiguous()
        q = q.contiguous() if N > 1 else torch.zeros(B, 1, 3, dtype=torch.float64, device=T.device)
        coeffs = torch.empty(B, 2 * S * N, 3, dtype=torch.float64, device=T.device)
        energy = torch.empty(B, dtype=torch.float64, device=T.device)
        gdC = torch.empty_like(coeffs)
        gdT = torch.empty(B, N, dtype=torch.float64, device=T.device)
        _bind_to_torch_stream(mb, T.device)
        mb.minco_forward_device(B, N, head, tail, q, T, coeffs_asc=coeffs, energy=energy, gdC=gdC, gdT=gdT)
        ctx.mb = mb
        ctx.save_for_backward(head, tail, q, T, gdC, gdT)
        return energy, coeffs

    @staticmethod
    def backward(ctx, g_energy, g_coeffs):
        head, tail, q, T, gdC, gdT = ctx.saved_tensors
        mb = ctx.mb
        B, N = T.shape
        # L = L(E(q,T), c(q,T)):  dL/dc|partial = g_E * dE/dc + g_c ;  dL/dT|partial = g_E * dE/dT
        pc = torch.zeros_like(gdC) if g_coeffs is None else g_coeffs.contiguous().clone()
        pt = torch.zeros_like(gdT)
        if g_energy is not None:
            pc += g_energy[:, None, None] * gdC
            pt += g_energy[:, None] * gdT
        gq = torch.zeros(B, max(N - 1, 1), 3, dtype=torch.float64, device=T.device)
        gT = torch.empty(B, N, dtype=torch.float64, device=T.device)
        _bind_to_torch_stream(mb, T.device)
        mb.minco_propagate_device(B, N, head, tail, q, T, pc, pt, gq, gT)
        return None, None, None, (gq if N > 1 else None), gT


def minco_layer(mb: MincoBatch, head: torch.Tensor, tail: torch.Tensor, q: torch.Tensor, T: torch.Tensor):
    """(energy [B], coeffs [B][2S*N][3] ascending powers, row 2S*i+k = c_k of piece i) for B trajectories."""
    return _MincoFunction.apply(mb, head, tail, q, T)
