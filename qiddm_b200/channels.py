"""Readout noise channels of the `*_noise` probability models (reference: the `add_noise` branches of
nn/qdense.py:98-104 (`QDenseUndirected_old_noise`), :174-180 (`QNN_A`), :431-439 (`differN_noise`), evaluated on
PennyLane's `default.mixed` by src/mnist_noise.py:211-229).

In these three classes the channel is applied to every wire AFTER the last entangling layer and immediately before
`qml.probs`.  A single-qubit channel in front of a computational-basis measurement only changes populations, so the
density-matrix simulation reduces exactly to a classical map on the probability vector, p' = (M x ... x M) p:

* add_noise = 1: PhaseShift(0.05) / PhaseDamping(g) — diagonal Kraus operators, M = I (no effect on probs);
* add_noise = 2: AmplitudeDamping(g), Kraus K0 = diag(1, sqrt(1-g)), K1 = sqrt(g)|0><1|  =>  M = [[1, g], [0, 1-g]];
* add_noise = 3: DepolarizingChannel(q), Kraus sqrt(1-q) I, sqrt(q/3) {X, Y, Z}  =>  bit flip with probability 2q/3.

The classes whose channels sit in the MIDDLE of the circuit (after every RZ re-upload gate: `QIDDM_*_noise` nn/qdense.py:1405-1417,
:1504-1516, :1599-1617, :1693-1705; `differN_noise_befor` :515-527) run a full density-matrix simulation on the device
(`run_noisy_stage` -> qiddm_noisy_forward, csrc/qiddm_dm.cu): inference only, as src/mnist_noise.py:211-229 uses them
(flag flipped on a trained net under `diff.eval()`)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

PHASE, AMPLITUDE_DAMPING, DEPOLARIZING = 1, 2, 3


def channel_matrix(add_noise: int, param: float):
    """2 x 2 column-stochastic matrix (m00, m01, m10, m11) of the readout channel, or None for the identity."""
    if add_noise in (0, None, PHASE):
        return None
    if add_noise == AMPLITUDE_DAMPING:
        return (1.0, param, 0.0, 1.0 - param)
    if add_noise == DEPOLARIZING:
        q = 2.0 * param / 3.0
        return (1.0 - q, q, q, 1.0 - q)
    raise NotImplementedError(f"add_noise={add_noise}")


class _ReadoutChannel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, n, m):
        ctx.n, ctx.m = n, m
        return _launch(probs, n, m)

    @staticmethod
    def backward(ctx, g):
        m00, m01, m10, m11 = ctx.m
        return _launch(g, ctx.n, (m00, m10, m01, m11)), None, None        # transposed map


def _launch(p: torch.Tensor, n: int, m) -> torch.Tensor:
    if not p.is_cuda:
        raise L.QiddmError("readout channels run on CUDA tensors only (no CPU fallback)")
    dt = {torch.float32: L.DTYPE_F32, torch.float64: L.DTYPE_F64}[p.dtype]
    p = p.contiguous()
    out = torch.empty_like(p)
    with torch.cuda.device(p.device):
        L.check(L.load_library().qiddm_readout_channel(L._ptr(p), L._ptr(out), dt, p.shape[0], n, *[float(v) for v in m],
                                                       C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)),
                "qiddm_readout_channel")
    return out


def apply_readout_channel(probs: torch.Tensor, n_qubits: int, add_noise: int, param: float) -> torch.Tensor:
    """(B, 2**n) probabilities after the class's noise channel on every wire (differentiable)."""
    m = channel_matrix(add_noise, param)
    if m is None:
        return probs
    if probs.shape[-1] != 1 << n_qubits:
        raise L.QiddmError(f"expected (B, {1 << n_qubits}) probabilities, got {tuple(probs.shape)}")
    return _ReadoutChannel.apply(probs, n_qubits, m)


def mid_circuit_channel(add_noise: int, params: dict):
    """(f_off, (m00, m01, m10, m11)) of the single-qubit channel the class applies after every re-upload RZ: populations
    (rho00, rho11) -> M (rho00, rho11), coherences *= f_off.  `params` = {1: PhaseDamping g, 2: AmplitudeDamping g, 3: Depolarizing p}."""
    if add_noise == PHASE:
        g = params[1]
        return (1.0 - g) ** 0.5, (1.0, 0.0, 0.0, 1.0)
    if add_noise == AMPLITUDE_DAMPING:
        g = params[2]
        return (1.0 - g) ** 0.5, (1.0, g, 0.0, 1.0 - g)
    if add_noise == DEPOLARIZING:
        p = params[3]
        q = 2.0 * p / 3.0
        return 1.0 - 4.0 * p / 3.0, (1.0 - q, q, q, 1.0 - q)
    raise NotImplementedError(f"add_noise={add_noise}")


def run_noisy_stage(spec: "L.StageSpec", angles: torch.Tensor, weights: torch.Tensor, add_noise: int, params: dict) -> torch.Tensor:
    """One QNode call of a re-upload class with its mid-circuit channels, (B, n) angles -> (B, n_out), on `default.mixed`
    semantics.  Forward only: the reference evaluates these under `diff.eval()` / torch.no_grad()."""
    if torch.is_grad_enabled() and (angles.requires_grad or weights.requires_grad):
        raise NotImplementedError(
            "mid-circuit noise channels are an inference path (src/mnist_noise.py:211-229 flips add_noise on a trained net and "
            "samples under no_grad); there is no density-matrix backward -- wrap the call in torch.no_grad()")
    f_off, m = mid_circuit_channel(add_noise, params)
    out = L.Plan.get(spec).noisy_forward(angles.detach(), weights.detach(), f_off, m)
    return out.to(angles.dtype)
