"""Drop-in `QConv2d` (reference `nn/qconv.py:8-126`, alias `:307`).

Forward semantics = the INTENDED forward of `_QConv2d_FAST` (the reference lost the
`x = self.qnode(x)` line between `:78` and `:79`): unfold -> +0.1 -> AmplitudeEmbedding(pad_with=0.5,
normalize) -> StronglyEntanglingLayers(pi*tanh(W)) -> probs -> *2^n/2 -> clamp(0,1) -> [::2] ->
[:out_channels].  Two B200 paths behind the same module (chosen per call from the patch count, `self.path`):
gate by gate (qiddm_qconv_forward: unfold, post-processing and NCHW re-layout fused into the gate kernel, adjoint
backward with fused col2im) or unitary collapse (qiddm_qconv_gemm_forward: the SEL block collapsed once per
weight version, every patch one row of a tcgen05 GEMM with the unfold fused into the operand preparation)."""
from __future__ import annotations

import math
import os
import warnings

import torch

from .. import _lib as L
from ..functional import build_unitary, run_qconv, run_qconv_reference_map, run_stage

# `reference_forward=True` (or QIDDM_QCONV_REFERENCE_FORWARD=1): reproduce what the reference's forward LITERALLY computes --
# unfold -> +0.1 -> * F * 0.5 -> clamp -> [::2] -> [:out_channels], no circuit (nn/qconv.py:71-90) -- which is the map every
# QConv checkpoint of the reference was trained through (its `weights` never received a gradient).  Default: the intended
# circuit forward (H1).
REFERENCE_FORWARD_DEFAULT = os.environ.get("QIDDM_QCONV_REFERENCE_FORWARD", "0").lower() in ("1", "true", "yes")


class _QConv2d_FAST(torch.nn.Module):
    """Fastest version of QConv2d.  nn/qconv.py:8-126."""

    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), padding=1, qdepth=2, reference_forward=None):
        super().__init__()
        self.reference_forward = REFERENCE_FORWARD_DEFAULT if reference_forward is None else bool(reference_forward)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
        self.padding = padding if isinstance(padding, tuple) else (padding, padding)
        self.qdepth = qdepth
        wires_for_inp = math.ceil(math.log2(self.kernel_size[0] * self.kernel_size[1] * in_channels))
        wires_for_out = math.ceil(math.log2(out_channels))
        self.wires = max(wires_for_inp, wires_for_out, 1)
        if self.wires > 10:
            warnings.warn(f"Too many wires ({self.wires}). This might cause performance issues.")
        if self.wires > L.MAX_QUBITS:
            raise ValueError(f"QConv2d needs {self.wires} wires; the B200 kernels support <= {L.MAX_QUBITS}")
        weights = torch.rand((qdepth, self.wires, 3), dtype=torch.double) * math.pi - math.pi / 2
        self.weights = torch.nn.Parameter(weights)
        self.qdev = "qiddm_b200:sm_100a"
        self.path = L.PATH_AUTO        # AUTO: unitary-collapse tcgen05 GEMM once patches >= 2 * 2**wires, else gate by gate
        self.qnode = self._circuit
        self.sample_qnode = None
        self.sample_matrix = None

    def _spec(self, full=False):
        n = self.wires
        return L.StageSpec(
            n_qubits=n, n_blocks=1, layers_per_block=self.qdepth, init=L.INIT_AMPLITUDE,
            n_features=self.kernel_size[0] * self.kernel_size[1] * self.in_channels, pad_value=0.5,
            add_offset=0.0 if full else 0.1, imprimitive=L.IMP_CNOT, remap=L.REMAP_PI_TANH,
            readout=L.READ_PROBS, read_count=(1 << n) if full else self.out_channels,
            read_stride=1 if full else 2, post_scale=1.0 if full else 0.5 * (1 << n), clamp=not full,
            path=L.PATH_GATE if full else self.path)

    def _circuit(self, features):
        """(P, F) patch rows -> un-scaled probs (P, 2**wires): what the reference QNode returns (:51-56)."""
        return run_stage(self._spec(full=True), features, self.weights)

    def forward(self, x):
        b, c, h_in, w_in = x.shape
        assert c == self.in_channels, f"Expected {self.in_channels} channels, got {c}"
        if self.reference_forward:
            return run_qconv_reference_map(x, self.kernel_size, self.padding, self.out_channels)
        return run_qconv(self._spec(), x, self.weights, self.kernel_size, self.padding)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        if prefix + "weights" in state_dict and not self.reference_forward:
            warnings.warn(
                "QConv2d: loading circuit weights into the circuit forward (H1).  Checkpoints written by the reference were "
                "trained through its literal forward, which never calls the circuit (nn/qconv.py:71-90): their `weights` are "
                "the untrained initial draw.  Pass reference_forward=True (or QIDDM_QCONV_REFERENCE_FORWARD=1) to reproduce "
                "the reference's outputs for such a checkpoint.", stacklevel=2)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def __repr__(self):
        return (f"QConv2d({self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, "
                f"padding={self.padding}, wires={self.wires})")

    def train(self, mode=True):
        """Eval mode collapses the SEL block into one unitary (nn/qconv.py:92-126); kept for API parity
        (`sample_matrix`), built by the gate kernel on the 2**n basis states."""
        super().train(mode)
        if not mode and self.sample_matrix is None and self.weights.is_cuda:
            self.sample_matrix = build_unitary(self._spec(), self.weights)
            self.sample_qnode = self._circuit
        if mode:
            self.sample_qnode = None
            self.sample_matrix = None
        return self


QConv2d = _QConv2d_FAST
