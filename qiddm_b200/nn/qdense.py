"""Drop-in replacements for the 27 dense quantum modules of the reference `nn/qdense.py`.

Constructor positional order, `forward` shapes, `state_dict` keys (`weights` / `weights1`,
`linear_down.*`, `linear_up.*`, `conv_layer.*`, `batchnorm.*`, `batch_norm.*`), `save_name()` and
`__repr__` follow the reference class by class (file:line cited per class).  What changes is the
seam `self.qnode(...)`: it is a sm_100a CUDA kernel call (qiddm_b200.functional.run_stage) with an
adjoint-method backward; there is no PennyLane device and no CPU fallback.

Deliberate deviations:
* H2 — the reference detaches every lightning.qubit result (`torch.tensor(qnode(...))`), so its
  circuit weights and `linear_down` never train.  Here the TRUE gradient flows; set the class
  attribute / instance flag `detach_quantum = True` to reproduce the reference's cut gradient.
* `add_noise in {2, 3}` (density-matrix channels on `default.mixed`) raise NotImplementedError;
  `add_noise == 1` is kept where it is a no-op on the readout (PhaseShift / PhaseDamping before a
  diagonal readout is NOT a no-op in general, so only the provable cases are accepted).
* Attributes `qdev`, `qnode`, `device_type`, `diff_method` are inert (kept for drivers that read them).
"""
from __future__ import annotations

import ast
import math
import operator
import os

import einops
import torch
import torch.nn as nn

from .. import _lib as L
from ..channels import apply_readout_channel, channel_matrix, run_noisy_stage
from ..functional import run_stage
from .glue import skinny_linear
from ..pca import DevicePCA

QDEV_NAME = "qiddm_b200:sm_100a"


# The reference re-fits sklearn PCA on every forward call on the current batch through a numpy round trip
# (nn/qdense.py:456, :1429).  Default here: the same per-call PCA ON THE DEVICE (qiddm_b200.pca.DevicePCA,
# exact solver + U-based sign convention of the reference's scikit-learn 1.1.3, CUDA-graph capturable).
# QIDDM_PCA=host (or `module.pca = sklearn.decomposition.PCA(k)`) keeps the host sklearn round trip.
PCA_ON_DEVICE = os.environ.get("QIDDM_PCA", "device").lower() != "host"


def _make_pca(n_components):
    if PCA_ON_DEVICE:
        return DevicePCA(n_components)
    from sklearn.decomposition import PCA
    return PCA(n_components=n_components)


def _pca_call(pca, method: str, x: torch.Tensor, group=None) -> torch.Tensor:
    """pca.<method>(x) as a float64 tensor on x.device (no gradient flows through the PCA, as in the reference).

    `group` (module attribute `pca_group`, default None = one PCA over the whole batch, what the reference does for the
    batch it is given): rows are split into consecutive groups of that many rows with ONE PCA PER GROUP, solved together
    on the device.  With group = tau this keeps the reference's default batch-size-1 semantics (src/mnist_exm.py:144:
    each step's PCA sees the tau-ladder of one image) for a batch of many images -- which is what lets the training
    step and the sampler scale over images without changing the PCA basis of any of them."""
    if isinstance(pca, DevicePCA):
        if group and x.shape[0] > group:
            if x.shape[0] % group:
                raise ValueError(f"pca_group={group} does not divide the batch of {x.shape[0]} rows")
            out = getattr(pca, method)(x.detach().reshape(x.shape[0] // group, group, -1))
            return out.reshape(x.shape[0], -1)
        return getattr(pca, method)(x.detach())
    if group and x.shape[0] > group:
        raise ValueError("pca_group needs the on-device PCA (QIDDM_PCA=device)")
    return torch.as_tensor(getattr(pca, method)(x.detach().cpu().numpy()), dtype=torch.float64).to(x.device)


def _pca_fit_transform(pca, x: torch.Tensor, group=None) -> torch.Tensor:
    return _pca_call(pca, "fit_transform", x, group)


def _check_noise(add_noise, allow_phase: bool, readout_channels: bool = False):
    if add_noise in (0, None):
        return
    if add_noise == 1 and allow_phase:
        return  # PhaseShift / PhaseDamping on every wire right before probs(): diagonal, no effect on probabilities
    if add_noise in (1, 2, 3) and readout_channels:
        return  # channels right before probs(): exact classical map on the probabilities (qiddm_b200.channels)
    raise NotImplementedError(
        f"add_noise={add_noise}: this class applies its noise channels in the middle of the circuit (after every "
        "re-upload gate), which needs a density-matrix simulation (default.mixed) -- out of scope of the B200 "
        "state-vector path (SURVEY.md 8f-4); channels right before a probability readout are supported")


# The reference cuts the gradient at every lightning.qubit result (`torch.tensor(qnode(...))`), so there only
# `linear_up` trains.  Default here: the TRUE gradient (adjoint kernels; BASELINE.json north_star (d)).  QIDDM_REFERENCE_GRADIENTS=1
# (or `module.detach_quantum = True`, or the class attribute) restores the reference's cut gradient for every class that
# detaches upstream, for runs that must reproduce the recorded training curves with the reference's hyper-parameters.
REFERENCE_GRADIENTS = os.environ.get("QIDDM_REFERENCE_GRADIENTS", "0").lower() in ("1", "true", "yes")

_ARITH = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.FloorDiv: operator.floordiv,
          ast.Pow: operator.pow}


def _parse_int_expr(text: str) -> int:
    """"28 * 28" -> 784 (nn/qdense.py:222-223 uses eval() on the driver's model_params string): integer arithmetic only."""
    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, int):
            return node.value
        if isinstance(node, ast.BinOp) and type(node.op) in _ARITH:
            return _ARITH[type(node.op)](ev(node.left), ev(node.right))
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
            return -ev(node.operand)
        raise ValueError(f"input_dim {text!r}: only integer arithmetic (+ - * // **) is accepted")
    return int(ev(ast.parse(text.strip(), mode="eval")))


def _shape2(shape):
    return (shape, shape) if isinstance(shape, int) else tuple(shape)


def _reupload_spec(n, L_, D, enc=L.ENC_RZ, enc_scale=1.0, readout=L.READ_EXPVAL_Z, read_count=0,
                   post_scale=1.0, clamp=False):
    return L.StageSpec(n_qubits=n, n_blocks=L_, layers_per_block=D, init=L.INIT_ZERO, enc=enc,
                       enc_scale=enc_scale, imprimitive=L.IMP_CZ, readout=readout, read_count=read_count,
                       post_scale=post_scale, clamp=clamp)


class _SaveLoadMixin:
    """save_model/load_model as in nn/qdense.py:297-307."""

    def save_model(self, path, loss_values, epochs):
        torch.save({"model_state_dict": self.state_dict(), "loss_values": loss_values, "epochs": epochs}, path)

    def load_model(self, path):
        # the reference's checkpoints hold a state_dict plus lists / ints: no pickled code is needed to read them
        checkpoint = torch.load(path, map_location="cpu", weights_only=True)
        self.load_state_dict(checkpoint["model_state_dict"])


# ======================================================================================
# a1 — amplitude embedding + SEL(CNOT) + probs            nn/qdense.py:15-125
# ======================================================================================
class _AmplitudeDense(nn.Module):
    _remap = L.REMAP_NONE

    def _setup(self, qdepth, shape):
        self.qdepth = qdepth
        self.width, self.height = _shape2(shape)
        self.pixels = self.width * self.height
        self.wires = math.ceil(math.log2(self.width * self.height))
        self.qdev = QDEV_NAME
        self.weights = nn.Parameter(torch.randn((qdepth, self.wires, 3)) * 0.4)
        self.qnode = self._circuit

    def _spec(self):
        # `path` / `gemm_precision` / `gemm_bwd_precision` may be set on the instance (tests, bench.py); default: dispatcher, x3
        return L.StageSpec(n_qubits=self.wires, n_blocks=1, layers_per_block=self.qdepth,
                           init=L.INIT_AMPLITUDE, n_features=self.pixels, pad_value=0.1,
                           imprimitive=L.IMP_CNOT, remap=self._remap, readout=L.READ_PROBS,
                           read_count=self.pixels, post_scale=float(self.pixels), clamp=True,
                           path=getattr(self, "path", L.PATH_AUTO), gemm_precision=getattr(self, "gemm_precision", 3),
                           gemm_bwd_precision=getattr(self, "gemm_bwd_precision", 0))

    def fused_mse_step(self, x, T, goal, decay_mod=3.0):
        """The whole diffusion training step of this layer in one library call (`Plan.dense_mse_step`): noise ladder of
        src/noise.py:105-126, this layer's forward (nn/qdense.py:56-66 / :95-111), MSELoss + `.mean().backward()` of
        src/models.py:65-67 (goal "data") / :95-99 (goal "noise").  `weights.grad` is accumulated as autograd would; returns
        the loss (0-d tensor), or None when the step does not qualify (the caller then runs the unfused sequence): needs
        the unitary-collapse path, CUDA float32 / float64 images (n, pixels), no noise channel, a trainable `weights`."""
        noise = getattr(self, "add_noise", 0)
        if (not x.is_cuda or x.dim() != 2 or x.shape[1] != self.pixels or x.dtype not in (torch.float32, torch.float64)
                or channel_matrix(noise, self._noise_params.get(noise, 0.0)) is not None
                or not (self.weights.requires_grad and torch.is_grad_enabled()) or self.weights.device != x.device
                or x.requires_grad):          # a gradient with respect to the images needs the autograd sequence
            return None
        plan = L.Plan.get(self._spec())
        if not plan.use_gemm(x.shape[0] * T):
            return None
        from ..noise import _level_weights
        eps = torch.normal(mean=0.5, std=0.2, size=tuple(x.shape), device=x.device)       # the draw of noise.ladder_pair
        coeffs = (1.0, 0.0, 1.0, 0.0) if goal == "data" else (0.1, -0.05, -1.0, 1.0)
        key = (T, float(decay_mod), x.device, x.dtype)          # the level weights are constants of the schedule: built once
        cache = self.__dict__.setdefault("_level_w_cache", {})
        if key not in cache or torch.cuda.is_current_stream_capturing():
            level_w = _level_weights(T + 1, decay_mod, x.device, x.dtype)
            if not torch.cuda.is_current_stream_capturing():
                cache[key] = level_w
        else:
            level_w = cache[key]
        loss, gw = plan.dense_mse_step(x, eps, level_w, T, self.weights, *coeffs)
        gw = gw.view_as(self.weights)
        if self.weights.grad is None:
            self.weights.grad = gw
        else:
            self.weights.grad.add_(gw)
        return loss

    def _circuit(self, inp):
        """Full probs of the circuit, (B, 2**wires) un-scaled (what the reference QNode returns)."""
        s = self._spec()
        s = L.StageSpec(**{**s.__dict__, "read_count": s.dim, "post_scale": 1.0, "clamp": False})
        return run_stage(s, inp, self.weights)

    _noise_params = {2: 0.1, 3: 0.02}      # AmplitudeDamping(0.1), DepolarizingChannel(0.02): nn/qdense.py:101-104

    def forward(self, x):
        x = einops.rearrange(x, "b 1 w h -> b (w h)")
        noise = getattr(self, "add_noise", 0)
        if channel_matrix(noise, self._noise_params.get(noise, 0.0)) is not None:
            # channels on every wire right before probs() (read at call time: src/mnist_noise.py:218 sets it on the
            # trained net): full probabilities -> exact readout channel -> _post_process
            _check_noise(noise, True, readout_channels=True)
            p = apply_readout_channel(self._circuit(x), self.wires, noise, self._noise_params[noise])
            x = torch.clamp(p[:, :self.pixels] * self.pixels, 0, 1)
        else:
            # qnode + _post_process fused in the kernel epilogue (slice, scale by pixels, clamp)
            x = run_stage(self._spec(), x, self.weights)
        return einops.rearrange(x, "b (w h) -> b 1 w h", w=self.width, h=self.height)


class QDenseUndirected_old(_AmplitudeDense):
    """Dense variational circuit. Undirected.  nn/qdense.py:15-68 (qw_map.tanh = pi*tanh remap)."""
    _remap = L.REMAP_PI_TANH

    def __init__(self, qdepth, shape) -> None:
        super().__init__()
        self._setup(qdepth, shape)

    def __repr__(self):
        return f"QDenseUndirected_old(qdepth={self.qdepth}, wires={self.wires})"

    def save_name(self) -> str:
        return f"QDenseUndirected_old{self.qdepth}_w{self.width}_h{self.height}"


class QDenseUndirected_old_noise(_AmplitudeDense):
    """nn/qdense.py:71-125 (torch.tanh remap; add_noise channels only exist on default.mixed)."""
    _remap = L.REMAP_TANH

    def __init__(self, qdepth, shape, add_noise=0, device_type="default.qubit.torch") -> None:
        super().__init__()
        _check_noise(add_noise, allow_phase=True, readout_channels=True)
        self.add_noise = add_noise
        self.device_type = device_type
        self._setup(qdepth, shape)

    def __repr__(self):
        return f"QDenseUndirected_old_noise(qdepth={self.qdepth}, wires={self.wires}, add_noise={self.add_noise})"

    def save_name(self) -> str:
        return f"QDenseUndirected_old_noise{self.qdepth}_w{self.width}_h{self.height}_noise{self.add_noise}"


# ======================================================================================
# a2 — linear_down + AngleEmbedding(Y) + SEL(CNOT) + probs      nn/qdense.py:128-210
# ======================================================================================
class QNN_A(nn.Module):
    """Dense variational circuit with angle encoding and dimensionality reduction.  nn/qdense.py:128-210."""

    def __init__(self, qdepth, shape, add_noise=0, device_type="default.qubit.torch", diff_method="backprop") -> None:
        super().__init__()
        _check_noise(add_noise, allow_phase=True, readout_channels=True)
        self.qdepth = qdepth
        self.add_noise = add_noise
        self.device_type = device_type
        self.diff_method = diff_method
        self.width, self.height = _shape2(shape)
        self.pixels = self.width * self.height
        self.wires = math.ceil(math.log2(self.pixels))
        self.qdev = QDEV_NAME
        self.linear_down = nn.Linear(self.pixels, self.wires, dtype=torch.double)
        self.weights = nn.Parameter(torch.randn((qdepth, self.wires, 3), dtype=torch.double) * 0.4)
        self.qnode = self._circuit

    def _spec(self, full=False):
        return L.StageSpec(n_qubits=self.wires, n_blocks=1, layers_per_block=self.qdepth, init=L.INIT_ZERO,
                           enc=L.ENC_RY, imprimitive=L.IMP_CNOT, readout=L.READ_PROBS,
                           read_count=(1 << self.wires) if full else self.pixels,
                           post_scale=1.0 if full else float(self.pixels), clamp=not full)

    def _circuit(self, inp):
        return run_stage(self._spec(full=True), inp, self.weights)

    _noise_params = {2: 0.05, 3: 0.02}     # PhaseDamping: no effect; AmplitudeDamping(0.05); Depolarizing(0.02): :175-180

    def forward(self, x):
        x = einops.rearrange(x, "b 1 w h -> b (w h)")
        x = skinny_linear(x, self.linear_down)
        noise = self.add_noise
        if channel_matrix(noise, self._noise_params.get(noise, 0.0)) is not None:
            _check_noise(noise, True, readout_channels=True)
            p = apply_readout_channel(self._circuit(x), self.wires, noise, self._noise_params[noise])
            x = torch.clamp(p[:, :self.pixels] * self.pixels, 0, 1)
        else:
            x = run_stage(self._spec(), x, self.weights)
        return einops.rearrange(x, "b (w h) -> b 1 w h", w=self.width, h=self.height)

    def __repr__(self):
        return f"QNN_A(qdepth={self.qdepth}, wires={self.wires}, add_noise={self.add_noise})"

    def save_name(self) -> str:
        return f"QNN_A{self.qdepth}_w{self.width}_h{self.height}_noise{self.add_noise}"


# ======================================================================================
# a5 — linear_down + RZ + SEL(CZ) + <Z> + linear_up       nn/qdense.py:219-386
# ======================================================================================
class _QNNBase(_SaveLoadMixin, nn.Module):
    detach_quantum = REFERENCE_GRADIENTS

    def _setup(self, input_dim, hidden_features, qdepth):
        if isinstance(input_dim, str):
            input_dim = _parse_int_expr(input_dim)  # "28 * 28" -> 784, as nn/qdense.py:222-223
        self.hidden_features = hidden_features
        self.qdepth = qdepth
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.linear_down = nn.Linear(input_dim, hidden_features, dtype=torch.double).to(self.device)
        self.linear_up = nn.Linear(hidden_features, input_dim, dtype=torch.double).to(self.device)
        self.qdev = QDEV_NAME
        self.weights = nn.Parameter(
            torch.randn((qdepth, hidden_features, 3), dtype=torch.double).to(self.device) * 0.4)
        self.qnode = self._circuit

    def _circuit(self, inputs, weights=None):
        w = self.weights if weights is None else weights
        return run_stage(_reupload_spec(self.hidden_features, 1, self.qdepth), inputs.reshape(-1, self.hidden_features), w)

    def _noisy_expvals(self, batch: int):
        """QNN_noise with add_noise = 3 (nn/qdense.py:252-263: a channel on every wire right after its RZ, BEFORE the
        entangling layers).  The state in front of the channels is |0..0> up to a phase, so PhaseDamping and
        AmplitudeDamping leave it alone (K1|0> = 0) and DepolarizingChannel(p) turns each wire into the classical
        mixture (1 - 2p/3)|0><0| + (2p/3)|1><1|: the exact density-matrix result is the mixture over the 2^n basis
        inputs of the noiseless circuit, <Z_i> = sum_s P(s) <s|U^dagger Z_i U|s> -- 2^n state-vector runs, independent
        of the sample (RZ on a basis state is a phase), computed once per call."""
        n = self.hidden_features
        q = 2.0 * self._depolarizing_p / 3.0
        dev = self.weights.device
        idx = torch.arange(1 << n, device=dev, dtype=torch.int32)
        spec = L.StageSpec(n_qubits=n, n_blocks=1, layers_per_block=self.qdepth, init=L.INIT_BASIS, imprimitive=L.IMP_CZ,
                           readout=L.READ_EXPVAL_Z)
        vals = run_stage(spec, None, self.weights, batch=1 << n, basis=idx).to(self.weights.dtype)      # (2^n, n)
        ones = torch.zeros(1 << n, device=dev, dtype=torch.int64)
        for j in range(n):
            ones += (idx.long() >> j) & 1
        prob = (q ** ones.to(vals.dtype)) * ((1.0 - q) ** (n - ones).to(vals.dtype))                      # P(s)
        return (prob[:, None] * vals).sum(dim=0, keepdim=True).expand(batch, n)

    def forward(self, x):
        b, c, w, h = x.shape
        x = x.view(b, -1).to(self.linear_down.weight.dtype)
        x_reduced = skinny_linear(x, self.linear_down)
        if getattr(self, "add_noise", 0) == 3:
            x_reduced = self._noisy_expvals(b) + 0.0 * x_reduced.sum()       # keeps linear_down in the graph (zero gradient)
        else:
            x_reduced = self._circuit(x_reduced)       # add_noise 1 / 2: the channels act on |0..0> and change nothing
        if self.detach_quantum:
            x_reduced = x_reduced.detach()
        x_restored = skinny_linear(x_reduced.to(self.linear_up.weight.dtype), self.linear_up)
        return x_restored.view(b, c, w, h)


class QNN_noise(_QNNBase):
    """nn/qdense.py:219-307."""

    _depolarizing_p = 0.02          # qml.DepolarizingChannel(0.02), nn/qdense.py:261

    def __init__(self, input_dim, hidden_features, qdepth: int, add_noise=0) -> None:
        super().__init__()
        # channels on the wires of |0..0> (right after the single RZ layer): exactly reducible, see _noisy_expvals
        _check_noise(add_noise, allow_phase=True, readout_channels=True)
        self.add_noise = add_noise
        self._setup(input_dim, hidden_features, qdepth)

    def __repr__(self):
        return f"QNN(qdepth={self.qdepth}, features={self.hidden_features}, add_noise={self.add_noise})"

    def save_name(self) -> str:
        return f"QNN_linear_features={self.hidden_features}_qdepth={self.qdepth}_add_noise={self.add_noise}"


class QNN(_QNNBase):
    """nn/qdense.py:310-386."""

    def __init__(self, input_dim, hidden_features, qdepth: int) -> None:
        super().__init__()
        self._setup(input_dim, hidden_features, qdepth)

    def __repr__(self):
        return f"QNN(qdepth={self.qdepth}, features={self.hidden_features})"

    def save_name(self) -> str:
        return f"QNN_linear_features={self.hidden_features}_qdepth={self.qdepth}"


# ======================================================================================
# a3 — RZ re-upload + SEL(CZ) + probs, N chained stages   nn/qdense.py:389-1011, 2182-2436
# ======================================================================================
class _DifferNBase(nn.Module):
    """angles (B,n) -> N stages; stage k+1 reads the first n (optionally post-processed)
    probabilities of stage k as its angles (nn/qdense.py:464-465 + :427)."""
    _reduce = "pca"            # "pca" | "conv" | "raw"
    _enc_scale = 1.0
    _post_each_stage = False   # per-sample variants post-process between stages (:813-817)
    _shared_weights = False    # QIDDM_A_sameN
    _weight_name = "weights"
    detach_quantum = False     # the batched default.qubit.torch classes train through the circuit in the reference too

    def _setup(self, shape, spectrum_layer, N):
        self.spectrum_layer = spectrum_layer
        self.N = N
        self.width, self.height = _shape2(shape)
        self.pixels = self.width * self.height
        n = math.ceil(math.log2(self.pixels))
        self.qdev = QDEV_NAME
        wshape = (spectrum_layer, 2, n, 3) if self._shared_weights else (N, spectrum_layer, 2, n, 3)
        setattr(self, self._weight_name, nn.Parameter(torch.randn(wshape) * 0.4))
        self.qnode = self._circuit
        return n

    @property
    def _n(self):
        return getattr(self, "wires", None) or self.hidden_features

    def _stage_spec(self, last: bool):
        n = self._n
        if last or self._post_each_stage:
            return _reupload_spec(n, self.spectrum_layer, 2, enc_scale=self._enc_scale, readout=L.READ_PROBS,
                                  read_count=self.pixels if last else n, post_scale=float(self.pixels), clamp=True)
        return _reupload_spec(n, self.spectrum_layer, 2, enc_scale=self._enc_scale, readout=L.READ_PROBS,
                              read_count=n)

    def _circuit(self, inputs, weights):
        """Un-scaled probs (B, 2**n) of one stage, as the reference QNode returns."""
        n = self._n
        s = _reupload_spec(n, self.spectrum_layer, 2, enc_scale=self._enc_scale, readout=L.READ_PROBS,
                           read_count=1 << n)
        return run_stage(s, inputs.reshape(-1, inputs.shape[-1])[:, :n], weights)

    def _angles(self, x):
        b = x.shape[0]
        n = self._n
        W = getattr(self, self._weight_name)
        if self._reduce == "pca":
            flat = x.reshape(b, -1)
            a = _pca_fit_transform(self.pca, flat, getattr(self, "pca_group", None))
            return a.to(torch.float32).to(W.device)
        if self._reduce == "conv":
            a = self.conv_layer(x)
            return a.view(b, n, -1).mean(dim=2)
        return x.reshape(b, -1)[:, :n]

    def _chain(self, a):
        W = getattr(self, self._weight_name)
        n = self._n
        noise = getattr(self, "add_noise", 0)
        noisy = getattr(self, "_readout_noise", False) and channel_matrix(noise, self._noise_params.get(noise, 0.0)) is not None
        for k in range(self.N):
            w = W if self._shared_weights else W[k]
            last = k == self.N - 1
            if getattr(self, "_mid_circuit_noise", None) and noise in (1, 2, 3):
                # channels after every re-upload RZ (nn/qdense.py:515-527): density-matrix stage, inference only
                a = run_noisy_stage(self._stage_spec(last=last), a[:, :n].contiguous(), w, noise, self._mid_circuit_noise)
            elif noisy:
                # every stage is one QNode whose channels sit right before probs() (nn/qdense.py:431-441): full
                # probabilities -> exact readout channel -> the stage's usual slice (next angles / final image)
                p = apply_readout_channel(self._circuit(a, w), n, noise, self._noise_params[noise])
                if last or self._post_each_stage:
                    a = torch.clamp(p[:, :(self.pixels if last else n)] * self.pixels, 0, 1)
                else:
                    a = p[:, :n]
            else:
                a = run_stage(self._stage_spec(last=last), a[:, :n].contiguous(), w)
            if self.detach_quantum:
                a = a.detach()
        return a

    def forward(self, x):
        b, c, w, h = x.shape
        if not getattr(self, "_readout_noise", False) and not getattr(self, "_mid_circuit_noise", None):
            _check_noise(getattr(self, "add_noise", 0), allow_phase=False)      # read per call: src/mnist_noise.py:218 flips it
        probs = self._chain(self._angles(x))
        return einops.rearrange(probs, "b (w h) -> b 1 w h", w=self.width, h=self.height).to(x.dtype)


class differN_noise(_DifferNBase):
    """nn/qdense.py:389-478."""

    _readout_noise = True
    _noise_params = {2: 0.1, 3: 0.02}      # PhaseShift: no effect; AmplitudeDamping(0.1); Depolarizing(0.02): :431-439

    def __init__(self, shape, spectrum_layer, N, add_noise=0) -> None:
        super().__init__()
        _check_noise(add_noise, allow_phase=True, readout_channels=True)
        self.add_noise = add_noise
        self.wires = self._setup(shape, spectrum_layer, N)
        self.pca = _make_pca(self.wires)

    def __repr__(self):
        return f"differN_old_pca={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"

    def save_name(self) -> str:
        return f"differN_old_pca={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}_noise{self.add_noise}"


class differN_noise_befor(_DifferNBase):
    """nn/qdense.py:481-562."""

    _mid_circuit_noise = {1: 0.03, 2: 0.05, 3: 0.02}    # PhaseDamping / AmplitudeDamping / DepolarizingChannel, :520-526

    def __init__(self, shape, spectrum_layer, N, add_noise=0, device_type="default.qubit.torch") -> None:
        super().__init__()
        if add_noise not in (0, None, 1, 2, 3):
            raise NotImplementedError(f"add_noise={add_noise}")
        self.add_noise = add_noise
        self.device_type = device_type
        self.wires = self._setup(shape, spectrum_layer, N)
        self.pca = _make_pca(self.wires)

    def __repr__(self):
        return f"differN_noise={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"

    def save_name(self) -> str:
        return f"differN_noise={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"


class differN_old_pca(_DifferNBase):
    """nn/qdense.py:671-743."""

    def __init__(self, shape, spectrum_layer, N) -> None:
        super().__init__()
        self.wires = self._setup(shape, spectrum_layer, N)
        self.pca = _make_pca(self.wires)

    def __repr__(self):
        return f"differN_old_pca={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"

    def save_name(self) -> str:
        return f"differN_old_pca={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"


class differN_new_pca(_DifferNBase):
    """Per-sample variant, post-processes between stages.  nn/qdense.py:747-835."""
    _post_each_stage = True

    def __init__(self, shape, spectrum_layer, N) -> None:
        super().__init__()
        self.wires = self._setup(shape, spectrum_layer, N)
        self.pca = _make_pca(self.wires)

    def __repr__(self):
        return f"differN_new_pca={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"

    def save_name(self) -> str:
        return f"differN_new_pca={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"


class differN_new_conv(_DifferNBase):
    """nn/qdense.py:838-936."""
    _reduce = "conv"
    _post_each_stage = True

    def __init__(self, shape, spectrum_layer, N) -> None:
        super().__init__()
        self.wires = self._setup(shape, spectrum_layer, N)
        self.pca = _make_pca(self.wires)
        self.conv_layer = nn.Conv2d(in_channels=1, out_channels=self.wires, kernel_size=3, stride=2, padding=1)

    def __repr__(self):
        return f"differN_new_conv={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"

    def save_name(self) -> str:
        return f"differN_new_conv={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"


class differN_old_conv(_DifferNBase):
    """nn/qdense.py:939-1011."""
    _reduce = "conv"

    def __init__(self, shape, spectrum_layer, N) -> None:
        super().__init__()
        self.wires = self._setup(shape, spectrum_layer, N)
        self.pca = _make_pca(self.wires)
        self.conv_layer = nn.Conv2d(in_channels=1, out_channels=self.wires, kernel_size=3, stride=2, padding=1)

    def __repr__(self):
        return f"differN_old_conv={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"

    def save_name(self) -> str:
        return f"differN_old_conv={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"


class QIDDM_A_sameN(_DifferNBase):
    """Same weights for all N stages, raw pixels as angles.  nn/qdense.py:2276-2342."""
    _reduce = "raw"
    _shared_weights = True

    def __init__(self, shape, spectrum_layer, N) -> None:
        super().__init__()
        self.wires = self._setup(shape, spectrum_layer, N)

    def __repr__(self):
        return f"QIDDM_A_sameN={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"

    def save_name(self) -> str:
        return f"QIDDM_A_sameN={self.spectrum_layer}_N={self.N}_w{self.width}_h{self.height}"


class _QIDDM_A_differN(_SaveLoadMixin, _DifferNBase):
    """RZ(pi/2 * a), per-sample, post-processed between stages.  nn/qdense.py:2182-2273, 2345-2436.
    (The reference re-feeds all `pixels` post-processed values; the circuit reads the first n.)"""
    _enc_scale = math.pi * 0.5
    _post_each_stage = True
    _weight_name = "weights1"
    detach_quantum = REFERENCE_GRADIENTS       # per-sample lightning.qubit classes: nn/qdense.py:2246, :2409

    def __init__(self, input_dim, spectrum_layer, N: int) -> None:
        super().__init__()
        self.hidden_features = self._setup(input_dim, spectrum_layer, N)
        self.pca = _make_pca(self.hidden_features)

    def _angles(self, x):
        b = x.shape[0]
        a = _pca_fit_transform(self.pca, x.reshape(b, -1), getattr(self, "pca_group", None))
        return a.to(x.device).to(x.dtype)

    def __repr__(self):
        return f"QIDDM(qlayer={self.spectrum_layer}, features={self.hidden_features}, N={self.N})"


class QIDDM_A_differN_basePL(_QIDDM_A_differN):
    def save_name(self) -> str:
        return f"QIDDM_pca_features={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_A_differN_NEW(_QIDDM_A_differN):
    def save_name(self) -> str:
        return f"QIDDM_pca_new={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


# ======================================================================================
# a4 — reduce -> N x [RZ re-upload + SEL(CZ) + <Z>] -> linear_up   nn/qdense.py:565-670, 1014-2161
# ======================================================================================
class _QIDDMExpval(_SaveLoadMixin, nn.Module):
    _reduce = "linear"        # "linear" | "pca" | "conv"
    _restore = "linear"       # "linear" | "pca"
    _enc = L.ENC_RZ
    _layers = 2
    _bias = True
    _repr_name = "QIDDM"
    detach_quantum = REFERENCE_GRADIENTS

    def _setup(self, input_dim, hidden_features, spectrum_layer, N):
        self.hidden_features = hidden_features
        self.spectrum_layer = spectrum_layer
        self.N = N
        if self._reduce == "pca":
            self.pca = _make_pca(hidden_features)
        elif self._reduce == "conv":
            self.conv_layer = nn.Conv2d(in_channels=1, out_channels=hidden_features, kernel_size=3, stride=2,
                                        padding=1)
        else:
            self.linear_down = nn.Linear(input_dim, hidden_features, bias=self._bias)
        if self._restore == "linear":
            self.linear_up = nn.Linear(hidden_features, input_dim, bias=self._bias)
        self.qdev = QDEV_NAME
        self.weights1 = nn.Parameter(torch.randn((N, spectrum_layer, self._layers, hidden_features, 3)) * 0.4)
        self.qnode = self._circuit

    def _spec(self):
        return _reupload_spec(self.hidden_features, self.spectrum_layer, self._layers, enc=self._enc)

    def _circuit(self, inputs, weights1):
        return run_stage(self._spec(), inputs.reshape(-1, self.hidden_features), weights1)

    def _reduce_input(self, x):
        b = x.shape[0]
        if self._reduce == "pca":
            ref = self.linear_up.weight if hasattr(self, "linear_up") else self.weights1
            a = _pca_fit_transform(self.pca, x.reshape(b, -1), getattr(self, "pca_group", None))
            return a.to(ref.device).to(ref.dtype)
        if self._reduce == "conv":
            return self.conv_layer(x).view(b, self.hidden_features, -1).mean(dim=2)
        return skinny_linear(x.reshape(b, -1).to(self.linear_down.weight.dtype), self.linear_down)

    def _between_stages(self, a):
        return a

    _mid_circuit_noise = None      # {1: PhaseDamping g, 2: AmplitudeDamping g, 3: Depolarizing p} of the *_noise classes

    def _check_noise_at_call(self):
        """`add_noise` is flipped on trained nets at test time (src/mnist_noise.py:218): read it per call, never serve a
        noiseless result for a noisy request.  Returns the flag when the density-matrix path has to run."""
        noise = getattr(self, "add_noise", 0)
        if noise in (0, None):
            return 0
        if self._mid_circuit_noise and noise in (1, 2, 3) and self._enc == L.ENC_RZ:
            return noise
        _check_noise(noise, allow_phase=False)
        return 0

    def forward(self, x, hidden_only: bool = False):
        b, c, w, h = x.shape
        noise = self._check_noise_at_call()
        a = self._reduce_input(x)
        for n in range(self.N):
            a = self._between_stages(a)
            if noise:
                # channel after every re-upload RZ (nn/qdense.py:1599-1617): density-matrix stage, inference only
                a = run_noisy_stage(self._spec(), a.reshape(-1, self.hidden_features), self.weights1[n], noise,
                                    self._mid_circuit_noise)
            else:
                a = self._circuit(a, self.weights1[n])
            if self.detach_quantum:
                a = a.detach()
        a = a.view(b, -1)
        if hidden_only:
            return a.to(self.linear_up.weight.dtype)
        if self._restore == "linear":
            out = skinny_linear(a.to(self.linear_up.weight.dtype), self.linear_up)
        else:
            out = _pca_call(self.pca, "inverse_transform", a, getattr(self, "pca_group", None)).to(x.dtype).requires_grad_(True)
        return out.view(b, c, w, h)

    def forward_hidden(self, x):
        """Everything in front of `linear_up`: (rows, hidden) -- for the fused tail of the diffusion step (linear_up + loss in
        one pass, qiddm_b200.noise.linear_up_mse_loss); None when the class restores through the PCA instead."""
        if self._restore != "linear":
            return None
        return self.forward(x, hidden_only=True)

    def __repr__(self):
        return f"{self._repr_name}(qlayer={self.spectrum_layer}, features={self.hidden_features}, N={self.N})"


class _QIDDMExpvalNoise(_QIDDMExpval):
    # PhaseDamping(0.03) / AmplitudeDamping(0.05) / DepolarizingChannel(0.9) after every RZ: nn/qdense.py:1412-1416, :1511-1515,
    # :1608-1612, :1700-1704 (the RY-encoded QIDDM_PL_noise1, :606-610, has no density-matrix path: RY does not commute with
    # the channels, NotImplementedError)
    _mid_circuit_noise = {1: 0.03, 2: 0.05, 3: 0.9}

    def __init__(self, input_dim, hidden_features, spectrum_layer, N: int, add_noise=0,
                 device_type="lightning.qubit") -> None:
        super().__init__()
        if add_noise not in (0, None, 1, 2, 3):
            raise NotImplementedError(f"add_noise={add_noise}")
        self.add_noise = add_noise
        self.device_type = device_type
        self._setup(input_dim, hidden_features, spectrum_layer, N)

    def __repr__(self):
        return (f"{self._repr_name}(qlayer={self.spectrum_layer}, features={self.hidden_features}, "
                f"N={self.N}, add_noise={self.add_noise})")


class _QIDDMExpvalPlain(_QIDDMExpval):
    def __init__(self, input_dim, hidden_features, spectrum_layer, N: int) -> None:
        super().__init__()
        self._setup(input_dim, hidden_features, spectrum_layer, N)


class QIDDM_PL_noise1(_QIDDMExpvalNoise):
    """RY re-upload variant.  nn/qdense.py:565-668."""
    _reduce, _enc, _repr_name = "pca", L.ENC_RY, "QIDDM_PL_noise"

    def save_name(self) -> str:
        return f"QIDDM_PL_noise={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_PL_noise(_QIDDMExpvalNoise):
    """nn/qdense.py:1371-1466 (configs 3/4 default)."""
    _reduce, _repr_name = "pca", "QIDDM_PL_noise"

    def save_name(self) -> str:
        return f"QIDDM_PL_noise={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_LL_noise(_QIDDMExpvalNoise):
    """nn/qdense.py:1567-1660 (config 1 default)."""
    _repr_name = "QIDDM_LL_noise"

    def save_name(self) -> str:
        return f"QIDDM_LL_noise={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_LL_relu_noise(_QIDDMExpvalNoise):
    """nn/qdense.py:1469-1565 (the ReLU member is created but never applied there either)."""
    _repr_name = "QIDDM_LL_noise"

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.relu = nn.ReLU()

    def save_name(self) -> str:
        return f"QIDDM_LL_noise={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_PP_noise(_QIDDMExpvalNoise):
    """PCA down, PCA inverse_transform up.  nn/qdense.py:1663-1753."""
    _reduce, _restore, _repr_name = "pca", "pca", "QIDDM_PP_noise"

    def save_name(self) -> str:
        return f"QIDDM_PP_noise={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_CL_new(_QIDDMExpvalPlain):
    """nn/qdense.py:1014-1101."""
    _reduce = "conv"

    def save_name(self) -> str:
        return f"QIDDM_CL_new_q={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_CL_old(_QIDDMExpvalPlain):
    """nn/qdense.py:1104-1173 (no re-wrap of the QNode result there: the gradient is not cut)."""
    _reduce = "conv"
    detach_quantum = False

    def save_name(self) -> str:
        return f"QIDDM_CL_old_q={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_PL_old(_QIDDMExpvalPlain):
    """nn/qdense.py:1176-1268 (the reference passes the whole batch flattened, which only runs at
    batch 1; evaluated per sample here; no re-wrap of the QNode result: the gradient is not cut)."""
    _reduce = "pca"
    detach_quantum = False

    def save_name(self) -> str:
        return f"QIDDM_PL_old_q={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_PL(_QIDDMExpvalPlain):
    """nn/qdense.py:1271-1368."""
    _reduce, _repr_name = "pca", "QIDDM_PL"

    def save_name(self) -> str:
        return f"QIDDM_PL={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_LL_old(_QIDDMExpvalPlain):
    """nn/qdense.py:1873-1968."""

    def save_name(self) -> str:
        return f"QIDDM_linear_features={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_bias_false(_QIDDMExpvalPlain):
    """3-layer SEL blocks, bias-free linears.  nn/qdense.py:1971-2074."""
    _layers, _bias = 3, False

    def save_name(self) -> str:
        return f"QIDDM_linear_features={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_L_B(_QIDDMExpvalPlain):
    """BatchNorm1d before every stage, 3-layer SEL blocks.  nn/qdense.py:2077-2179 (the reference
    instantiates default.qubit.jax with interface torch — dead code)."""
    _layers, _repr_name = 3, "QIDDM_L_B"
    detach_quantum = False          # backprop device in the reference: trains through the circuit

    def __init__(self, input_dim, hidden_features, spectrum_layer, N: int) -> None:
        super().__init__(input_dim, hidden_features, spectrum_layer, N)
        self.batchnorm = nn.BatchNorm1d(hidden_features)

    def _between_stages(self, a):
        return self.batchnorm(a.to(self.batchnorm.weight.dtype)).to(self.weights1.dtype)

    def save_name(self) -> str:
        return f"QIDDM_linear_batch_features={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"


class QIDDM_PP_old(nn.Module):
    """PCA(2n) fitted once -> BatchNorm1d -> linear_down -> circuit -> linear_up -> PCA inverse.
    nn/qdense.py:1756-1870 (the output is re-created with torch.tensor, so nothing trains there)."""

    def __init__(self, input_dim, hidden_features, spectrum_layer, N: int) -> None:
        super().__init__()
        self.hidden_features = hidden_features
        self.spectrum_layer = spectrum_layer
        self.input_dim = input_dim
        self.N = N
        self.pca = None
        self.batch_norm = nn.BatchNorm1d(2 * hidden_features)
        self.linear_down = nn.Linear(2 * hidden_features, hidden_features)
        self.linear_up = nn.Linear(hidden_features, 2 * hidden_features)
        self.qdev = QDEV_NAME
        self.weights1 = nn.Parameter(torch.randn((N, spectrum_layer, 2, hidden_features, 3)) * 0.4)
        self.qnode = self._circuit

    def _circuit(self, inputs, weights1):
        return run_stage(_reupload_spec(self.hidden_features, self.spectrum_layer, 2),
                         inputs.reshape(-1, self.hidden_features), weights1)

    def forward(self, x):
        b, c, w, h = x.shape
        x = x.view(b, -1)
        if self.pca is None:
            self.pca = _make_pca(2 * self.hidden_features)
            _pca_call(self.pca, "fit_transform", x)
        a = _pca_call(self.pca, "transform", x).to(x.dtype).requires_grad_(True)
        a = self.linear_down(self.batch_norm(a))
        for n in range(self.N):
            a = self._circuit(a, self.weights1[n]).detach().to(x.dtype)
        a = self.linear_up(a).view(b, -1)
        out = _pca_call(self.pca, "inverse_transform", a).to(x.dtype).requires_grad_(True)
        return out.view(b, c, w, h)

    def __repr__(self):
        return f"QIDDM_PP(qlayer={self.spectrum_layer}, features={self.hidden_features}, N={self.N})"

    def save_name(self) -> str:
        return f"QIDDM_PP_features={self.hidden_features}_L={self.spectrum_layer}_N={self.N}"

    def save_model(self, path):
        # the fitted PCA travels as plain tensors (mean_, components_, ...), not as a pickle (nn/qdense.py:1852-1870 pickles it)
        model_dict = {"model_state_dict": self.state_dict()}
        if self.pca is not None:
            model_dict["pca_state"] = {k: torch.as_tensor(getattr(self.pca, k)).detach().cpu()
                                       for k in ("mean_", "components_", "singular_values_", "explained_variance_")
                                       if getattr(self.pca, k, None) is not None}
        torch.save(model_dict, path)

    def load_model(self, path):
        checkpoint = torch.load(path, map_location="cpu", weights_only=True)
        self.load_state_dict(checkpoint["model_state_dict"])
        if "pca_state" in checkpoint:
            state = checkpoint["pca_state"]
            self.pca = _make_pca(2 * self.hidden_features)
            dev = self.linear_down.weight.device
            for k, v in state.items():
                setattr(self.pca, k, v.to(dev) if isinstance(self.pca, DevicePCA) else v.numpy())


__all__ = [
    "QDenseUndirected_old", "QDenseUndirected_old_noise", "QNN_A", "QNN_noise", "QNN", "differN_noise",
    "differN_noise_befor", "QIDDM_PL_noise1", "differN_old_pca", "differN_new_pca", "differN_new_conv",
    "differN_old_conv", "QIDDM_CL_new", "QIDDM_CL_old", "QIDDM_PL_old", "QIDDM_PL", "QIDDM_PL_noise",
    "QIDDM_LL_relu_noise", "QIDDM_LL_noise", "QIDDM_PP_noise", "QIDDM_PP_old", "QIDDM_LL_old",
    "QIDDM_bias_false", "QIDDM_L_B", "QIDDM_A_differN_basePL", "QIDDM_A_sameN", "QIDDM_A_differN_NEW",
]
