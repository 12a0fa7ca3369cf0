"""CPU oracle for the QIDDM quantum-layer hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path
(``qiddm_b200``) never does and fails loudly when its CUDA library is missing.

What this is
------------
A complex128 torch restatement of the state-vector arithmetic that the reference
delegates to third-party simulators that are NOT vendored in ``/root/reference`` and
NOT installable here (no network):

* PennyLane 0.29.0 (``requirements.txt:46``): ``AmplitudeEmbedding``, ``AngleEmbedding``,
  ``RZ``, ``RY``, ``Rot``, ``CNOT``, ``CZ``, ``StronglyEntanglingLayers``, ``probs``,
  ``expval(PauliZ)``, device ``default.qubit.torch``;
* PennyLane-Lightning 0.30.0 (``requirements.txt:47``): device ``lightning.qubit``;
* qW-Map 0.1.2 (``requirements.txt:68``): ``qw_map.tanh`` (= pi*tanh; package source absent, pinned by
  fixture F3, see below).

Parity status: the reference has no tests for this path.  The conventions restated
here (wire 0 = MSB, ``Rot = RZ(omega) RY(theta) RZ(phi)``, SEL ranges, CNOT direction,
CZ, RZ re-upload chaining, ``AmplitudeEmbedding(pad_with, normalize)``, probs order,
``torch.tanh`` remap) are pinned by fixture F1: checkpoints trained by the real
PennyLane stack (``results/emnist.zip``) produce recognisable letters only under these
conventions (``tests/test_oracle.py``, golden vectors in ``tests/golden``); the sign of the
``expval(PauliZ)`` readout by the checkpoints of the <Z> families; ``qw_map.tanh = pi*tanh`` by
fixture F3: the one shipped ``QDenseUndirected_old`` checkpoint with its recorded training losses
and its 100 training images (``results_rebuttal_complex_dataset/logo2kplus.zip``) reproduces the
recorded loss (19.3-19.8 per epoch) only with pi*tanh (19.5; tanh 22.6, identity 23.6).
``AngleEmbedding(rotation="Y")`` and QConv's ``pad_with=0.5``/``[::2]`` stay "parity unpinned"
(documentation only: no shipped artefact exercises them).

Reference call sites each function follows are cited as ``nn/qdense.py:LINE`` etc.
(paths relative to ``/root/reference``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

CDTYPE = torch.complex128
RDTYPE = torch.float64

# enums shared (by value) with include/qiddm.h
INIT_ZERO, INIT_AMPLITUDE, INIT_BASIS = 0, 1, 2
ENC_NONE, ENC_RZ, ENC_RY = 0, 1, 2
IMP_CNOT, IMP_CZ = 0, 1
REMAP_NONE, REMAP_TANH, REMAP_PI_TANH = 0, 1, 2
READ_PROBS, READ_EXPVAL_Z, READ_STATE = 0, 1, 2


# --------------------------------------------------------------------------------------
# gates (PennyLane conventions)
# --------------------------------------------------------------------------------------
def remap_weights(w: torch.Tensor, remap: int) -> torch.Tensor:
    """nn/qdense.py:45 (qw_map.tanh = pi*tanh), :97 (torch.tanh), :171 (raw)."""
    if remap == REMAP_NONE:
        return w
    if remap == REMAP_TANH:
        return torch.tanh(w)
    if remap == REMAP_PI_TANH:
        return math.pi * torch.tanh(w)
    raise ValueError(remap)


def rot_matrix(phi, theta, omega) -> torch.Tensor:
    """qml.Rot(phi, theta, omega) = RZ(omega) RY(theta) RZ(phi); returns (..., 2, 2)."""
    phi = torch.as_tensor(phi, dtype=RDTYPE)
    theta = torch.as_tensor(theta, dtype=RDTYPE)
    omega = torch.as_tensor(omega, dtype=RDTYPE)
    c = torch.cos(theta / 2).to(CDTYPE)
    s = torch.sin(theta / 2).to(CDTYPE)
    ep = torch.exp(0.5j * (phi + omega).to(CDTYPE))
    em = torch.exp(0.5j * (phi - omega).to(CDTYPE))
    m00 = ep.conj() * c
    m01 = -em * s
    m10 = em.conj() * s
    m11 = ep * c
    return torch.stack([torch.stack([m00, m01], -1), torch.stack([m10, m11], -1)], -2)


def rz_matrix(a) -> torch.Tensor:
    a = torch.as_tensor(a, dtype=RDTYPE).to(CDTYPE)
    z = torch.zeros_like(a)
    e = torch.exp(-0.5j * a)
    return torch.stack([torch.stack([e, z], -1), torch.stack([z, e.conj()], -1)], -2)


def ry_matrix(a) -> torch.Tensor:
    a = torch.as_tensor(a, dtype=RDTYPE)
    c = torch.cos(a / 2).to(CDTYPE)
    s = torch.sin(a / 2).to(CDTYPE)
    return torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)


def apply_1q(state: torch.Tensor, mat: torch.Tensor, wire: int, n: int) -> torch.Tensor:
    """Apply a 2x2 (or per-sample (B,2,2)) matrix on ``wire``; state is (B, 2**n), wire 0 = MSB."""
    B = state.shape[0]
    st = state.reshape(B, 2 ** wire, 2, 2 ** (n - 1 - wire))
    s0, s1 = st[:, :, 0, :], st[:, :, 1, :]
    if mat.dim() == 3:
        m = mat[:, :, :, None, None]
        m00, m01, m10, m11 = m[:, 0, 0], m[:, 0, 1], m[:, 1, 0], m[:, 1, 1]
    else:
        m00, m01, m10, m11 = mat[0, 0], mat[0, 1], mat[1, 0], mat[1, 1]
    n0 = m00 * s0 + m01 * s1
    n1 = m10 * s0 + m11 * s1
    return torch.stack([n0, n1], dim=2).reshape(B, 2 ** n)


_PERM_CACHE: dict = {}


def _bit(k: torch.Tensor, wire: int, n: int) -> torch.Tensor:
    return (k >> (n - 1 - wire)) & 1


def ring_permutation(n: int, r: int) -> torch.Tensor:
    """Gather index for the CNOT ring of one SEL layer: new[k] = old[src[k]].

    CNOT(control=i, target=(i+r) mod n) for i = 0..n-1 in order (PennyLane
    StronglyEntanglingLayers), basis index k = sum_i b_i 2**(n-1-i).
    """
    key = ("cnot", n, r)
    if key not in _PERM_CACHE:
        src = torch.arange(2 ** n)
        for i in range(n):
            c, t = i, (i + r) % n
            k = torch.arange(2 ** n)
            # new[k] = old[k ^ (bit_c(k) << pos_t)]  (CNOT is an involution)
            g = k ^ (_bit(k, c, n) << (n - 1 - t))
            src = src[g]
        _PERM_CACHE[key] = src
    return _PERM_CACHE[key]


def ring_cz_sign(n: int, r: int) -> torch.Tensor:
    key = ("cz", n, r)
    if key not in _PERM_CACHE:
        k = torch.arange(2 ** n)
        par = torch.zeros_like(k)
        for i in range(n):
            par = par ^ (_bit(k, i, n) & _bit(k, (i + r) % n, n))
        _PERM_CACHE[key] = (1 - 2 * par).to(RDTYPE)
    return _PERM_CACHE[key]


def apply_ring(state: torch.Tensor, n: int, r: int, imprimitive: int) -> torch.Tensor:
    if n < 2:
        return state
    if imprimitive == IMP_CNOT:
        return state[:, ring_permutation(n, r)]
    return state * ring_cz_sign(n, r).to(state.dtype)


def sel_ranges(n_layers: int, n: int):
    """PennyLane default: r_l = (l mod (n-1)) + 1; each SEL call restarts l at 0."""
    if n < 2:
        return [0] * n_layers
    return [(l % (n - 1)) + 1 for l in range(n_layers)]


def strongly_entangling_layers(state, W, n: int, imprimitive: int, first_layer_right=None):
    """W: (D, n, 3) already re-mapped.  ``first_layer_right`` (optional list of n per-sample
    (B,2,2) matrices) is right-multiplied into the first layer's Rot (the encoding gate that
    immediately precedes it on the same wire)."""
    D = W.shape[0]
    ranges = sel_ranges(D, n)
    for l in range(D):
        for i in range(n):
            m = rot_matrix(W[l, i, 0], W[l, i, 1], W[l, i, 2])
            if l == 0 and first_layer_right is not None:
                m = m[None] @ first_layer_right[i]
            state = apply_1q(state, m, i, n)
        state = apply_ring(state, n, ranges[l], imprimitive)
    return state


# --------------------------------------------------------------------------------------
# one circuit "stage" = what a single QNode call computes
# --------------------------------------------------------------------------------------
@dataclass
class StageDesc:
    """Mirror of ``qiddm_circuit_desc`` (include/qiddm.h)."""
    n_qubits: int
    n_blocks: int = 1            # L (re-upload blocks); 1 for plain SEL circuits
    layers_per_block: int = 1    # SEL depth inside a block
    init: int = INIT_ZERO
    n_features: int = 0          # AMPLITUDE: number of real features F <= 2**n
    pad_value: float = 0.0       # AMPLITUDE: pad_with
    add_offset: float = 0.0      # AMPLITUDE: constant added to the features first (QConv +0.1)
    enc: int = ENC_NONE          # per-block single-qubit data gate before the block's SEL
    enc_scale: float = 1.0       # angle = enc_scale * input
    imprimitive: int = IMP_CNOT
    remap: int = REMAP_NONE
    readout: int = READ_PROBS
    read_count: int = 0          # PROBS: number of outputs K
    read_stride: int = 1         # PROBS: out[m] = p[m*stride]
    post_scale: float = 1.0
    clamp: bool = False
    clamp_lo: float = 0.0
    clamp_hi: float = 1.0

    @property
    def dim(self) -> int:
        return 2 ** self.n_qubits

    @property
    def n_in(self) -> int:
        if self.init == INIT_AMPLITUDE:
            return self.n_features
        if self.enc != ENC_NONE:
            return self.n_qubits
        return 0

    @property
    def n_out(self) -> int:
        if self.readout == READ_PROBS:
            return self.read_count
        if self.readout == READ_EXPVAL_Z:
            return self.n_qubits
        return 2 * self.dim


def amplitude_embedding(x: torch.Tensor, n: int, pad_with: float, add_offset: float = 0.0):
    """qml.AmplitudeEmbedding(features, pad_with=c, normalize=True) (nn/qdense.py:41-43,
    nn/qconv.py:52-54): append c up to 2**n, divide by the L2 norm, cast complex."""
    B, F = x.shape
    A = 2 ** n
    x = x.to(RDTYPE) + add_offset
    if F < A:
        x = torch.cat([x, torch.full((B, A - F), pad_with, dtype=RDTYPE)], dim=1)
    x = x / torch.linalg.vector_norm(x, dim=1, keepdim=True)
    return x.to(CDTYPE)


def expval_z(state: torch.Tensor, n: int) -> torch.Tensor:
    p = (state.real ** 2 + state.imag ** 2)
    k = torch.arange(2 ** n)
    signs = torch.stack([(1 - 2 * _bit(k, j, n)).to(RDTYPE) for j in range(n)], dim=1)  # (A, n)
    return p @ signs


def run_stage(desc: StageDesc, x: Optional[torch.Tensor], weights: torch.Tensor,
              batch: Optional[int] = None, basis_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One QNode evaluation for a batch.  ``weights``: (L, D, n, 3) raw (un-remapped).
    Returns (B, n_out) float64 (READ_STATE: interleaved re/im)."""
    n, A = desc.n_qubits, desc.dim
    W = remap_weights(weights.to(RDTYPE).reshape(desc.n_blocks, desc.layers_per_block, n, 3), desc.remap)
    if desc.init == INIT_AMPLITUDE:
        state = amplitude_embedding(x, n, desc.pad_value, desc.add_offset)
        B = state.shape[0]
        ang = None
    elif desc.init == INIT_BASIS:
        B = basis_index.shape[0]
        state = torch.zeros(B, A, dtype=CDTYPE)
        state[torch.arange(B), basis_index] = 1.0
        ang = None
    else:
        B = x.shape[0] if x is not None else int(batch)
        state = torch.zeros(B, A, dtype=CDTYPE)
        state[:, 0] = 1.0
        ang = None
    if desc.enc != ENC_NONE:
        ang = x.to(RDTYPE)[:, :n] * desc.enc_scale
    for blk in range(desc.n_blocks):
        right = None
        if desc.enc == ENC_RZ:
            right = [rz_matrix(ang[:, j]) for j in range(n)]
        elif desc.enc == ENC_RY:
            right = [ry_matrix(ang[:, j]) for j in range(n)]
        state = strongly_entangling_layers(state, W[blk], n, desc.imprimitive, right)
    if desc.readout == READ_STATE:
        return torch.view_as_real(state).reshape(B, 2 * A)
    if desc.readout == READ_EXPVAL_Z:
        out = expval_z(state, n)
    else:
        p = state.real ** 2 + state.imag ** 2
        out = p[:, : desc.read_count * desc.read_stride: desc.read_stride]
    out = out * desc.post_scale
    if desc.clamp:
        out = torch.clamp(out, desc.clamp_lo, desc.clamp_hi)
    return out


def circuit_unitary(desc: StageDesc, weights: torch.Tensor) -> torch.Tensor:
    """Full (A, A) complex unitary of the weight-only part of a single-block circuit
    (what nn/qconv.py:92-126 builds with qml.matrix for eval mode)."""
    d = StageDesc(**{**desc.__dict__, "init": INIT_BASIS, "enc": ENC_NONE, "readout": READ_STATE})
    A = desc.dim
    cols = run_stage(d, None, weights, basis_index=torch.arange(A))      # row c = U e_c
    U_t = torch.view_as_complex(cols.reshape(A, A, 2).contiguous())
    return U_t.transpose(0, 1)


# --------------------------------------------------------------------------------------
# descriptors of the reference module families
# --------------------------------------------------------------------------------------
def desc_qdense(qdepth: int, pixels: int, remap: int) -> StageDesc:
    """a1: QDenseUndirected_old (remap pi*tanh) / _old_noise (tanh)  nn/qdense.py:40-54, 95-111."""
    n = math.ceil(math.log2(pixels))
    return StageDesc(n_qubits=n, n_blocks=1, layers_per_block=qdepth, init=INIT_AMPLITUDE,
                     n_features=pixels, pad_value=0.1, imprimitive=IMP_CNOT, remap=remap,
                     readout=READ_PROBS, read_count=pixels, post_scale=float(pixels), clamp=True)


def desc_qnn_a(qdepth: int, pixels: int) -> StageDesc:
    """a2: QNN_A  nn/qdense.py:162-190 (AngleEmbedding Y, SEL CNOT, probs)."""
    n = math.ceil(math.log2(pixels))
    return StageDesc(n_qubits=n, n_blocks=1, layers_per_block=qdepth, init=INIT_ZERO, enc=ENC_RY,
                     imprimitive=IMP_CNOT, readout=READ_PROBS, read_count=pixels,
                     post_scale=float(pixels), clamp=True)


def desc_reupload(n: int, L: int, layers: int = 2, enc: int = ENC_RZ, enc_scale: float = 1.0,
                  readout: int = READ_EXPVAL_Z, read_count: int = 0, post_scale: float = 1.0,
                  clamp: bool = False) -> StageDesc:
    """a3/a4/a5: RZ (or RY) re-upload blocks + SEL(CZ)  nn/qdense.py:422-441, 1599-1617, 249-265."""
    return StageDesc(n_qubits=n, n_blocks=L, layers_per_block=layers, init=INIT_ZERO, enc=enc,
                     enc_scale=enc_scale, imprimitive=IMP_CZ, readout=readout, read_count=read_count,
                     post_scale=post_scale, clamp=clamp)


def desc_qconv(in_channels: int, out_channels: int, kernel_size, qdepth: int) -> StageDesc:
    """a6: _QConv2d_FAST (with the H1 fix)  nn/qconv.py:24-28, 51-69."""
    kh, kw = kernel_size
    F = in_channels * kh * kw
    n = max(math.ceil(math.log2(F)), math.ceil(math.log2(out_channels)), 1)
    return StageDesc(n_qubits=n, n_blocks=1, layers_per_block=qdepth, init=INIT_AMPLITUDE,
                     n_features=F, pad_value=0.5, add_offset=0.1, imprimitive=IMP_CNOT,
                     remap=REMAP_PI_TANH, readout=READ_PROBS, read_count=out_channels, read_stride=2,
                     post_scale=0.5 * 2 ** n, clamp=True)


# --------------------------------------------------------------------------------------
# module-level forwards (what the reference nn.Modules compute around the QNode)
# --------------------------------------------------------------------------------------
def qdense_forward(x: torch.Tensor, weights: torch.Tensor, remap: int) -> torch.Tensor:
    """QDenseUndirected_old[_noise].forward  nn/qdense.py:56-62 / 113-119.  x: (B,1,w,h)."""
    B, _, w, h = x.shape
    d = desc_qdense(weights.shape[0], w * h, remap)
    return run_stage(d, x.reshape(B, w * h), weights[None]).reshape(B, 1, w, h)


def differN_forward(angles: torch.Tensor, weights: torch.Tensor, pixels: int) -> torch.Tensor:
    """differN_noise / differN_old_pca after the PCA  nn/qdense.py:463-469: N chained stages,
    each probs (B, 2**n); the next stage reads columns 0..n-1 as its angles (:427)."""
    N, L, D, n, _ = weights.shape
    a = angles
    for k in range(N):
        last = k == N - 1
        d = desc_reupload(n, L, D, readout=READ_PROBS,
                          read_count=pixels if last else n,
                          post_scale=float(pixels) if last else 1.0, clamp=last)
        a = run_stage(d, a, weights[k])
    return a


def qiddm_expval_chain(angles: torch.Tensor, weights1: torch.Tensor) -> torch.Tensor:
    """QIDDM_{LL,PL}_noise inner loop  nn/qdense.py:1631-1635: N stages of <Z_j> chaining."""
    N, L, D, n, _ = weights1.shape
    a = angles
    for k in range(N):
        a = run_stage(desc_reupload(n, L, D), a, weights1[k])
    return a


def qiddm_ll_forward(x, weights1, w_down, b_down, w_up, b_up):
    """QIDDM_LL_noise.forward  nn/qdense.py:1620-1642 (with the TRUE gradient)."""
    B, c, w, h = x.shape
    a = x.reshape(B, -1).to(RDTYPE) @ w_down.to(RDTYPE).T + b_down.to(RDTYPE)
    a = qiddm_expval_chain(a, weights1)
    out = a @ w_up.to(RDTYPE).T + b_up.to(RDTYPE)
    return out.reshape(B, c, w, h)


def qnn_forward(x, weights, w_down, b_down, w_up, b_up):
    """QNN / QNN_noise.forward  nn/qdense.py:267-289."""
    B, c, w, h = x.shape
    qdepth, n, _ = weights.shape
    a = x.reshape(B, -1).to(RDTYPE) @ w_down.to(RDTYPE).T + b_down.to(RDTYPE)
    a = run_stage(desc_reupload(n, 1, qdepth), a, weights[None])
    out = a @ w_up.to(RDTYPE).T + b_up.to(RDTYPE)
    return out.reshape(B, c, w, h)


def unfold_patches(x: torch.Tensor, kernel_size, padding) -> torch.Tensor:
    """nn/qconv.py:76-77: Unfold then '(batch feat) channel'.  Returns (B*H_out*W_out, C*kh*kw)."""
    cols = torch.nn.functional.unfold(x.to(RDTYPE), kernel_size=kernel_size, padding=padding)
    B, F, P = cols.shape
    return cols.permute(0, 2, 1).reshape(B * P, F)


def qconv_forward(x: torch.Tensor, weights: torch.Tensor, out_channels: int, kernel_size=(3, 3),
                  padding=(1, 1)) -> torch.Tensor:
    """_QConv2d_FAST.forward with the H1 fix (qnode between :78 and :79)  nn/qconv.py:71-87."""
    B, C, H, W = x.shape
    kh, kw = kernel_size
    ph, pw = padding
    h_out, w_out = H + 2 * ph - kh + 1, W + 2 * pw - kw + 1
    d = desc_qconv(C, out_channels, kernel_size, weights.shape[0])
    patches = unfold_patches(x, kernel_size, padding)
    out = run_stage(d, patches, weights[None])                      # (B*P, out)
    return out.reshape(B, h_out, w_out, out_channels).permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------------------
# diffusion wrapper + noise ladder (src/models.py, src/noise.py)
# --------------------------------------------------------------------------------------
def noise_ladder(data: torch.Tensor, eps: torch.Tensor, tau: int, decay_mod: float = 3.0) -> torch.Tensor:
    """add_normal_noise_multiple with the normal draw ``eps`` passed in  src/noise.py:105-126.
    Returns ((batch tau), pixels)."""
    w = torch.linspace(0, 1, tau, dtype=data.dtype) ** decay_mod
    w = w / w.max()
    w = w[:, None, None]
    noisy = data[None] * (1 - w) + eps[None] * w
    noisy = noisy.clamp(0, 1)
    return noisy.permute(1, 0, 2).reshape(-1, data.shape[1])


def training_targets(data, eps, T: int, shape):
    """src/models.py:44-63: returns (noisy, clean) as ((batch T),1,w,h)."""
    whole = noise_ladder(data, eps, T + 1).reshape(data.shape[0], T + 1, -1)
    wd, ht = shape
    noisy = whole[:, 1:, :].reshape(-1, 1, wd, ht)
    clean = whole[:, :-1, :].reshape(-1, 1, wd, ht)
    return noisy, clean


def diffusion_loss(net_fn, data, eps, T: int, shape, goal: str = "data"):
    """src/models.py:64-67 / 94-99: mean MSE of the training step."""
    noisy, clean = training_targets(data, eps, T, shape)
    pred = net_fn(noisy)
    if goal == "data":
        return ((pred - clean) ** 2).mean()
    pred = (pred - 0.5) * 0.1
    return ((pred - (noisy - clean)) ** 2).mean()


def sample(net_fn, first_x: torch.Tensor, n_iters: int, goal: str = "data", noise_factor: float = 1.0):
    """Diffusion.sample(only_last=True)  src/models.py:106-139."""
    x = first_x
    with torch.no_grad():
        for _ in range(n_iters):
            pred = net_fn(x)
            if goal == "data":
                x = pred
            else:
                x = torch.clamp(x - (pred - 0.5) * 0.1 * noise_factor, 0, 1)
    return x


# --------------------------------------------------------------------------------------
# noise channels of the `*_noise` probability models (nn/qdense.py:98-104, :174-180, :431-439 on default.mixed)
# Kraus operators as documented for PennyLane 0.29 (qml.PhaseDamping, qml.AmplitudeDamping, qml.DepolarizingChannel,
# qml.PhaseShift); PennyLane itself is not installable here, so this part of the oracle is "parity unpinned".
NOISE_PHASE, NOISE_AMPLITUDE_DAMPING, NOISE_DEPOLARIZING = 1, 2, 3


def kraus_operators(kind: int, param: float, phase_shift: bool = False):
    z = torch.zeros((), dtype=CDTYPE)
    one = torch.ones((), dtype=CDTYPE)
    if kind == NOISE_PHASE:
        if phase_shift:                                      # qml.PhaseShift(phi) = diag(1, e^{i phi}) (a unitary)
            return [torch.stack([torch.stack([one, z]), torch.stack([z, torch.exp(1j * torch.tensor(param, dtype=RDTYPE)).to(CDTYPE)])])]
        g = torch.tensor(param, dtype=RDTYPE)                # qml.PhaseDamping(g)
        return [torch.diag(torch.stack([one, torch.sqrt(1 - g).to(CDTYPE)])),
                torch.diag(torch.stack([z, torch.sqrt(g).to(CDTYPE)]))]
    if kind == NOISE_AMPLITUDE_DAMPING:                      # qml.AmplitudeDamping(g)
        g = torch.tensor(param, dtype=RDTYPE)
        k0 = torch.diag(torch.stack([one, torch.sqrt(1 - g).to(CDTYPE)]))
        k1 = torch.zeros(2, 2, dtype=CDTYPE)
        k1[0, 1] = torch.sqrt(g)
        return [k0, k1]
    if kind == NOISE_DEPOLARIZING:                           # qml.DepolarizingChannel(p): sqrt(1-p) I, sqrt(p/3) X, Y, Z
        p = torch.tensor(param, dtype=RDTYPE)
        eye = torch.eye(2, dtype=CDTYPE)
        x = torch.tensor([[0, 1], [1, 0]], dtype=CDTYPE)
        y = torch.tensor([[0, -1j], [1j, 0]], dtype=CDTYPE)
        zz = torch.tensor([[1, 0], [0, -1]], dtype=CDTYPE)
        return [torch.sqrt(1 - p) * eye, torch.sqrt(p / 3) * x, torch.sqrt(p / 3) * y, torch.sqrt(p / 3) * zz]
    raise ValueError(kind)


def density_matrix_readout(state: torch.Tensor, n: int, kind: int, param: float, phase_shift: bool = False) -> torch.Tensor:
    """What default.mixed computes: rho = |psi><psi|, the channel on every wire (Kraus sum), then the diagonal.
    state (B, 2**n) complex -> (B, 2**n) probabilities.  Exponential in n: small cases only."""
    B, A = state.shape
    rho = state[:, :, None] * state.conj()[:, None, :]                  # (B, A, A)
    ks = kraus_operators(kind, param, phase_shift)
    for wire in range(n):
        shape = (B,) + (2,) * (2 * n)
        r = rho.reshape(shape)
        row_ax, col_ax = 1 + wire, 1 + n + wire
        acc = torch.zeros_like(r)
        for k in ks:
            t = torch.movedim(torch.tensordot(k, torch.movedim(r, row_ax, 0), dims=([1], [0])), 0, row_ax)            # K rho
            t = torch.movedim(torch.tensordot(k.conj(), torch.movedim(t, col_ax, 0), dims=([1], [0])), 0, col_ax)    # .. K^dagger
            acc = acc + t
        rho = acc.reshape(B, A, A)
    return torch.diagonal(rho, dim1=1, dim2=2).real


def readout_channel_probs(p: torch.Tensor, n: int, kind: int, param: float) -> torch.Tensor:
    """The same thing as a classical map on the probabilities (what the CUDA path implements): per wire
    AmplitudeDamping (p0, p1) -> (p0 + g p1, (1 - g) p1); Depolarizing: bit flip with probability 2p/3; phase: identity."""
    if kind == NOISE_PHASE:
        return p
    if kind == NOISE_AMPLITUDE_DAMPING:
        m = torch.tensor([[1.0, param], [0.0, 1.0 - param]], dtype=p.dtype)
    else:
        q = 2.0 * param / 3.0
        m = torch.tensor([[1.0 - q, q], [q, 1.0 - q]], dtype=p.dtype)
    B = p.shape[0]
    t = p.reshape((B,) + (2,) * n)
    for wire in range(n):
        t = torch.movedim(torch.tensordot(m, torch.movedim(t, 1 + wire, 0), dims=([1], [0])), 0, 1 + wire)
    return t.reshape(B, -1)


# --------------------------------------------------------------------------------------
# mid-circuit noise channels of the re-upload classes on `default.mixed` (nn/qdense.py:515-527, :1405-1417, :1599-1617;
# src/mnist_noise.py:211-229).  Literal tape order on the full density matrix: per block, per wire j: RZ(a_j) then the channel's
# Kraus sum on wire j; then StronglyEntanglingLayers (Rot on every wire, CZ ring).  Exponential in n: small cases only.
# Same status as the readout channels above: Kraus operators as documented for PennyLane 0.29 -- "parity unpinned".
def _embed_1q(mat: torch.Tensor, wire: int, n: int) -> torch.Tensor:
    """Full 2**n x 2**n matrix of a 2 x 2 operator on ``wire`` (wire 0 = MSB)."""
    full = torch.ones(1, 1, dtype=CDTYPE)
    for i in range(n):
        full = torch.kron(full, mat.to(CDTYPE) if i == wire else torch.eye(2, dtype=CDTYPE))
    return full


def noisy_reupload_stage(desc: StageDesc, angles: torch.Tensor, weights: torch.Tensor, kind: int, param: float) -> torch.Tensor:
    """One QNode call of a re-upload class with `add_noise = kind` (1 PhaseDamping, 2 AmplitudeDamping, 3 DepolarizingChannel,
    parameter ``param``) after every RZ(a_j).  angles (B, n), weights (L, D, n, 3) -> (B, n_out) float64."""
    assert desc.init == INIT_ZERO and desc.enc == ENC_RZ
    n, A = desc.n_qubits, desc.dim
    B = angles.shape[0]
    W = weights.to(RDTYPE).reshape(desc.n_blocks, desc.layers_per_block, n, 3)
    ang = angles.to(RDTYPE)[:, :n] * desc.enc_scale
    rho = torch.zeros(B, A, A, dtype=CDTYPE)
    rho[:, 0, 0] = 1.0
    kraus = [[_embed_1q(k, j, n) for k in kraus_operators(kind, param)] for j in range(n)]
    k_idx = torch.arange(A)
    ranges = sel_ranges(desc.layers_per_block, n)
    for blk in range(desc.n_blocks):
        for j in range(n):
            sgn = (2 * _bit(k_idx, j, n) - 1).to(RDTYPE)                       # RZ(a) = diag(e^{-ia/2}, e^{+ia/2})
            d = torch.exp(0.5j * (ang[:, j:j + 1] * sgn[None, :]).to(CDTYPE))  # (B, A)
            rho = d[:, :, None] * rho * d.conj()[:, None, :]
            rho = sum(k[None] @ rho @ k.conj().T[None] for k in kraus[j])
        for l in range(desc.layers_per_block):
            for i in range(n):
                u = _embed_1q(rot_matrix(W[blk, l, i, 0], W[blk, l, i, 1], W[blk, l, i, 2]), i, n)
                rho = u[None] @ rho @ u.conj().T[None]
            if n > 1:
                if desc.imprimitive == IMP_CZ:
                    s = ring_cz_sign(n, ranges[l]).to(CDTYPE)
                    rho = s[None, :, None] * rho * s[None, None, :]
                else:
                    src = ring_permutation(n, ranges[l])
                    rho = rho[:, src][:, :, src]
    p = torch.diagonal(rho, dim1=1, dim2=2).real
    if desc.readout == READ_EXPVAL_Z:
        signs = torch.stack([(1 - 2 * _bit(k_idx, j, n)).to(RDTYPE) for j in range(n)], dim=1)
        out = p @ signs
    else:
        out = p[:, : desc.read_count * desc.read_stride: desc.read_stride]
    out = out * desc.post_scale
    if desc.clamp:
        out = torch.clamp(out, desc.clamp_lo, desc.clamp_hi)
    return out
