"""Readout noise channels (add_noise of QDenseUndirected_old_noise / QNN_A / differN_noise, nn/qdense.py:98-104, :174-180,
:431-439): the classical probability map the CUDA path implements equals the density-matrix (Kraus)
simulation default.mixed performs, and the modules equal the oracle pipeline."""
import pytest
import torch

from conftest import rel_to_max
from oracle import qiddm_oracle as O


@pytest.mark.parametrize("n", [1, 2, 3, 5])
@pytest.mark.parametrize("kind,param,ps", [(1, 0.05, True), (1, 0.05, False), (2, 0.1, False), (2, 0.05, False), (3, 0.02, False)])
def test_channel_before_measurement_is_a_classical_map(n, kind, param, ps):
    g = torch.Generator().manual_seed(n * 10 + kind)
    psi = torch.randn(4, 1 << n, generator=g, dtype=torch.float64) + 1j * torch.randn(4, 1 << n, generator=g, dtype=torch.float64)
    psi = psi / psi.abs().pow(2).sum(1, keepdim=True).sqrt()
    dm = O.density_matrix_readout(psi, n, kind, param, phase_shift=ps)
    cl = O.readout_channel_probs(psi.abs() ** 2, n, kind, param)
    assert (dm - cl).abs().max() <= 1e-14
    assert (dm.sum(1) - 1).abs().max() <= 1e-14            # trace preserving


def test_mid_circuit_noise_classes_construct_and_read_the_flag_at_call_time():
    """`add_noise` is a constructor argument AND is flipped on trained nets at test time (src/mnist_noise.py:218).  The RZ
    re-upload classes route it to the density-matrix path (inference only: refuses under autograd); the RY-encoded
    QIDDM_PL_noise1 has no such path and refuses at call time -- never a silent noiseless result."""
    from qiddm_b200 import nn
    for make in (lambda: nn.QIDDM_PL_noise(64, 4, 2, 2, add_noise=2), lambda: nn.QIDDM_LL_noise(64, 4, 2, 2, add_noise=3),
                 lambda: nn.differN_noise_befor(8, 2, 2, add_noise=2)):
        assert make().add_noise in (2, 3)
    with pytest.raises(NotImplementedError):
        nn.QIDDM_LL_noise(64, 4, 2, 2, add_noise=7)
    m = nn.QIDDM_LL_noise(64, 4, 2, 2)
    m.add_noise = 2                                         # flipped after construction, as mnist_noise.test() does
    with pytest.raises(NotImplementedError, match="inference path"):
        m(torch.rand(1, 1, 8, 8))                            # autograd enabled: there is no density-matrix backward
    ry = nn.QIDDM_PL_noise1(64, 4, 2, 2)
    ry.add_noise = 1
    with pytest.raises(NotImplementedError):
        ry(torch.rand(2, 1, 8, 8))
    assert nn.QDenseUndirected_old_noise(2, 8, add_noise=3).add_noise == 3
    assert nn.differN_noise(8, 2, 2, add_noise=2).add_noise == 2


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_mid_circuit_oracle_reduces_to_the_state_vector_stage_and_keeps_the_trace(kind):
    """Zero-strength channels give the noiseless stage; any strength keeps Tr rho = 1; DepolarizingChannel(3/4) on every wire of
    the last block's input is the fully mixed state only at p = 3/4 -- with p = 0.9 (the reference's value) <Z> stays non-trivial."""
    g = torch.Generator().manual_seed(kind)
    d = O.desc_reupload(4, 3, 2)
    W = torch.randn(3, 2, 4, 3, generator=g, dtype=torch.float64) * 0.4
    a = torch.randn(5, 4, generator=g, dtype=torch.float64)
    assert (O.noisy_reupload_stage(d, a, W, kind, 0.0) - O.run_stage(d, a, W)).abs().max() <= 1e-13
    dp = O.desc_reupload(4, 3, 2, readout=O.READ_PROBS, read_count=16)
    p = O.noisy_reupload_stage(dp, a, W, kind, {1: 0.03, 2: 0.05, 3: 0.9}[kind])
    assert (p.sum(1) - 1).abs().max() <= 1e-13 and p.min() >= -1e-15
    assert (p - O.run_stage(dp, a, W)).abs().max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("n,L_,D,readout", [(4, 3, 2, "z"), (5, 2, 2, "z"), (6, 3, 2, "z"), (6, 2, 3, "p"), (3, 4, 2, "p")])
@pytest.mark.parametrize("kind,param", [(1, 0.03), (2, 0.05), (3, 0.9), (3, 0.02)])
def test_mid_circuit_noise_stage_matches_the_density_matrix_oracle(n, L_, D, readout, kind, param):
    """qiddm_noisy_forward (class-wise channel butterflies + gate kernels on the rows of rho^T) against the literal Kraus
    simulation, n = 3 ... 6 (both gate-kernel schedules: n = 6 with 3-bit tiles is the resident one)."""
    from qiddm_b200 import _lib as L
    from qiddm_b200.channels import run_noisy_stage
    g = torch.Generator().manual_seed(100 * n + kind)
    if readout == "z":
        d = O.desc_reupload(n, L_, D)
    else:
        d = O.desc_reupload(n, L_, D, readout=O.READ_PROBS, read_count=(1 << n) - 1)
        d.post_scale, d.clamp = float(1 << n), True
    W = torch.randn(L_, D, n, 3, generator=g, dtype=torch.float64) * 0.4
    a = torch.randn(7, n, generator=g, dtype=torch.float64)
    ref = O.noisy_reupload_stage(d, a, W, kind, param)
    spec = L.StageSpec(n_qubits=n, n_blocks=L_, layers_per_block=D, init=L.INIT_ZERO, enc=L.ENC_RZ, imprimitive=L.IMP_CZ,
                       readout=L.READ_EXPVAL_Z if readout == "z" else L.READ_PROBS, read_count=d.read_count,
                       post_scale=d.post_scale, clamp=d.clamp)
    with torch.no_grad():
        out = run_noisy_stage(spec, a.cuda(), W.cuda(), kind, {kind: param})
    assert out.shape == ref.shape
    assert rel_to_max(out, ref, floor=1e-3) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("noise", [1, 2, 3])
def test_qiddm_ll_noise_module_with_the_flag_flipped_at_test_time(noise):
    """src/mnist_noise.py:211-229: a trained QIDDM_LL_noise gets `add_noise = k` and is sampled under no_grad on default.mixed."""
    from qiddm_b200 import nn
    torch.manual_seed(noise)
    m = nn.QIDDM_LL_noise(64, 4, 3, 2).to("cuda", torch.float64)
    x = torch.rand(3, 1, 8, 8, dtype=torch.float64)
    ps = {k: v.detach().cpu() for k, v in m.named_parameters()}
    clean = m(x.cuda()).detach()
    m.add_noise = noise
    with torch.no_grad():
        out = m(x.cuda())
    a = x.reshape(3, 64) @ ps["linear_down.weight"].T + ps["linear_down.bias"]
    for k in range(2):
        a = O.noisy_reupload_stage(O.desc_reupload(4, 3, 2), a, ps["weights1"][k], noise, {1: 0.03, 2: 0.05, 3: 0.9}[noise])
    ref = (a @ ps["linear_up.weight"].T + ps["linear_up.bias"]).reshape(3, 1, 8, 8)
    assert rel_to_max(out, ref) <= 2e-5
    assert rel_to_max(out, clean) > 1e-3                    # the flag is honoured


@pytest.mark.gpu
def test_differn_noise_befor_chain_with_mid_circuit_channels():
    """nn/qdense.py:481-562 with add_noise = 2: two chained stages, each a density-matrix run; the next stage's angles are the
    first n probabilities."""
    from qiddm_b200 import nn
    torch.manual_seed(5)
    m = nn.differN_noise_befor(4, 2, 2).cuda()              # 4 x 4 pixels -> 4 wires
    a = torch.randn(5, m.wires, dtype=torch.float64)
    W = m.weights.detach().cpu().double()
    m.add_noise = 2
    with torch.no_grad():
        out = m._chain(a.cuda())
    n = m.wires
    d_mid = O.desc_reupload(n, 2, 2, readout=O.READ_PROBS, read_count=n)
    d_last = O.desc_reupload(n, 2, 2, readout=O.READ_PROBS, read_count=16)
    d_last.post_scale, d_last.clamp = 16.0, True
    ref = O.noisy_reupload_stage(d_last, O.noisy_reupload_stage(d_mid, a, W[0], 2, 0.05), W[1], 2, 0.05)
    assert rel_to_max(out, ref) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n,kind,param", [(1, 2, 0.1), (4, 2, 0.05), (6, 3, 0.02), (10, 2, 0.1), (12, 3, 0.02)])
def test_readout_channel_kernel_and_its_transpose(dtype, n, kind, param):
    from qiddm_b200 import _lib as L
    from qiddm_b200.channels import apply_readout_channel
    g = torch.Generator().manual_seed(n + kind)
    p = torch.rand(5, 1 << n, generator=g, dtype=torch.float64)
    p = p / p.sum(1, keepdim=True)
    pr = p.clone().requires_grad_(True)
    ref = O.readout_channel_probs(pr, n, kind, param)
    go = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * go).sum().backward()
    pd = p.to(dtype).cuda().requires_grad_(True)
    n0 = L.launch_count()
    out = apply_readout_channel(pd, n, kind, param)
    (out * go.to(dtype).cuda()).sum().backward()
    assert L.launch_count() - n0 == 2
    tol = 1e-12 if dtype == torch.float64 else 2e-6
    assert rel_to_max(out, ref) <= tol and rel_to_max(pd.grad, pr.grad) <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("noise", [1, 2, 3])
def test_qdense_noise_module_matches_the_density_matrix_oracle(noise):
    """QDenseUndirected_old_noise(add_noise) == AmplitudeEmbedding -> SEL(tanh W) -> channels (Kraus, density matrix) ->
    probs -> _post_process, incl. the weight gradient through the channel."""
    from qiddm_b200 import nn
    torch.manual_seed(2)
    m = nn.QDenseUndirected_old_noise(3, 4, add_noise=noise).to("cuda", torch.float64)      # 4x4 image, n = 4
    x = torch.rand(6, 1, 4, 4, dtype=torch.float64)
    W = m.weights.detach().cpu().clone().requires_grad_(True)
    d = O.desc_qdense(3, 16, O.REMAP_TANH)
    full = O.StageDesc(**{**d.__dict__, "readout": O.READ_STATE, "clamp": False, "post_scale": 1.0})
    st = O.run_stage(full, x.reshape(6, 16), W.reshape(1, 3, 4, 3))
    psi = torch.view_as_complex(st.reshape(6, 16, 2).contiguous())
    param = {1: 0.05, 2: 0.1, 3: 0.02}[noise]
    probs = O.density_matrix_readout(psi, 4, noise, param, phase_shift=(noise == 1))
    ref = torch.clamp(probs[:, :16] * 16, 0, 1).reshape(6, 1, 4, 4)
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    out = m(x.cuda())
    (out * g.cuda()).sum().backward()
    assert rel_to_max(out, ref) <= 1e-5
    assert rel_to_max(m.weights.grad, W.grad) <= 1e-4
    # src/mnist_noise.py:218 flips add_noise on a trained net at test time
    m.add_noise = 0
    clean = m(x.cuda())
    assert (noise == 1) == (rel_to_max(clean, out) <= 1e-6)


@pytest.mark.gpu
def test_differn_noise_chain_with_readout_channels():
    from qiddm_b200 import nn
    torch.manual_seed(4)
    m = nn.differN_noise(4, 2, 2, add_noise=3).cuda()                 # 4x4 image, n = 4, L = 2, N = 2
    ang = torch.rand(5, 4, dtype=torch.float64)
    W = m.weights.detach().cpu().double()
    a = ang
    for k in range(2):
        d = O.desc_reupload(4, 2, 2, readout=O.READ_STATE)
        st = O.run_stage(d, a[:, :4], W[k])
        psi = torch.view_as_complex(st.reshape(5, 16, 2).contiguous())
        a = O.density_matrix_readout(psi, 4, 3, 0.02)
    ref = torch.clamp(a[:, :16] * 16, 0, 1)
    out = m._chain(ang.float().cuda())
    assert rel_to_max(out, ref) <= 1e-5
    whole = m(torch.rand(8, 1, 4, 4).cuda())                          # with the PCA in front
    assert whole.shape == (8, 1, 4, 4)


def _density_qnn(angles, W, n, kind, param):
    """Density-matrix simulation of QNN_noise._circuit (nn/qdense.py:249-265): RZ(a_j) + channel on wire j, SEL(CZ), <Z_i>."""
    B = angles.shape[0]
    A = 1 << n
    psi = torch.zeros(B, A, dtype=O.CDTYPE)
    psi[:, 0] = 1.0
    for j in range(n):
        psi = O.apply_1q(psi, O.rz_matrix(angles[:, j]), j, n)
    rho = psi[:, :, None] * psi.conj()[:, None, :]
    ks = O.kraus_operators(kind, param)
    for wire in range(n):
        r = rho.reshape((B,) + (2,) * (2 * n))
        acc = torch.zeros_like(r)
        for k in ks:
            t = torch.movedim(torch.tensordot(k, torch.movedim(r, 1 + wire, 0), dims=([1], [0])), 0, 1 + wire)
            t = torch.movedim(torch.tensordot(k.conj(), torch.movedim(t, 1 + n + wire, 0), dims=([1], [0])), 0, 1 + n + wire)
            acc = acc + t
        rho = acc.reshape(B, A, A)
    d = O.StageDesc(n_qubits=n, n_blocks=1, layers_per_block=W.shape[0], init=O.INIT_BASIS, imprimitive=O.IMP_CZ, readout=O.READ_STATE)
    U = O.circuit_unitary(d, W.reshape(1, W.shape[0], n, 3))
    rho = U[None] @ rho @ U.conj().T[None]
    p = torch.diagonal(rho, dim1=1, dim2=2).real
    k = torch.arange(A)
    return torch.stack([(p * (1 - 2 * ((k >> (n - 1 - i)) & 1)).to(p.dtype)).sum(1) for i in range(n)], dim=1)


def test_qnn_noise_reduction_matches_density_matrix_on_cpu():
    """The reduction QNN_noise uses (mixture over basis inputs for Depolarizing; nothing for the dampings) == the Kraus
    density-matrix simulation of the reference circuit."""
    n, depth = 3, 2
    g = torch.Generator().manual_seed(5)
    W = torch.randn(depth, n, 3, generator=g, dtype=torch.float64) * 0.4
    a = torch.randn(4, n, generator=g, dtype=torch.float64)
    d = O.desc_reupload(n, 1, depth)
    clean = O.run_stage(d, a, W.reshape(1, depth, n, 3))
    for kind, param in ((1, 0.03), (2, 0.05)):
        assert (_density_qnn(a, W, n, kind, param) - clean).abs().max() <= 1e-13
    dm = _density_qnn(a, W, n, 3, 0.02)
    db = O.StageDesc(n_qubits=n, n_blocks=1, layers_per_block=depth, init=O.INIT_BASIS, imprimitive=O.IMP_CZ, readout=O.READ_EXPVAL_Z)
    vals = O.run_stage(db, None, W.reshape(1, depth, n, 3), basis_index=torch.arange(1 << n))
    q = 2 * 0.02 / 3
    ones = torch.tensor([bin(s).count("1") for s in range(1 << n)], dtype=torch.float64)
    mix = ((q ** ones) * ((1 - q) ** (n - ones)))[:, None] * vals
    assert (dm - mix.sum(0, keepdim=True)).abs().max() <= 1e-13


@pytest.mark.gpu
def test_qnn_noise_module_with_depolarizing_channel():
    from qiddm_b200 import nn
    torch.manual_seed(7)
    m = nn.QNN_noise(64, 4, 3, add_noise=3)
    x = torch.rand(5, 1, 8, 8, dtype=torch.float64, device=m.weights.device)
    W = m.weights.detach().cpu().clone().requires_grad_(True)
    dm = _density_qnn(torch.zeros(1, 4, dtype=torch.float64), W, 4, 3, 0.02)          # input-independent
    ref = (dm @ m.linear_up.weight.detach().cpu().T + m.linear_up.bias.detach().cpu()).expand(5, 64).reshape(5, 1, 8, 8)
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    out = m(x)
    (out * g.to(out.device)).sum().backward()
    assert rel_to_max(out, ref) <= 1e-5
    assert rel_to_max(m.weights.grad, W.grad) <= 1e-4
    m.add_noise = 2                                         # AmplitudeDamping on |0..0>: identical to the noiseless net
    m2 = m(x)
    m.add_noise = 0
    assert rel_to_max(m2, m(x)) <= 1e-7
