"""CPU tests pinning the oracle: analytic known answers, the committed golden vectors, and (when the
reference tree is mounted) fixture F1 — checkpoints trained by the real PennyLane stack only produce
letters under the restated conventions."""
import io
import math
import zipfile
from pathlib import Path

import pytest
import torch

from conftest import GOLDEN
from oracle import qiddm_oracle as O

REF_ZIP = Path("/root/reference/results/emnist.zip")


def test_probs_sum_to_one_and_zero_weights_permute_basis():
    d = O.StageDesc(n_qubits=6, layers_per_block=7, init=O.INIT_AMPLITUDE, n_features=50, pad_value=0.1,
                    readout=O.READ_PROBS, read_count=64)
    p = O.run_stage(d, torch.rand(5, 50, dtype=torch.float64), torch.randn(1, 7, 6, 3, dtype=torch.float64))
    assert torch.allclose(p.sum(1), torch.ones(5, dtype=torch.float64), atol=1e-12)
    d0 = O.desc_reupload(5, 3, 2, readout=O.READ_PROBS, read_count=32)
    d0.imprimitive = O.IMP_CNOT
    p0 = O.run_stage(d0, torch.randn(4, 5, dtype=torch.float64), torch.zeros(3, 2, 5, 3, dtype=torch.float64))
    assert torch.allclose(p0[:, 0], torch.ones(4, dtype=torch.float64), atol=1e-12)


def test_ry_expval_is_cos_theta_and_rz_on_zero_state_is_invisible():
    d = O.StageDesc(n_qubits=1, layers_per_block=1, readout=O.READ_EXPVAL_Z)
    for th in (0.3, 1.1, 2.5):
        z = O.run_stage(d, None, torch.tensor([[[[0.0, th, 0.0]]]], dtype=torch.float64), batch=1)
        assert abs(z.item() - math.cos(th)) < 1e-12
    # QNN/QNN_noise: RZ once on |0..0> is a global phase -> output independent of the input
    dq = O.desc_reupload(4, 1, 3)
    W = torch.randn(1, 3, 4, 3, dtype=torch.float64)
    a, b = O.run_stage(dq, torch.randn(2, 4, dtype=torch.float64), W), O.run_stage(dq, torch.zeros(2, 4, dtype=torch.float64), W)
    assert torch.allclose(a, b, atol=1e-12)


def test_cnot_direction_and_wire_order():
    """|10> (wire 0 set) --CNOT(0->1)--> |11>; index = b0*2 + b1 (wire 0 = MSB)."""
    st = torch.zeros(1, 4, dtype=O.CDTYPE)
    st[0, 2] = 1
    out = O.apply_ring(st, 2, 1, O.IMP_CNOT)  # ring on 2 wires = CNOT(0,1) then CNOT(1,0)
    # CNOT(0,1): |10>->|11>; CNOT(1,0): |11>->|01>
    assert out[0, 1].abs() == 1
    sign = O.ring_cz_sign(3, 1)
    assert sign[0b110] == -1 and sign[0b101] == -1 and sign[0b111] == -1 and sign[0b100] == 1


def test_parameter_shift_identity_matches_autograd():
    """d f / d theta = (f(theta + pi/2) - f(theta - pi/2)) / 2 for every Rot angle (what the reference's
    diff_method='parameter-shift' would compute, nn/qdense.py:246)."""
    d = O.desc_reupload(3, 2, 2)
    g = torch.Generator().manual_seed(0)
    W = torch.randn(2, 2, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    x = torch.randn(1, 3, generator=g, dtype=torch.float64)
    c = torch.randn(3, generator=g, dtype=torch.float64)
    (O.run_stage(d, x, W) @ c).sum().backward()
    for idx in [(0, 0, 0, 0), (1, 1, 2, 1), (0, 1, 1, 2), (1, 0, 0, 1)]:
        Wp, Wm = W.detach().clone(), W.detach().clone()
        Wp[idx] += math.pi / 2
        Wm[idx] -= math.pi / 2
        ps = 0.5 * ((O.run_stage(d, x, Wp) @ c).sum() - (O.run_stage(d, x, Wm) @ c).sum())
        assert abs(ps.item() - W.grad[idx].item()) < 1e-10


def test_unitary_collapse_equals_gate_path():
    """Cross-implementation (nn/qconv.py:92-126): U from basis states reproduces the gate-by-gate probs."""
    d = O.desc_qconv(2, 4, (3, 3), 3)
    g = torch.Generator().manual_seed(1)
    W = torch.randn(1, 3, d.n_qubits, 3, generator=g, dtype=torch.float64)
    U = O.circuit_unitary(d, W)
    assert torch.allclose(U.conj().T @ U, torch.eye(d.dim, dtype=O.CDTYPE), atol=1e-12)
    x = torch.rand(6, d.n_features, generator=g, dtype=torch.float64)
    psi0 = O.amplitude_embedding(x, d.n_qubits, d.pad_value, d.add_offset)
    p = (psi0 @ U.T).abs() ** 2
    ref = torch.clamp(p * d.post_scale, 0, 1)[:, ::2][:, : d.read_count]
    assert torch.allclose(O.run_stage(d, x, W), ref, atol=1e-12)


def test_golden_stage_vectors_regression():
    vec = torch.load(GOLDEN / "stage_vectors.pt", weights_only=False)
    assert len(vec) >= 10
    for name, v in vec.items():
        d = O.StageDesc(**v["desc"])
        W, x = v["weights"].clone().requires_grad_(True), v["x"].clone().requires_grad_(True)
        out = O.run_stage(d, x, W)
        (out * v["grad_out"]).sum().backward()
        assert torch.allclose(out, v["out"], atol=1e-12), name
        assert torch.allclose(W.grad, v["grad_w"], atol=1e-10), name
        assert torch.allclose(x.grad, v["grad_x"], atol=1e-10), name


def test_golden_f1_qdense_one_forward():
    gold = torch.load(GOLDEN / "f1_qdense_label14.pt", weights_only=True)
    out = O.qdense_forward(gold["first_x"], gold["weights"], O.REMAP_TANH)
    assert torch.allclose(out, gold["one_forward"], atol=1e-12)
    assert gold["contrast"] > 0.45


def test_golden_f1_differn_forward():
    gold = torch.load(GOLDEN / "f1_differn_label14.pt", weights_only=True)
    out = O.differN_forward(gold["angles"], gold["weights"].double(), 784)
    assert torch.allclose(out, gold["out"], atol=1e-12)


@pytest.mark.skipif(not REF_ZIP.exists(), reason="reference artefacts not mounted (GPU box)")
def test_f1_reference_checkpoint_draws_letter_only_with_reference_conventions():
    """Appendix A: label-14 QDense checkpoint -> letter 'O' (centre brighter than border by > 0.3);
    swapping the re-map to pi*tanh destroys it."""
    z = zipfile.ZipFile(REF_ZIP)
    ck = torch.load(io.BytesIO(z.read(
        "emnist14/noise_0/QDenseUndirected_old_noise60_w28_h28_noise0_noise_14.pt")), weights_only=False)
    W = ck["model_state_dict"]["net.weights"]
    torch.manual_seed(0)
    x0 = (torch.rand(1, 784, dtype=torch.float64) * 0.75 + 0.5).reshape(1, 1, 28, 28)

    def contrast(remap, iters):
        img = O.sample(lambda v: O.qdense_forward(v, W, remap), x0, iters, goal="noise")[0, 0]
        return (img[6:22, 6:22].mean() - (img.sum() - img[6:22, 6:22].sum()) / (784 - 256)).item()

    assert contrast(O.REMAP_TANH, 20) > 0.3
    assert contrast(O.REMAP_PI_TANH, 20) < 0.1


def test_golden_f1_expval_families_forward():
    """a4 / a5 regression on weights trained through lightning.qubit + expval(PauliZ) (tests/golden/make_golden.py)."""
    gold = torch.load(GOLDEN / "f1_expval_label14.pt", weights_only=True)
    pl, qnn = gold["qiddm_pl"], gold["qnn"]
    chain = O.qiddm_expval_chain(pl["angles"], pl["weights1"].double())
    assert torch.allclose(chain, pl["chain_out"], atol=1e-12) and chain.abs().max() <= 1 + 1e-12
    out = O.qnn_forward(qnn["x"], qnn["weights"].double(), qnn["linear_down.weight"], qnn["linear_down.bias"],
                        qnn["linear_up.weight"], qnn["linear_up.bias"])
    assert torch.allclose(out, qnn["out"], atol=1e-12)
    # QNN: RZ on |0..0> is a global phase -> the trained output does not depend on the image
    assert torch.allclose(out[0], out[1], atol=1e-12)


def _letter_contrast(img):
    return (img[6:22, 6:22].mean() - (img.sum() - img[6:22, 6:22].sum()) / (784 - 256)).item()


def test_f1_expval_checkpoints_pin_the_sign_of_pauli_z():
    """The label-14 checkpoints of the <Z> families draw their letter through `Diffusion.sample` (contrast +0.5)
    only with <Z> = P(0) - P(1); with the opposite sign the image is flat (contrast 0), and the bias alone gives a
    negative contrast - so the readout sign of families a4 / a5 is pinned by weights the real PennyLane stack trained.
    (Measured in the build container: the same checkpoints are NOT sensitive to the PCA sign convention, the wire order
    of the readout list or the SEL ranges, which therefore stay pinned only through families a1 / a3.)"""
    gold = torch.load(GOLDEN / "f1_expval_label14.pt", weights_only=True)
    qnn, pl = gold["qnn"], gold["qiddm_pl"]
    torch.manual_seed(0)
    x0 = torch.rand(2, 1, 28, 28, dtype=torch.float64) * 0.75 + 0.5

    def qnn_net(sign):
        def f(v):
            a = v.reshape(v.shape[0], -1) @ qnn["linear_down.weight"].T + qnn["linear_down.bias"]
            z = sign * O.run_stage(O.desc_reupload(8, 1, 6), a, qnn["weights"].double()[None])
            return (z @ qnn["linear_up.weight"].T + qnn["linear_up.bias"]).reshape(-1, 1, 28, 28)
        return f

    assert _letter_contrast(O.sample(qnn_net(+1.0), x0, 20, goal="noise")[0, 0]) > 0.4
    assert abs(_letter_contrast(O.sample(qnn_net(-1.0), x0, 20, goal="noise")[0, 0])) < 0.1
    assert _letter_contrast(qnn["linear_up.bias"].reshape(28, 28)) < 0.0

    def pl_net(sign):          # the PCA scores of the sampler are replaced by fixed angles: the letter does not depend on them
        def f(v):
            a = pl["angles"][: v.shape[0]]
            for k in range(2):
                a = sign * O.run_stage(O.desc_reupload(8, 6, 2), a, pl["weights1"].double()[k])
            return (a @ pl["linear_up.weight"].T + pl["linear_up.bias"]).reshape(-1, 1, 28, 28)
        return f

    assert _letter_contrast(O.sample(pl_net(+1.0), x0, 20, goal="noise")[0, 0]) > 0.4
    assert abs(_letter_contrast(O.sample(pl_net(-1.0), x0, 20, goal="noise")[0, 0])) < 0.1


@pytest.mark.skipif(not REF_ZIP.exists(), reason="reference artefacts not mounted (GPU box)")
def test_f4_reference_checkpoints_show_the_cut_circuit_gradient():
    """Fixture F4: the shipped QIDDM_PL_noise checkpoints of different labels carry bit-identical circuit
    weights (the reference detaches the QNode output, so `weights1` never trains) while their `linear_up` differs - the
    behaviour `detach_quantum=True` reproduces in the product modules."""
    z = zipfile.ZipFile(REF_ZIP)
    sds = [torch.load(io.BytesIO(z.read(n)), weights_only=False, map_location="cpu")["model_state_dict"]
           for n in sorted(z.namelist()) if "QIDDM_PL_noise=8_L=6_N=2" in n and n.endswith(".pt")]
    assert len(sds) >= 5
    same = max(([sd for sd in sds if torch.equal(sd["net.weights1"], ref["net.weights1"])] for ref in sds), key=len)
    assert len(same) >= 5
    assert not torch.equal(same[0]["net.linear_up.weight"], same[1]["net.linear_up.weight"])


def test_f3_recorded_training_loss_pins_qw_map_tanh_as_pi_tanh():
    """`QDenseUndirected_old` re-maps its weights with the un-vendored `qw_map.tanh` (nn/qdense.py:45; QConv2d does the same,
    nn/qconv.py:55).  The reference ships one such checkpoint together with its recorded per-epoch training losses AND the
    100 training images (results_rebuttal_complex_dataset/logo2kplus.zip, copied to tests/golden/f3_qw_map_logo_ascari.pt).
    Re-evaluating the training loss of `src/bloodmnist.py:181-191` (batch 1, tau 10, goal "data": the epoch loss is the sum of
    the per-image MSE) with the oracle reproduces the recorded final losses (19.26 ... 19.81) only with qw_map.tanh = pi * tanh
    (19.5); torch.tanh (22.7) and the identity (23.0) stay at the level of the first, untrained epoch (22.0)."""
    gold = torch.load(GOLDEN / "f3_qw_map_logo_ascari.pt", weights_only=True)
    W = gold["weights"].double()
    X = gold["train_images_u8"].double().reshape(100, 784) / 255
    recorded = gold["loss_values"]
    n_img = 40                                       # a subset keeps the CPU suite short; the loss is a per-image mean
    g = torch.Generator().manual_seed(0)
    eps = torch.normal(0.5, 0.2, size=(n_img, 784), generator=g).double()

    def epoch_loss(remap):
        with torch.no_grad():
            per_image = O.diffusion_loss(lambda v: O.qdense_forward(v, W, remap), X[:n_img], eps, 10, (28, 28), goal="data")
        return per_image.item() * 100                # 100 batches of one image per epoch

    trained = recorded[-5:].mean().item()
    assert abs(epoch_loss(O.REMAP_PI_TANH) - trained) < 0.03 * trained
    assert epoch_loss(O.REMAP_TANH) > trained * 1.1
    assert epoch_loss(O.REMAP_NONE) > trained * 1.1


def test_f3_recorded_training_loss_pins_the_expval_readout_order_of_the_reupload_family():
    """Family a4 (QIDDM_PL_noise(784,8,6,2): PCA -> 2 chained stages of RZ re-upload + SEL(CZ) -> [<Z_0> .. <Z_7>] -> linear_up).
    `linear_up` of the shipped checkpoint (logo2kplus "Sanyo") was trained on the expectation values the real lightning.qubit
    produced, in its order; re-evaluating the training loss on the checkpoint's own training images gives 6.2 per epoch
    with the restated readout order (recorded: 4.5 ... 4.9; the images were re-normalised when the reference saved them as
    PNGs), 55 with the list reversed and 30 with CNOT instead of CZ.  (The PCA sign convention and the sign of the angles do
    not move this loss - measured - so they stay unpinned by this fixture.)"""
    from sklearn.decomposition import PCA
    gold = torch.load(GOLDEN / "f3_qiddm_pl_logo_sanyo.pt", weights_only=True)
    W1, wu, bu = gold["weights1"].double(), gold["linear_up.weight"].double(), gold["linear_up.bias"].double()
    X = gold["train_images_u8"].double().reshape(100, 784) / 255

    def epoch_loss(mut):
        def net(v):
            a = torch.tensor(PCA(n_components=8).fit_transform(v.reshape(v.shape[0], -1).numpy()))
            for k in range(2):
                d = O.desc_reupload(8, 6, 2)
                if mut == "cnot":
                    d.imprimitive = O.IMP_CNOT
                a = O.run_stage(d, a, W1[k])
                if mut == "reversed":
                    a = a.flip(1)
            return (a @ wu.T + bu).reshape(-1, 1, 28, 28)
        g = torch.Generator().manual_seed(0)
        tot = 0.0
        with torch.no_grad():
            for i in range(0, 100, 4):               # 25 images, batch 1 each: the PCA runs on one image's tau-ladder
                eps = torch.normal(0.5, 0.2, size=(1, 784), generator=g).double()
                tot += O.diffusion_loss(net, X[i:i + 1], eps, 10, (28, 28), goal="data").item()
        return tot * 4

    recorded = gold["loss_values"][-5:].mean().item()
    ours, rev, cnot = epoch_loss(None), epoch_loss("reversed"), epoch_loss("cnot")
    assert ours < 1.5 * recorded
    assert rev > 3 * ours and cnot > 3 * ours


def test_reference_generated_images_are_reproduced_by_the_restated_sampler():
    """The strongest pin of family a4: images the REAL stack generated.  The reference saved, next to the Sanyo checkpoint,
    `first_x` and the 5 iterations of `Diffusion.sample` it ran with it (10 samples, 8-bit PNGs, each min-max normalised by
    plt.imsave after a clamp to [0, 1]).  Starting from the recoverable part of `first_x` (values above 1 were clamped away:
    replaced by their mean 1.125) the restated sampler lands on the reference's images: after 5 iterations the mean
    absolute difference is 2.3 of 255 grey levels (correlation 0.9997); with the <Z> list reversed it is 57 (0.84)."""
    from sklearn.decomposition import PCA
    gold = torch.load(GOLDEN / "f3_qiddm_pl_logo_sanyo.pt", weights_only=True)
    W1, wu, bu = gold["weights1"].double(), gold["linear_up.weight"].double(), gold["linear_up.bias"].double()
    S = gold["sample_steps_u8"].double()                       # (6 steps, 10 samples, 28, 28)
    x0 = 0.5 + S[0] / 255 * 0.5
    x0[S[0] == 255] = 1.125

    def run(reverse):
        x = x0.reshape(10, 1, 28, 28)
        for _ in range(5):
            a = torch.tensor(PCA(n_components=8).fit_transform(x.reshape(10, -1).numpy()))
            for k in range(2):
                a = O.run_stage(O.desc_reupload(8, 6, 2), a, W1[k])
                a = a.flip(1) if reverse else a
            x = (a @ wu.T + bu).reshape(10, 1, 28, 28)
        img = x[:, 0].clamp(0, 1)
        lo, hi = img.amin(dim=(1, 2), keepdim=True), img.amax(dim=(1, 2), keepdim=True)
        pred = (img - lo) / (hi - lo) * 255
        corr = torch.stack([torch.corrcoef(torch.stack([pred[i].flatten(), S[5][i].flatten()]))[0, 1] for i in range(10)])
        return (pred - S[5]).abs().mean().item(), corr.mean().item()

    diff, corr = run(False)
    assert diff < 4.0 and corr > 0.999
    diff_r, corr_r = run(True)
    assert diff_r > 30 and corr_r < 0.9


def test_noise_ladder_and_training_targets():
    """src/noise.py:105-126 + src/models.py:46-63 layout: '(batch tau) pixels', w_0 = 0, w_last = 1."""
    x = torch.rand(3, 16, dtype=torch.float64)
    eps = torch.rand(3, 16, dtype=torch.float64)
    lad = O.noise_ladder(x, eps, 11).reshape(3, 11, 16)
    assert torch.allclose(lad[:, 0], x) and torch.allclose(lad[:, -1], eps.clamp(0, 1))
    noisy, clean = O.training_targets(x, eps, 10, (4, 4))
    assert noisy.shape == (30, 1, 4, 4) and torch.allclose(noisy[:9], clean[1:10])
