"""The C restatement (oracle/statevec_oracle.c, gate by gate in place like lightning.qubit) against the torch
restatement (oracle/qiddm_oracle.py, batched tensor ops like default.qubit.torch) and the committed golden vectors:
two independent implementations of the conventions of SURVEY.md §8c must agree to 1e-12.  CPU only."""
import pytest
import torch

from conftest import GOLDEN
from oracle import c_oracle as C
from oracle import qiddm_oracle as O


def test_c_oracle_reproduces_every_golden_stage_vector():
    vec = torch.load(GOLDEN / "stage_vectors.pt", weights_only=False)
    assert len(vec) >= 10
    for name, v in vec.items():
        d = O.StageDesc(**v["desc"])
        out = C.run_stage(d, v["x"], v["weights"])
        assert torch.allclose(out, v["out"], atol=1e-12), name


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7])
@pytest.mark.parametrize("imp", [O.IMP_CNOT, O.IMP_CZ])
def test_c_oracle_full_state_matches_torch_oracle(n, imp):
    """Full state (not only probabilities): ring order, CNOT direction, wire order and the Rot decomposition."""
    g = torch.Generator().manual_seed(10 * n + imp)
    d = O.StageDesc(n_qubits=n, n_blocks=3, layers_per_block=max(2, n), enc=O.ENC_RZ, enc_scale=0.7, imprimitive=imp,
                    remap=O.REMAP_PI_TANH, readout=O.READ_STATE)
    W = torch.randn(3, max(2, n), n, 3, generator=g, dtype=torch.float64)
    x = torch.randn(4, n, generator=g, dtype=torch.float64)
    assert torch.allclose(C.run_stage(d, x, W), O.run_stage(d, x, W), atol=1e-12)
    d.enc = O.ENC_RY
    assert torch.allclose(C.run_stage(d, x, W), O.run_stage(d, x, W), atol=1e-12)


def test_c_oracle_unitary_columns_and_reference_checkpoint():
    """Basis-state columns (the collapse of nn/qconv.py:92-126) and the F1 checkpoint forward."""
    d = O.desc_qconv(2, 4, (3, 3), 3)
    g = torch.Generator().manual_seed(5)
    W = torch.randn(1, 3, d.n_qubits, 3, generator=g, dtype=torch.float64)
    db = O.StageDesc(**{**d.__dict__, "init": O.INIT_BASIS, "readout": O.READ_STATE})
    idx = torch.arange(d.dim)
    assert torch.allclose(C.run_stage(db, None, W, basis_index=idx), O.run_stage(db, None, W, basis_index=idx), atol=1e-12)
    gold = torch.load(GOLDEN / "f1_qdense_label14.pt", weights_only=True)
    dq = O.desc_qdense(60, 784, O.REMAP_TANH)
    out = C.run_stage(dq, gold["first_x"].reshape(1, 784), gold["weights"][None])
    assert torch.allclose(out.reshape(gold["one_forward"].shape), gold["one_forward"], atol=1e-12)
    pl = torch.load(GOLDEN / "f1_expval_label14.pt", weights_only=True)["qiddm_pl"]
    a = pl["angles"]
    for k in range(2):
        a = C.run_stage(O.desc_reupload(8, 6, 2), a, pl["weights1"][k])
    assert torch.allclose(a, pl["chain_out"], atol=1e-12)


def test_c_oracle_rejects_bad_descriptors():
    d = O.StageDesc(n_qubits=3, init=O.INIT_AMPLITUDE, n_features=9, readout=O.READ_PROBS, read_count=4)
    with pytest.raises(ValueError):
        C.run_stage(d, torch.rand(2, 9, dtype=torch.float64), torch.zeros(1, 1, 3, 3, dtype=torch.float64))


@pytest.mark.parametrize("family", ["amplitude_probs", "reupload_expval", "ry_probs_chain"])
def test_autograd_gradients_of_the_torch_oracle_match_central_differences_of_the_c_forward(family):
    """Gradients without autograd: d/d(weights, inputs) of a random linear functional of the stage output, by central
    differences on the C forward (h = 1e-5), against the torch oracle's autograd (the golden gradients' source)."""
    g = torch.Generator().manual_seed({"amplitude_probs": 1, "reupload_expval": 2, "ry_probs_chain": 3}[family])
    if family == "amplitude_probs":
        d = O.StageDesc(n_qubits=4, layers_per_block=3, init=O.INIT_AMPLITUDE, n_features=11, pad_value=0.3, add_offset=0.1,
                        imprimitive=O.IMP_CNOT, remap=O.REMAP_TANH, readout=O.READ_PROBS, read_count=6, read_stride=2,
                        post_scale=8.0)
        x = torch.rand(2, 11, generator=g, dtype=torch.float64)
    elif family == "reupload_expval":
        d = O.desc_reupload(3, 3, 2)
        x = torch.randn(2, 3, generator=g, dtype=torch.float64)
    else:
        d = O.desc_reupload(4, 2, 2, enc=O.ENC_RY, readout=O.READ_PROBS, read_count=4)
        x = torch.randn(2, 4, generator=g, dtype=torch.float64)
    W = torch.randn(d.n_blocks, d.layers_per_block, d.n_qubits, 3, generator=g, dtype=torch.float64) * 0.5
    c = torch.randn(2, d.n_out, generator=g, dtype=torch.float64)
    Wr, xr = W.clone().requires_grad_(True), x.clone().requires_grad_(True)
    (O.run_stage(d, xr, Wr) * c).sum().backward()
    f = lambda xv, wv: (C.run_stage(d, xv, wv) * c).sum().item()
    h = 1e-5
    for t, grad, is_w in ((W, Wr.grad, True), (x, xr.grad, False)):
        flat = t.reshape(-1)
        for i in range(0, flat.numel(), max(1, flat.numel() // 12)):     # a dozen coordinates of each
            tp, tm = flat.clone(), flat.clone()
            tp[i] += h
            tm[i] -= h
            fd = ((f(x, tp.reshape(t.shape)) - f(x, tm.reshape(t.shape))) if is_w
                  else (f(tp.reshape(t.shape), W) - f(tm.reshape(t.shape), W))) / (2 * h)
            assert abs(fd - grad.reshape(-1)[i].item()) < 1e-8 * max(1.0, abs(fd)) + 1e-9, (family, is_w, i)


def test_c_adjoint_gradients_reproduce_the_golden_autograd_gradients():
    """The adjoint method in C (no autograd anywhere) against the golden gradients the torch oracle's autograd produced."""
    vec = torch.load(GOLDEN / "stage_vectors.pt", weights_only=False)
    for name, v in vec.items():
        d = O.StageDesc(**v["desc"])
        gw, gx = C.stage_grads(d, v["x"], v["weights"], v["grad_out"])
        assert torch.allclose(gw, v["grad_w"], atol=1e-10), name
        assert torch.allclose(gx, v["grad_x"], atol=1e-10), name
