"""Generates the golden fixtures under tests/golden/ (run in the build container, where
/root/reference is mounted; the GPU box never sees /root/reference).

* f1_qdense_label14.pt / f1_differn_label14.pt — weights trained by the REAL PennyLane stack, taken
  verbatim from the reference artefact results/emnist.zip (fixture F1), plus the oracle's
  sampler output / stage outputs for them.
* f1_expval_label14.pt — the same for the <Z> families: QIDDM_PL_noise(784,8,6,2) (a4) and QNN(784,8,6) (a5) checkpoints of
  label 14 with the oracle's chain / layer outputs on seeded inputs (`python make_golden.py f1_expval` writes only this one).
* f1_unet_label14.pt — the classical UNetUndirected(3, 8, qdepth=0) checkpoint of label 14 (row a7: the UNet glue around the
  convolutions) + 20 sampler iterations of the product module on the CPU (`python make_golden.py f1_unet`).
* f3_qw_map_logo_ascari.pt — fixture F3 turned into a pin of `qw_map.tanh`: the QDenseUndirected_old(60, 28) checkpoint of
  results_rebuttal_complex_dataset/logo2kplus.zip (label "Ascari"), its recorded per-epoch training losses and the 100 training
  images the reference saved next to it (8-bit PNGs) (`python make_golden.py f3_qw_map`).
* f3_qiddm_pl_logo_sanyo.pt — the same kind of fixture for family a4: QIDDM_PL_noise(784,8,6,2), recorded losses, training
  images (`python make_golden.py f3_qiddm_pl`).
* ref_qconv_literal_forward.pt — OUTPUTS OF THE REFERENCE'S OWN CODE: `/root/reference/nn/qconv.py::_QConv2d_FAST` imported with
  stub `pennylane` / `qw_map` modules (its forward never calls the QNode, nn/qconv.py:71-90, so the stubs are never executed)
  and run on seeded float64 images: forward outputs and autograd image gradients for five layer shapes
  (`python make_golden.py ref_qconv`).  Pins the `reference_forward=True` mode of the product's QConv2d.
* f2_state_dict_contract.json — key names / shapes / dtypes of every shipped checkpoint family (F2).
* stage_vectors.pt — oracle outputs and gradients for seeded inputs of each circuit family.
"""
import io
import json
import sys
import zipfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import qiddm_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference")


def load_ck(z, name):
    return torch.load(io.BytesIO(z.read(name)), weights_only=False, map_location="cpu")


def f1_expval(z):
    """F1 a4 / a5: checkpoints trained through lightning.qubit + expval(PauliZ): weights verbatim + oracle outputs."""
    src_pl = "emnist14/noise_0/QIDDM_PL_noise=8_L=6_N=2_noise_14.pt"
    sd = load_ck(z, src_pl)["model_state_dict"]
    torch.manual_seed(2)
    ang = torch.randn(4, 8, dtype=torch.float64)
    chain = O.qiddm_expval_chain(ang, sd["net.weights1"].double())
    pl = {"weights1": sd["net.weights1"], "linear_up.weight": sd["net.linear_up.weight"], "linear_up.bias": sd["net.linear_up.bias"],
          "angles": ang, "chain_out": chain,
          "image_out": chain @ sd["net.linear_up.weight"].double().T + sd["net.linear_up.bias"].double(), "source": "results/emnist.zip:" + src_pl}
    src_qnn = "emnist14/noise_0/QNN_linear_features=8_qdepth=6_add_noise=0_noise_14.pt"
    sd = load_ck(z, src_qnn)["model_state_dict"]
    x = torch.rand(2, 1, 28, 28, dtype=torch.float64)
    out = O.qnn_forward(x, sd["net.weights"].double(), sd["net.linear_down.weight"], sd["net.linear_down.bias"],
                        sd["net.linear_up.weight"], sd["net.linear_up.bias"])
    qnn = {**{k[4:]: v for k, v in sd.items()}, "x": x, "out": out, "source": "results/emnist.zip:" + src_qnn}
    torch.save({"qiddm_pl": pl, "qnn": qnn}, OUT / "f1_expval_label14.pt")


def f1_unet(z):
    """F1 a7: weights of the reference's classical UNet (torch.nn.Conv2d children) verbatim + the sampler's output."""
    from qiddm_b200 import nn as qnn
    src = "emnist14/noise_0/unet_undirected_d3_s8_d0_noise_14.pt"
    sd = {k[4:]: v for k, v in load_ck(z, src)["model_state_dict"].items()}
    net = qnn.UNetUndirected(3, 8, 0)
    net.load_state_dict(sd)
    net.eval()
    torch.manual_seed(0)
    x0 = torch.rand(2, 1, 28, 28, dtype=torch.float64) * 0.75 + 0.5
    smp = O.sample(lambda v: net(v), x0, 20, goal="noise")
    torch.save({"state_dict": sd, "first_x": x0, "sample": smp, "source": "results/emnist.zip:" + src}, OUT / "f1_unet_label14.pt")


def f3_qw_map():
    """The only shipped checkpoints of a class that calls qw_map.tanh (nn/qdense.py:45) + their training data."""
    import numpy as np
    from PIL import Image
    z = zipfile.ZipFile(REF / "results_rebuttal_complex_dataset/logo2kplus.zip")
    src = "logo2kplus/Ascari/QDenseUndirected_old60_w28_h28_0.pt"
    ck = load_ck(z, src)
    imgs = [np.asarray(Image.open(io.BytesIO(z.read(f"logo2kplus/Ascari/image_0/train_image_{i}.png"))).convert("L"))
            for i in range(1, 101)]
    torch.save({"weights": ck["model_state_dict"]["net.weights"], "loss_values": torch.tensor(ck["loss_values"]),
                "epochs": ck["epochs"], "train_images_u8": torch.tensor(np.stack(imgs), dtype=torch.uint8),
                "source": "results_rebuttal_complex_dataset/logo2kplus.zip:" + src + " + Ascari/image_0/train_image_*.png "
                          "(plt.imsave(cmap='gray') of the training tensors, src/bloodmnist.py:266-268)"},
               OUT / "f3_qw_map_logo_ascari.pt")


def f3_qiddm_pl():
    """QIDDM_PL_noise(784,8,6,2) checkpoint of logo2kplus "Sanyo" + recorded losses + its 100 training images (a4)."""
    import numpy as np
    from PIL import Image
    z = zipfile.ZipFile(REF / "results_rebuttal_complex_dataset/logo2kplus.zip")
    src = "logo2kplus/Sanyo/QIDDM_PL_noise=8_L=6_N=2_5.pt"
    ck = load_ck(z, src)
    imgs = [np.asarray(Image.open(io.BytesIO(z.read(f"logo2kplus/Sanyo/image_0/train_image_{i}.png"))).convert("L"))
            for i in range(1, 101)]
    # the reference's OWN sampler output for this checkpoint: image_{1..10}/step_{1..6}.png = first_x and 5 iterations of
    # Diffusion.sample (goal "data"), clamped to [0, 1] and saved with plt.imsave(cmap="gray") (src/bloodmnist.py:231-278)
    steps = [[np.asarray(Image.open(io.BytesIO(z.read(f"logo2kplus/Sanyo/image_{i}/step_{s}.png"))).convert("L"))
              for i in range(1, 11)] for s in range(1, 7)]
    torch.save({**ck["model_state_dict"], "loss_values": torch.tensor(ck["loss_values"]), "epochs": ck["epochs"],
                "train_images_u8": torch.tensor(np.stack(imgs), dtype=torch.uint8),
                "sample_steps_u8": torch.tensor(np.array(steps), dtype=torch.uint8),
                "source": "results_rebuttal_complex_dataset/logo2kplus.zip:" + src + " + Sanyo/image_0/train_image_*.png"
                          " + Sanyo/image_*/step_*.png"},
               OUT / "f3_qiddm_pl_logo_sanyo.pt")


def ref_qconv():
    """Run the reference's literal `_QConv2d_FAST.forward` (torch + einops only; PennyLane is stubbed, never executed)."""
    import importlib.util
    import types
    qml = types.ModuleType("pennylane")

    class _SEL:
        @staticmethod
        def shape(n_layers, n_wires):
            return (n_layers, n_wires, 3)

    qml.StronglyEntanglingLayers = _SEL
    qml.device = lambda *a, **k: object()
    qml.QNode = lambda *a, **k: (lambda *aa, **kk: (_ for _ in ()).throw(RuntimeError("stub QNode executed")))
    sys.modules.setdefault("pennylane", qml)
    sys.modules.setdefault("qw_map", types.ModuleType("qw_map"))
    spec = importlib.util.spec_from_file_location("ref_qconv", REF / "nn/qconv.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = {}
    for i, (cin, cout, k, pad, hw) in enumerate([(1, 8, 3, 1, (7, 6)), (8, 8, 3, 1, (6, 6)), (4, 1, 1, 0, (5, 5)),
                                                   (3, 4, (3, 2), (1, 0), (6, 7)), (16, 8, 3, 1, (4, 4))]):
        g = torch.Generator().manual_seed(500 + i)
        layer = mod._QConv2d_FAST(cin, cout, kernel_size=k, padding=pad, qdepth=2)
        x = (torch.rand(2, cin, *hw, generator=g, dtype=torch.float64) * 0.4 - 0.1).requires_grad_(True)   # spans both clamp edges
        out = layer(x)
        go = torch.randn(out.shape, generator=g, dtype=torch.float64)
        (out * go).sum().backward()
        cases[f"in{cin}_out{cout}_k{k}_p{pad}"] = {"args": (cin, cout, k, pad), "x": x.detach().clone(), "out": out.detach(),
                                                   "grad_out": go, "grad_x": x.grad.clone()}
    torch.save({"cases": cases, "source": "/root/reference/nn/qconv.py::_QConv2d_FAST.forward (:71-90), pennylane/qw_map stubbed"},
               OUT / "ref_qconv_literal_forward.pt")
    print({k: tuple(v["out"].shape) for k, v in cases.items()})


def main():
    if sys.argv[1:] == ["ref_qconv"]:
        ref_qconv()
        return
    if sys.argv[1:] == ["f3_qw_map"]:
        f3_qw_map()
        return
    if sys.argv[1:] == ["f3_qiddm_pl"]:
        f3_qiddm_pl()
        return
    z = zipfile.ZipFile(REF / "results/emnist.zip")
    if sys.argv[1:] == ["f1_expval"]:
        f1_expval(z)
        return
    if sys.argv[1:] == ["f1_unet"]:
        f1_unet(z)
        return
    f1_expval(z)
    f1_unet(z)
    f3_qw_map()
    f3_qiddm_pl()
    ref_qconv()
    # ---- F1 a1: QDenseUndirected_old_noise(60, 28), label 14 ("O")
    ck = load_ck(z, "emnist14/noise_0/QDenseUndirected_old_noise60_w28_h28_noise0_noise_14.pt")
    W = ck["model_state_dict"]["net.weights"]
    torch.manual_seed(0)
    first_x = (torch.rand(1, 784, dtype=torch.float64) * 0.75 + 0.5).reshape(1, 1, 28, 28)
    smp = O.sample(lambda v: O.qdense_forward(v, W, O.REMAP_TANH), first_x, 40, goal="noise")
    img = smp[0, 0]
    contrast = (img[6:22, 6:22].mean() - (img.sum() - img[6:22, 6:22].sum()) / (784 - 256)).item()
    one = O.qdense_forward(first_x, W, O.REMAP_TANH)
    torch.save({"weights": W, "first_x": first_x, "sample": smp, "contrast": contrast, "one_forward": one,
                "source": "results/emnist.zip:emnist14/noise_0/QDenseUndirected_old_noise60_w28_h28_noise0_noise_14.pt"},
               OUT / "f1_qdense_label14.pt")
    print("qdense contrast", contrast)

    # ---- F1 a3: differN_old_pca(28, 15, 2), label 14 — one forward on fixed angles (PCA excluded)
    ck = load_ck(z, "emnist14/noise_0/differN_old_pca=15_N=2_w28_h28_noise0_noise_14.pt")
    W3 = ck["model_state_dict"]["net.weights"]
    torch.manual_seed(1)
    ang = torch.randn(4, 10, dtype=torch.float64)
    out3 = O.differN_forward(ang, W3.double(), 784)
    torch.save({"weights": W3, "angles": ang, "out": out3,
                "source": "results/emnist.zip:emnist14/noise_0/differN_old_pca=15_N=2_w28_h28_noise0_noise_14.pt"},
               OUT / "f1_differn_label14.pt")

    # ---- F2: state-dict contract of every family shipped for label 14 + tune_results
    contract = {}
    for name in z.namelist():
        if name.startswith("emnist14/") and name.endswith(".pt"):
            sd = load_ck(z, name)["model_state_dict"]
            contract[Path(name).name] = {k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}
    tr = sorted((REF / "tune_results").rglob("*.pt"))
    if tr:
        sd = torch.load(tr[0], weights_only=False, map_location="cpu")["model_state_dict"]
        contract["tune_results:" + tr[0].name.split("_noise_")[0]] = {k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}
    (OUT / "f2_state_dict_contract.json").write_text(json.dumps(contract, indent=1, sort_keys=True))

    # ---- seeded stage vectors (forward + gradients) per family
    fams = {
        "qdense_60x28": (O.desc_qdense(60, 784, O.REMAP_TANH), 3),
        "qdense_pi_tanh_8x8": (O.desc_qdense(10, 64, O.REMAP_PI_TANH), 4),
        "qnn_a_8x8": (O.desc_qnn_a(4, 64), 4),
        "qiddm_ll_6_14": (O.desc_reupload(6, 14, 2), 5),
        "qiddm_pl_8_6": (O.desc_reupload(8, 6, 2), 5),
        "qnn_noise_8_14": (O.desc_reupload(8, 1, 14), 5),
        "differn_10_9_chain": (O.desc_reupload(10, 9, 2, readout=O.READ_PROBS, read_count=10), 3),
        "qconv_8_8_k3": (O.desc_qconv(8, 8, (3, 3), 3), 6),
        "qconv_1_8_k3": (O.desc_qconv(1, 8, (3, 3), 3), 6),
        "ry_reupload_5": (O.desc_reupload(5, 4, 2, enc=O.ENC_RY), 5),
    }
    vec = {}
    for i, (name, (d, B)) in enumerate(fams.items()):
        g = torch.Generator().manual_seed(1000 + i)
        W = (torch.randn(d.n_blocks, d.layers_per_block, d.n_qubits, 3, generator=g, dtype=torch.float64) * 0.4)
        x = (torch.rand(B, d.n_features, generator=g, dtype=torch.float64) if d.init == O.INIT_AMPLITUDE
             else torch.randn(B, d.n_qubits, generator=g, dtype=torch.float64))
        Wr, xr = W.clone().requires_grad_(True), x.clone().requires_grad_(True)
        out = O.run_stage(d, xr, Wr)
        go = torch.randn(out.shape, generator=g, dtype=torch.float64)
        (out * go).sum().backward()
        vec[name] = {"desc": dict(d.__dict__), "weights": W, "x": x, "out": out.detach(), "grad_out": go,
                     "grad_w": Wr.grad, "grad_x": xr.grad}
    torch.save(vec, OUT / "stage_vectors.pt")
    print("wrote", [p.name for p in OUT.iterdir()])


if __name__ == "__main__":
    main()
