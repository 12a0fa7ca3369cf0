"""CPU tests of the host-side mirror of the reference interface: constructors, state_dict contract (F2),
save_name strings, noise ladder, Diffusion wrapper semantics, UNet glue helpers."""
import json
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import qiddm_oracle as O


def test_state_dict_contract_matches_shipped_checkpoints():
    """F2: key names / shapes of the reference checkpoints (results/emnist.zip, tune_results) load unchanged."""
    from qiddm_b200 import models, nn
    contract = json.loads((GOLDEN / "f2_state_dict_contract.json").read_text())
    builders = {
        "QDenseUndirected_old_noise60": lambda: nn.QDenseUndirected_old_noise(60, 28),
        "differN_old_pca=15": lambda: nn.differN_noise(28, 15, 2),
        "QIDDM_PL_noise=8": lambda: nn.QIDDM_PL_noise(784, 8, 6, 2),
        "QNN_linear_features=8": lambda: nn.QNN_noise(784, 8, 6),
        "unet_undirected_d3_s8_d0": lambda: nn.UNetUndirected(3, 8, 0),
        "tune_results:differN_noise=9": lambda: nn.differN_noise_befor(28, 9, 2),
    }
    seen = 0
    for fname, keys in contract.items():
        for prefix, build in builders.items():
            if fname.startswith(prefix):
                diff = models.Diffusion(build(), None, "noise", (28, 28))
                sd = diff.state_dict()
                want = {k: v for k, v in keys.items() if "num_batches_tracked" not in k}
                have = {k: list(v.shape) for k, v in sd.items() if "num_batches_tracked" not in k}
                assert set(want) == set(have), (fname, set(want) ^ set(have))
                for k, (shape, _) in want.items():
                    assert have[k] == shape, (fname, k)
                seen += 1
    assert seen >= 6


def test_save_names_and_reprs_follow_the_reference():
    from qiddm_b200 import nn
    assert nn.QDenseUndirected_old(60, 28).save_name() == "QDenseUndirected_old60_w28_h28"
    assert nn.QDenseUndirected_old_noise(60, 28).save_name() == "QDenseUndirected_old_noise60_w28_h28_noise0"
    assert nn.QIDDM_LL_noise(784, 6, 14, 2).save_name() == "QIDDM_LL_noise=6_L=14_N=2"
    assert nn.QIDDM_PL_noise(784, 8, 6, 2).save_name() == "QIDDM_PL_noise=8_L=6_N=2"
    assert nn.QNN_noise("28 * 28", 8, 14).save_name() == "QNN_linear_features=8_qdepth=14_add_noise=0"
    assert nn.differN_noise(28, 9, 2).save_name() == "differN_old_pca=9_N=2_w28_h28_noise0"
    assert nn.differN_noise_befor(28, 9, 2).save_name() == "differN_noise=9_N=2_w28_h28"
    assert nn.UNetUndirected(3, 8, 3).save_name() == "unet_undirected_d3_s8_d3"
    assert nn.UNetUndirectedS(2, 4, 2).save_name() == "unet_s_undirected_d2_s4_d2"
    assert repr(nn.QConv2d(8, 8)) == "QConv2d(8, 8, kernel_size=(3, 3), padding=(1, 1), wires=7)"
    assert nn.QIDDM_LL_noise(784, 6, 14, 2).weights1.shape == (2, 14, 2, 6, 3)
    assert nn.QIDDM_bias_false(64, 4, 2, 1).weights1.shape == (1, 2, 3, 4, 3)
    assert nn.QIDDM_A_sameN(8, 3, 2).weights.shape == (3, 2, 6, 3)


def test_all_27_dense_classes_construct():
    from qiddm_b200 import nn
    from qiddm_b200.nn import qdense
    assert len(qdense.__all__) == 27
    args = {"QDenseUndirected_old": (4, 8), "QDenseUndirected_old_noise": (4, 8), "QNN_A": (4, 8),
            "QNN_noise": (64, 4, 2), "QNN": (64, 4, 2)}
    for name in qdense.__all__:
        cls = getattr(nn, name)
        if name in args:
            m = cls(*args[name])
        elif name.startswith("differN") or name == "QIDDM_A_sameN":
            m = cls(8, 2, 2)
        elif name.startswith("QIDDM_A_differN"):
            m = cls(8, 2, 2)
        else:
            m = cls(64, 4, 2, 2)
        assert callable(m.qnode) and isinstance(m.save_name(), str)


def test_qconv_wire_count_rule_and_noise_guard():
    from qiddm_b200 import nn
    cases = {(1, 8, 3): 4, (8, 8, 3): 7, (16, 16, 3): 8, (32, 32, 3): 9, (32, 16, 1): 5, (8, 1, 1): 3, (1, 1, 1): 1}
    for (cin, cout, k), wires in cases.items():
        assert nn.QConv2d(cin, cout, kernel_size=k, padding=k // 2).wires == wires
    assert nn.QIDDM_LL_noise(64, 4, 2, 2, add_noise=2).add_noise == 2     # mid-circuit channels: density-matrix path (inference)
    with pytest.raises(NotImplementedError):
        nn.QIDDM_LL_noise(64, 4, 2, 2, add_noise=4)
    with pytest.raises(NotImplementedError):
        nn.QDenseUndirected_old_noise(4, 8, add_noise=4)
    nn.QDenseUndirected_old_noise(4, 8, add_noise=1)       # PhaseShift before probs: a no-op
    nn.QDenseUndirected_old_noise(4, 8, add_noise=3)       # channels right before probs(): exact readout map (channels.py)


def test_noise_ladder_matches_oracle_and_reference_layout():
    from qiddm_b200 import noise
    x = torch.rand(3, 16, dtype=torch.float64)
    eps = torch.rand(3, 16, dtype=torch.float64)
    ours = noise.add_normal_noise_multiple(x, tau=11, decay_mod=3.0, eps=eps)
    ref = O.noise_ladder(x, eps, 11, 3.0)
    assert ours.shape == (33, 16) and torch.allclose(ours, ref, atol=1e-12)
    torch.manual_seed(0)
    a = noise.add_normal_noise_multiple(x[0], tau=4)
    assert a.shape == (4, 16) and torch.allclose(a[0], x[0])


def test_diffusion_wrapper_with_a_classical_net_on_cpu():
    """Diffusion keeps the reference semantics (backward inside forward, verbose outputs, sampler layout)."""
    from qiddm_b200 import models, noise, nn
    torch.manual_seed(0)
    net = nn.UNetUndirected(depth=2, start_channels=4, qdepth=0)
    diff = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (8, 8), torch.nn.MSELoss()).double()
    diff.train()
    x = torch.rand(2, 64, dtype=torch.float64)
    loss, recon = diff(x=x, T=5, verbose=True)
    assert recon.shape == (10, 1, 8, 8) and loss.dim() == 0
    assert all(p.grad is not None for p in diff.parameters())
    diff2 = models.Diffusion(net, noise.add_normal_noise_multiple, "noise", (8, 8), torch.nn.MSELoss()).double()
    diff2.train()
    (l2,) = diff2(x=x, T=5)
    assert l2.dim() == 0 and diff2.save_name() == "unet_undirected_d2_s4_d0_noise"
    diff2.eval()
    grid = diff2(torch.rand(3, 1, 8, 8, dtype=torch.float64), n_iters=4)
    assert grid.shape == (5 * 8, 3 * 8)
    last = diff2(torch.rand(3, 1, 8, 8, dtype=torch.float64), n_iters=4, only_last=True)
    assert last.shape == (3, 1, 8, 8) and last.min() >= 0 and last.max() <= 1


def test_f1_reference_unet_checkpoint_draws_its_letter_only_with_the_reference_wiring():
    """Row a7 pinned by fixture F1: the reference's trained classical UNet (results/emnist.zip, label 14, weights verbatim
    in tests/golden/f1_unet_label14.pt) loaded into the product `UNetUndirected(3, 8, qdepth=0)` draws its letter through
    `Diffusion.sample` (centre brighter than border by > 0.35); with the skip concatenation swapped ([down, up] instead of
    [up, down], nn/unet.py:73) the letter is gone."""
    from conftest import GOLDEN
    from qiddm_b200 import models, noise, nn
    import qiddm_b200.nn.unet as U
    gold = torch.load(GOLDEN / "f1_unet_label14.pt", weights_only=True)
    net = nn.UNetUndirected(3, 8, 0)
    net.load_state_dict(gold["state_dict"])
    diff = models.Diffusion(net, noise.add_normal_noise_multiple, "noise", (28, 28), torch.nn.MSELoss()).double()
    diff.eval()

    def contrast(img):
        return (img[6:22, 6:22].mean() - (img.sum() - img[6:22, 6:22].sum()) / (784 - 256)).item()

    out = diff.sample(20, first_x=gold["first_x"], only_last=True)
    assert torch.allclose(out, gold["sample"], atol=1e-8)
    assert min(contrast(out[i, 0]) for i in range(out.shape[0])) > 0.35

    def swapped(self, from_down, from_up):
        from_up = self.up_conv(from_up)
        from_down, from_up = U.autopad(from_down.double(), from_up.double())
        return self.net(torch.cat([from_down, from_up], dim=1).double())

    orig = U.UpBlock.forward
    U.UpBlock.forward = swapped
    try:
        bad = diff.sample(100, first_x=gold["first_x"], only_last=True)
    finally:
        U.UpBlock.forward = orig
    assert max(abs(contrast(bad[i, 0])) for i in range(bad.shape[0])) < 0.2


def test_unet_glue_helpers():
    from qiddm_b200.nn import autocrop, autopad, get_label_embedding
    big, small = torch.zeros(1, 1, 7, 7), torch.ones(1, 1, 4, 5)
    _, padded = autopad(big, small)
    assert padded.shape == big.shape and padded.sum() == 20
    assert padded[0, 0, 2, 1] == 1 and padded[0, 0, 1, 1] == 0     # ceil on the leading side
    x, cropped = autocrop(small, big)
    assert cropped.shape[2:] == (4, 5)
    m = get_label_embedding(torch.tensor([0.0, 1.0]), 6, 4)
    assert m.shape == (2, 1, 6, 4) and torch.allclose(m[1, 0, :, 0], 0.1 * torch.sin(1 + torch.arange(6) / 20))


def test_shard_batch_covers_everything_once():
    from qiddm_b200.train import shard_batch
    x = torch.arange(11)
    for world in (1, 2, 3, 4, 8):
        parts = [shard_batch(x, r, world) for r in range(world)]
        assert torch.equal(torch.cat(parts), x)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_path_dispatch_cost_model_against_the_measured_sweep():
    """Plan.use_gemm (gate-by-gate vs unitary collapse) replayed against the committed B200 sweep
    (profiles/r2_configs_sweep.jsonl, scripts/bench_configs.py --what sweep, round 2: gate path with the psi_final
    hand-over): wrong on few points, and never expensively."""
    import json
    from pathlib import Path
    from qiddm_b200 import _lib as L

    class FakePlan(L.Plan):                     # the rule only needs the descriptor
        def __init__(self, spec):
            self.spec = spec

        def gemm_supported(self):
            return True

    rows = [json.loads(l) for l in (Path(__file__).resolve().parent.parent / "profiles" / "r2_configs_sweep.jsonl").open()
            if l.startswith("{")]
    pts = {}
    for r in rows:
        if r["what"] == "sweep" and r["family"] == "qdense":
            pts.setdefault((r["n"], r["depth"], r["batch"]), {})[r["path"]] = r["fwd_bwd_ms"]
    both = {k: v for k, v in pts.items() if len(v) == 2}
    assert len(both) >= 100
    wrong, worst = 0, 1.0
    for (n, depth, batch), v in both.items():
        feats = 784 if n == 10 else 1 << n
        spec = L.StageSpec(n_qubits=n, layers_per_block=depth, init=L.INIT_AMPLITUDE, n_features=feats, read_count=feats)
        pick = "gemm" if FakePlan(spec).use_gemm(batch) else "gate"
        best = min(v, key=v.get)
        if pick != best:
            wrong += 1
            worst = max(worst, v[pick] / v[best])
    assert wrong <= 0.05 * len(both) and worst <= 1.1, (wrong, worst)
    # forced paths and the re-upload families
    spec = L.StageSpec(n_qubits=10, layers_per_block=60, init=L.INIT_AMPLITUDE, n_features=784, read_count=784, path=L.PATH_GATE)
    assert not FakePlan(spec).use_gemm(1 << 20)
    spec = L.StageSpec(n_qubits=10, layers_per_block=60, init=L.INIT_AMPLITUDE, n_features=784, read_count=784, path=L.PATH_GEMM)
    assert FakePlan(spec).use_gemm(1)


def test_fused_step_dispatch_conditions_on_the_host():
    """`Diffusion._fused_step` only hands the training step to the library for nets that offer `fused_mse_step`, the reference's
    ladder and MSELoss, and non-verbose calls; a QDense net declines CPU tensors itself (no CPU path anywhere)."""
    import torch
    from qiddm_b200 import models, nn, noise
    net = nn.QDenseUndirected_old_noise(2, 4)
    diff = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (4, 4), torch.nn.MSELoss())
    x = torch.rand(3, 16)
    assert diff._fused_step(x, {"T": 4}) is None                       # CPU tensor: the net declines
    assert diff._fused_step(x, {"T": 4, "verbose": True}) is None
    diff2 = models.Diffusion(net, lambda d, tau, decay_mod: d, "data", (4, 4), torch.nn.MSELoss())
    assert diff2._fused_step(x, {"T": 4}) is None                      # custom noise schedule
    diff3 = models.Diffusion(net, noise.add_normal_noise_multiple, "data", (4, 4), torch.nn.L1Loss())
    assert diff3._fused_step(x, {"T": 4}) is None                      # another loss
    diff4 = models.Diffusion(torch.nn.Linear(16, 16), noise.add_normal_noise_multiple, "data", (4, 4), torch.nn.MSELoss())
    assert diff4._fused_step(x, {"T": 4}) is None                      # a net without the hook
    os.environ["QIDDM_FUSED_STEP"] = "0"
    try:
        assert diff._fused_step(x, {"T": 4}) is None
    finally:
        os.environ.pop("QIDDM_FUSED_STEP", None)


def test_qconv_dispatch_direct_convolution_gemm_or_gate():
    """Plan.use_collapse_qconv: every layer shape of UNetUndirected(3, 8, 3) except the 32-output-channel ones has the direct
    fp32 convolution (csrc/qiddm_conv.cu) and takes it once the call has at least 2^n patches; an explicit PATH_GEMM / PATH_GATE
    is honoured; other windows fall back to the GEMM-vs-gate cost model.  Host logic only (the library answers without a GPU)."""
    import dataclasses
    from qiddm_b200 import _lib as L
    from qiddm_b200 import nn

    def plan_of(cin, cout, k, path=L.PATH_AUTO):
        m = nn.QConv2d(cin, cout, kernel_size=k, padding=k // 2, qdepth=3)
        return L.Plan.get(dataclasses.replace(m._spec(), path=path)), m

    for cin, cout, k, hw in ((1, 8, 3, 28), (8, 8, 3, 28), (16, 8, 3, 28), (16, 8, 1, 28), (8, 1, 1, 28), (8, 16, 3, 14),
                             (16, 16, 3, 14), (32, 16, 3, 14), (32, 16, 1, 14)):
        plan, m = plan_of(cin, cout, k)
        u = L.UnfoldDesc(cin, hw, hw, k, k, k // 2, k // 2)
        assert plan.qconv_direct(u), (cin, cout, k)
        assert plan.use_collapse_qconv(u, plan.spec.dim)             # as many patches as basis columns: direct
        assert not plan.use_collapse_qconv(u, plan.spec.dim - 1)     # fewer: gate by gate (no collapse to pay for)
        assert not plan_of(cin, cout, k, L.PATH_GATE)[0].use_collapse_qconv(u, 10 ** 6)
        pg = plan_of(cin, cout, k, L.PATH_GEMM)[0]
        assert not pg.qconv_direct(u) and pg.use_collapse_qconv(u, 1)
    for cin, cout, k, hw in ((16, 32, 3, 7), (32, 32, 3, 7)):            # N = 64 rows of U: tcgen05 GEMM or gate path
        plan, m = plan_of(cin, cout, k)
        u = L.UnfoldDesc(cin, hw, hw, k, k, 1, 1)
        assert not plan.qconv_direct(u)
        assert plan.use_collapse_qconv(u, 31360) == plan.use_gemm(31360)
    plan, m = plan_of(8, 8, 3)
    assert not plan.qconv_direct(L.UnfoldDesc(8, 28, 28, 3, 3, 0, 0))       # not "same" padding
    assert not plan.qconv_direct(L.UnfoldDesc(8, 4, 600, 3, 3, 1, 1))       # a row wider than a 512-pixel band


def test_qconv_direct_support_answers_over_a_sweep_of_shapes():
    """qiddm_qconv_direct_supported (the band / tile chooser behind it) over many layer shapes: it answers without a GPU, says yes
    exactly for 1 x 1 / 3 x 3 "same" windows with <= 16 output channels whose band fits (a row of at most 512 pixels, the weight-
    gradient kernel's warp layout), and the sizes it reports for such layers cover the saved rows and the dL/dY rows."""
    import ctypes as C
    import random
    from qiddm_b200 import _lib as L
    from qiddm_b200 import nn
    lib = L.load_library()
    rng = random.Random(0)
    seen_yes = seen_no = 0
    for _ in range(120):
        cin, cout = rng.choice([1, 2, 3, 8, 16, 32, 48]), rng.choice([1, 2, 5, 8, 16, 24, 32])
        k = rng.choice([1, 3, 5])
        h, w = rng.choice([1, 7, 14, 28, 64, 100]), rng.choice([1, 7, 14, 28, 64, 511, 512, 513])
        if cin * k * k > 1024:
            continue
        m = nn.QConv2d(cin, cout, kernel_size=k, padding=k // 2, qdepth=2)
        if 2 * cout > 2 ** m.wires:          # `[:, ::2][:, :out_channels]` of nn/qconv.py:69-70 needs 2 out_channels amplitudes
            continue
        plan = L.Plan.get(m._spec())
        u = L.UnfoldDesc(cin, h, w, k, k, k // 2, k // 2)
        ok = bool(lib.qiddm_qconv_direct_supported(plan.handle, C.byref(u)))
        n_out = plan.spec.read_count
        np_ = 4 if 2 * n_out <= 4 else (16 if 2 * n_out <= 16 else 32)
        # shared-memory footprints of the thinnest band (one row): weights + image tile, + the dL/dY rows
        tc = (w + k - 1) | 1
        cs = k * tc + ((3 - (k * tc) % 32) + 32) % 32
        tile, wdf = (cin * cs + 3) & ~3, ((cin * k * k + 1) * np_ + 3) & ~3
        ct = 16 if cin > 8 else 8
        fits = (4 * (wdf + tile + w * (np_ + 4) + 256) <= 200 * 1024
                and 4 * (-(-cin // ct) * ct * k * k * np_ + k * (w + k - 1) * (np_ + 4)) <= 200 * 1024)
        expect = (m.wires >= 3 and k in (1, 3) and 2 * n_out <= 32 and w <= 512 and fits
                  and -(-cin * k // 32) * (2 if 2 * n_out > 16 else 1) <= 8)
        assert ok == expect, (cin, cout, k, h, w, m.wires, n_out)
        if ok:
            seen_yes += 1
            n_img = 3
            rows = n_img * h * w * (np_ + 4) * 4
            assert lib.qiddm_qconv_gemm_saved_bytes(plan.handle, C.byref(u), n_img) >= rows
            assert lib.qiddm_qconv_gemm_workspace_bytes(plan.handle, C.byref(u), n_img) >= rows
        else:
            seen_no += 1
    assert seen_yes >= 20 and seen_no >= 20
