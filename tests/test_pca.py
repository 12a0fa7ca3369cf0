"""On-device PCA (qiddm_b200.pca.DevicePCA) against scikit-learn's exact solver: same subspace, same scores up to
the documented sign convention (U-based svd_flip of the reference's scikit-learn 1.1.3)."""
import numpy as np
import pytest
import torch

from qiddm_b200.pca import DevicePCA


def _u_based(t):
    """Flip every column so that its entry of largest magnitude is positive (svd_flip, u_based_decision=True)."""
    t = np.asarray(t, dtype=np.float64)
    idx = np.abs(t).argmax(axis=0)
    return t * np.sign(t[idx, np.arange(t.shape[1])])


@pytest.mark.parametrize("m,p,k", [(10, 784, 8), (10, 784, 10), (80, 784, 8), (40, 64, 6), (300, 64, 8), (30, 4096, 8)])
def test_device_pca_matches_sklearn_full_solver(m, p, k):
    from sklearn.decomposition import PCA
    rng = np.random.default_rng(m * 1000 + p)
    base = rng.random((m, 12)) @ rng.random((12, p))            # image-like: low rank + noise, entries > 0
    x = (base / base.max() + 0.05 * rng.random((m, p))).astype(np.float64)
    sk = PCA(n_components=k, svd_solver="full")
    ref = _u_based(sk.fit_transform(x))
    pca = DevicePCA(k)
    got = pca.fit_transform(torch.from_numpy(x)).numpy()
    scale = np.abs(ref).max()
    keep = sk.singular_values_ > 1e-9 * sk.singular_values_[0]      # a rank-deficient tail has arbitrary directions
    assert np.abs(got[:, keep] - ref[:, keep]).max() <= 1e-8 * scale
    assert np.abs(got[:, ~keep]).max(initial=0.0) <= 1e-6 * scale
    # convention: the largest-magnitude entry of every (non-degenerate) score column is positive
    idx = np.abs(got).argmax(axis=0)
    assert (got[idx, np.arange(k)][keep] > 0).all()
    # inverse_transform / transform round trip equals sklearn's
    inv_ref = sk.inverse_transform(sk.transform(x))
    inv = pca.inverse_transform(pca.transform(torch.from_numpy(x))).numpy()
    assert np.abs(inv - inv_ref).max() <= 1e-8


def test_device_pca_rejects_too_many_components():
    with pytest.raises(ValueError):
        DevicePCA(11).fit_transform(torch.rand(10, 784))


@pytest.mark.gpu
@pytest.mark.parametrize("m", [1, 2, 3, 10, 17, 80, 118])
def test_jacobi_eigh_kernel_matches_lapack(m):
    from qiddm_b200 import _lib as L
    g = torch.Generator().manual_seed(m)
    a = torch.randn(m, max(m, 3) + 5, generator=g, dtype=torch.float64)
    a = (a @ a.T).cuda()
    lam, vec = L.sym_eigh(a)
    ref = torch.linalg.eigvalsh(a.cpu()).flip(0)
    assert (lam.cpu() - ref).abs().max() <= 1e-12 * ref.abs().max()
    assert ((vec.T @ vec).cpu() - torch.eye(m, dtype=torch.float64)).abs().max() <= 1e-12
    assert ((a @ vec) - vec * lam).abs().max().item() <= 1e-11 * ref.abs().max().item()


@pytest.mark.gpu
def test_device_pca_on_gpu_equals_cpu_math():
    x = torch.rand(10, 784, dtype=torch.float64)
    a = DevicePCA(8).fit_transform(x)
    b = DevicePCA(8).fit_transform(x.cuda()).cpu()
    assert (a - b).abs().max() <= 1e-9 * a.abs().max()


def test_grouped_pca_equals_one_pca_per_group():
    """3-D input = independent PCAs per group (the reference's batch-1 semantics for a batch of images)."""
    torch.manual_seed(3)
    x = torch.rand(6, 10, 64, dtype=torch.float64)
    pca = DevicePCA(8)
    got = pca.fit_transform(x)
    inv = pca.inverse_transform(got)
    for g in range(6):
        one = DevicePCA(8)
        ref = one.fit_transform(x[g])
        assert (got[g] - ref).abs().max() <= 1e-10 * ref.abs().max()
        assert (inv[g] - one.inverse_transform(ref)).abs().max() <= 1e-10


def test_module_pca_group_matches_per_image_forward():
    """QIDDM_PL-style reduction with pca_group = tau on a batch of images == the per-image calls of the reference."""
    from qiddm_b200.nn import qdense
    torch.manual_seed(4)
    x = torch.rand(30, 64, dtype=torch.float64)                 # 3 images x tau = 10 rows
    pca = DevicePCA(4)
    grouped = qdense._pca_fit_transform(pca, x, 10)
    for i in range(3):
        ref = DevicePCA(4).fit_transform(x[10 * i:10 * i + 10])
        assert (grouped[10 * i:10 * i + 10] - ref).abs().max() <= 1e-10 * ref.abs().max()
    with pytest.raises(ValueError):
        qdense._pca_fit_transform(pca, x[:25], 10)


@pytest.mark.gpu
def test_batched_jacobi_kernel_and_grouped_module_on_gpu():
    from qiddm_b200 import _lib as L
    from qiddm_b200 import nn
    g = torch.Generator().manual_seed(8)
    a = torch.randn(37, 10, 40, generator=g, dtype=torch.float64)
    a = (a @ a.transpose(1, 2)).cuda()
    lam, vec = L.sym_eigh(a)
    ref = torch.linalg.eigvalsh(a.cpu()).flip(-1)
    assert (lam.cpu() - ref).abs().max() <= 1e-12 * ref.abs().max()
    assert ((a @ vec) - vec * lam.unsqueeze(1)).abs().max().item() <= 1e-10 * ref.abs().max().item()
    # a PL model on 4 images x tau rows with pca_group = tau equals 4 separate forwards
    torch.manual_seed(0)
    m = nn.QIDDM_PL_noise(64, 4, 3, 2).to("cuda", torch.float64)
    m.pca_group = 5
    x = torch.rand(20, 1, 8, 8, dtype=torch.float64, device="cuda")
    with torch.no_grad():
        whole = m(x)
        parts = torch.cat([m(x[5 * i:5 * i + 5]) for i in range(4)])
    assert (whole - parts).abs().max().item() <= 1e-6 * parts.abs().max().item()


def test_host_sklearn_pca_still_plugs_in():
    """`module.pca = sklearn.decomposition.PCA(k)` (or QIDDM_PCA=host) keeps the reference's host round trip."""
    from sklearn.decomposition import PCA
    from qiddm_b200.nn import qdense
    x = torch.rand(10, 64, dtype=torch.float64)
    a = qdense._pca_fit_transform(PCA(n_components=4), x)
    b = qdense._pca_fit_transform(DevicePCA(4), x)
    assert a.shape == b.shape == (10, 4) and a.dtype == torch.float64
    assert (a.abs() - b.abs()).abs().max() <= 1e-9 * b.abs().max()          # same scores up to the sign convention
    with pytest.raises(ValueError):
        qdense._pca_fit_transform(PCA(n_components=4), torch.rand(20, 64, dtype=torch.float64), 10)
