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
