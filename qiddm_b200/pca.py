"""On-device PCA for the `*_PL*` / `differN*pca*` / `QIDDM_PP*` modules.

The reference re-fits `sklearn.decomposition.PCA(n_components=k)` on EVERY forward call on the current batch
through a numpy round trip (nn/qdense.py:456-458, :1429).  `DevicePCA` keeps that per-call semantics but stays on
the GPU: centre the (m, P) batch, eigen-decompose its m x m Gram matrix with the library's single-CTA Jacobi kernel
(`qiddm_sym_eigh_f64`, float64, no host sync -> CUDA-graph capturable), scores = U * S.

Conventions follow the reference's pinned scikit-learn 1.1.3 (requirements.txt:76) with the exact ("full") solver:
`U, S, Vt = svd(X - mean)`, then `svd_flip(U, Vt)` with the U-BASED sign decision (the entry of largest magnitude in
every column of U is made positive), `fit_transform = U[:, :k] * S[:k]`.  (scikit-learn >= 1.5 switched to the V-based
decision, so a newer host sklearn differs from the reference by per-component signs; reference 1.1.3 would also pick
its `randomized` solver when 500 < max(m, P) and k < 0.8 min(m, P) — a stochastic approximation of the same
quantity, not reproduced.)  Batches with m > P or m above the kernel's limit use the P x P covariance with
`torch.linalg.eigh` on the device (not capturable)."""
from __future__ import annotations

import torch

from . import _lib as L


class DevicePCA:
    def __init__(self, n_components: int):
        self.n_components = n_components
        self.mean_ = None
        self.components_ = None          # (k, P), rows = principal axes (signs consistent with the scores)
        self.singular_values_ = None

    @staticmethod
    def _eigh_desc(a: torch.Tensor):
        if a.is_cuda and a.shape[-1] <= L.load_library().qiddm_sym_eigh_max_dim():
            return L.sym_eigh(a)
        if a.is_cuda and torch.cuda.is_current_stream_capturing():
            raise L.QiddmError(
                f"DevicePCA: a {a.shape[-1]} x {a.shape[-1]} eigenproblem exceeds the capturable Jacobi kernel "
                f"(<= {L.load_library().qiddm_sym_eigh_max_dim()}); torch.linalg.eigh synchronises and cannot be captured in a "
                "CUDA graph.  Set module.pca_group = tau (one PCA per image's tau-ladder, the reference's batch-1 semantics) "
                "or run the step eagerly.")
        lam, vec = torch.linalg.eigh(a)                       # ascending
        return lam.flip(-1), vec.flip(-1)

    @staticmethod
    def _u_based_signs(u: torch.Tensor) -> torch.Tensor:
        """(..., m, k) -> (..., 1, k): sign of the entry of largest magnitude in every column (svd_flip, U-based)."""
        idx = u.abs().argmax(dim=-2, keepdim=True)
        signs = torch.sign(u.gather(-2, idx))
        return torch.where(signs == 0, torch.ones_like(signs), signs)

    def fit_transform(self, x: torch.Tensor) -> torch.Tensor:
        """(m, P) -> (m, k) float64 scores on x.device.  A 3-D input (G, m, P) is G independent PCAs (one per group, e.g.
        one per image's tau-ladder: the reference's batch-1 semantics for a batch of G images) solved in one launch."""
        x = x.detach().to(torch.float64)
        m, p = x.shape[-2], x.shape[-1]
        k = self.n_components
        if k > min(m, p):
            raise ValueError(f"n_components={k} must be between 0 and min(n_samples, n_features)={min(m, p)}")
        self.mean_ = x.mean(dim=-2, keepdim=True)
        xc = x - self.mean_
        xt = xc.transpose(-1, -2)
        if m <= p:
            lam, u = self._eigh_desc(xc @ xt)                 # Gram matrix of the centred rows
            s = lam[..., :k].clamp_min(0).sqrt().unsqueeze(-2)        # (..., 1, k)
            u = u[..., :, :k]
            u = u * self._u_based_signs(u)
            scores = u * s
            # V_k^T = S^-1 U^T Xc; a rank-deficient tail (k >= m: the centred batch has rank m - 1) gets a zero axis
            inv_s = torch.where(s > 1e-12 * s[..., :1], 1.0 / s.clamp_min(1e-300), torch.zeros_like(s))
            self.components_ = (u * inv_s).transpose(-1, -2) @ xc
        else:
            lam, v = self._eigh_desc(xt @ xc)                 # covariance (un-normalised)
            s = lam[..., :k].clamp_min(0).sqrt().unsqueeze(-2)
            v = v[..., :, :k]
            scores = xc @ v
            signs = self._u_based_signs(scores)               # argmax |U| = argmax |U S| per column
            scores = scores * signs
            self.components_ = (v * signs).transpose(-1, -2)
        self.singular_values_ = s.squeeze(-2)
        if x.dim() == 2:
            self.mean_ = self.mean_.squeeze(0)
        return scores

    def fit(self, x: torch.Tensor) -> "DevicePCA":
        self.fit_transform(x)
        return self

    def transform(self, x: torch.Tensor) -> torch.Tensor:
        """(m, P) -> (m, k): (x - mean_) @ components_.T with the fitted basis."""
        x = x.detach().to(torch.float64)
        return (x - self.mean_.to(x.device)) @ self.components_.to(x.device).transpose(-1, -2)

    def inverse_transform(self, scores: torch.Tensor) -> torch.Tensor:
        """(m, k) -> (m, P): scores @ components_ + mean_ (sklearn `PCA.inverse_transform`, whiten=False)."""
        return scores.to(torch.float64) @ self.components_ + self.mean_
