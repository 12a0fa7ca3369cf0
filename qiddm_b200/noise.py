"""Noise ladder used by the diffusion training step (reference `src/noise.py:105-126`; the other
schedules there are unused by every driver)."""
import torch


def add_normal_noise_multiple(data: torch.Tensor, tau: int, decay_mod: float = 1.0, eps: torch.Tensor = None):
    """x_t = clamp(x (1 - w_t) + eps w_t, 0, 1), w_t = (t/(tau-1))**decay_mod, eps ~ N(0.5, 0.2) drawn once
    per sample (float32 draw as in the reference).  Returns ((batch tau), pixels), batch-major.
    `eps` may be passed in for reproducible tests; everything stays on `data.device`."""
    if data.dim() == 1:
        data = data.unsqueeze(0)
    batch, pixels = data.shape
    if eps is None:
        eps = torch.normal(mean=0.5, std=0.2, size=(batch, pixels), device=data.device)
    w = torch.linspace(0, 1, tau, device=data.device) ** decay_mod
    w = (w / w.max()).to(data.dtype)[None, :, None]                      # (1, tau, 1)
    noisy = data[:, None, :] * (1 - w) + eps.to(data.device)[:, None, :] * w
    return noisy.clamp(0, 1).reshape(batch * tau, pixels)
