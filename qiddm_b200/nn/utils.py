"""UNet glue helpers (reference `nn/utils.py:7-74`); QASM/Qiskit export (:77-129) is out of scope."""
import math
import warnings

import einops
import torch


def autocrop(x, y):
    """Center crop the larger of (x, y) to the smaller one's spatial size.  nn/utils.py:7-19."""
    xs, ys = x.shape, y.shape
    if xs > ys:
        warnings.warn("x is larger than y. Cropping x to match y")
        return autocrop(y, x)
    top, left = (ys[2] - xs[2]) // 2, (ys[3] - xs[3]) // 2
    return x, y[:, :, top:(ys[2] + xs[2]) // 2, left:(ys[3] + xs[3]) // 2]


def autopad(x, y):
    """Zero-pad y to the spatial size of x (ceil on the leading side).  nn/utils.py:22-39."""
    xs, ys = x.shape, y.shape
    if xs < ys:
        warnings.warn("x is smaller than y. Padding x to match y")
        return autopad(y, x)
    dh, dw = xs[2] - ys[2], xs[3] - ys[3]
    pad = (math.ceil(dw / 2), math.floor(dw / 2), math.ceil(dh / 2), math.floor(dh / 2))
    return x, torch.nn.functional.pad(y, pad, mode="constant", value=0)


def get_label_embedding(labels: torch.Tensor, width: int, height: int):
    """Sinusoidal label mask 0.1*sin(label + col/20), broadcast over rows.  nn/utils.py:42-56,74."""
    pos = torch.arange(width, device=labels.device) / 20
    mask = torch.sin(labels[:, None] + pos[None, :]) * 0.1
    return einops.repeat(mask, "b w -> b 1 w h", h=height)
