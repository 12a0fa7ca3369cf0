"""qiddm_b200 — B200-native (sm_100a) batched state-vector kernels behind the QIDDM quantum layers.

Public surface mirrors the reference's hot path only: `qiddm_b200.nn` (QDense*/QIDDM*/QNN*/differN*,
QConv2d, UNet callers), `qiddm_b200.models.Diffusion`, `qiddm_b200.noise`, and the data-parallel
training helper in `qiddm_b200.train`.  All quantum arithmetic runs in libqiddm_b200.so (C ABI in
include/qiddm.h); there is no CPU fallback."""
from . import _lib, functional, models, nn, noise  # noqa: F401
from ._lib import Plan, QiddmError, StageSpec, launch_count, load_library  # noqa: F401

__version__ = "0.1.0"
