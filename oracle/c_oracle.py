"""ctypes front-end of oracle/statevec_oracle.c — TEST INFRASTRUCTURE ONLY (see the header of that file).

`build()` compiles the C restatement with gcc into oracle/_build/ (git-ignored, travels with the repo snapshot);
`run_stage(desc, x, weights, ...)` has the signature of `oracle.qiddm_oracle.run_stage` (forward only)."""
from __future__ import annotations

import ctypes
import shutil
import subprocess
from pathlib import Path
from typing import Optional

import torch

from . import qiddm_oracle as O

SRC = Path(__file__).resolve().parent / "statevec_oracle.c"
LIB = Path(__file__).resolve().parent / "_build" / "libqiddm_oracle_c.so"


class _Desc(ctypes.Structure):
    _fields_ = [("n_qubits", ctypes.c_int), ("n_blocks", ctypes.c_int), ("layers_per_block", ctypes.c_int),
                ("init", ctypes.c_int), ("n_features", ctypes.c_int), ("pad_value", ctypes.c_double),
                ("add_offset", ctypes.c_double), ("enc", ctypes.c_int), ("enc_scale", ctypes.c_double),
                ("imprimitive", ctypes.c_int), ("remap", ctypes.c_int), ("readout", ctypes.c_int),
                ("read_count", ctypes.c_int), ("read_stride", ctypes.c_int), ("post_scale", ctypes.c_double),
                ("clamp", ctypes.c_int), ("clamp_lo", ctypes.c_double), ("clamp_hi", ctypes.c_double)]


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: the C oracle cannot be built")
    LIB.parent.mkdir(exist_ok=True)
    subprocess.run([gcc, "-O2", "-std=gnu11", "-fopenmp", "-shared", "-fPIC", "-o", str(LIB), str(SRC), "-lm"], check=True)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
        _lib.qc_forward.restype = ctypes.c_int
        _lib.qc_forward.argtypes = [ctypes.POINTER(_Desc), ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_long]
        _lib.qc_set_threads.restype = None
        _lib.qc_set_threads.argtypes = [ctypes.c_int]
        _lib.qc_backward.restype = ctypes.c_int
        _lib.qc_backward.argtypes = [ctypes.POINTER(_Desc), ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    return _lib


def _c_desc(desc: O.StageDesc) -> _Desc:
    return _Desc(desc.n_qubits, desc.n_blocks, desc.layers_per_block, desc.init, desc.n_features, desc.pad_value,
                 desc.add_offset, desc.enc, desc.enc_scale, desc.imprimitive, desc.remap, desc.readout, desc.read_count,
                 desc.read_stride, desc.post_scale, int(bool(desc.clamp)), desc.clamp_lo, desc.clamp_hi)


def stage_grads(desc: O.StageDesc, x: Optional[torch.Tensor], weights: torch.Tensor, grad_out: torch.Tensor,
                batch: Optional[int] = None, threads: Optional[int] = None):
    """Adjoint-method gradients of sum(out * grad_out): (grad_weights like `weights`, grad_x like `x` or None)."""
    lib = _load()
    d = _c_desc(desc)
    w = weights.detach().to(torch.float64).reshape(desc.n_blocks, desc.layers_per_block, desc.n_qubits, 3).contiguous()
    xs = x.detach().to(torch.float64).contiguous() if x is not None else None
    go = grad_out.detach().to(torch.float64).contiguous()
    B = xs.shape[0] if xs is not None else int(batch)
    gw = torch.empty_like(w)
    gx = torch.zeros_like(xs) if xs is not None else None
    if threads is not None:
        lib.qc_set_threads(int(threads))
    rc = lib.qc_backward(ctypes.byref(d), xs.data_ptr() if xs is not None else None, xs.shape[1] if xs is not None else 0,
                         None, w.data_ptr(), go.data_ptr(), gx.data_ptr() if gx is not None else None, gw.data_ptr(), B)
    if rc != 0:
        raise ValueError(f"qc_backward: invalid descriptor (rc={rc})")
    return gw.reshape(weights.shape), gx


def run_stage(desc: O.StageDesc, x: Optional[torch.Tensor], weights: torch.Tensor, batch: Optional[int] = None,
              basis_index: Optional[torch.Tensor] = None, threads: Optional[int] = None) -> torch.Tensor:
    """(B, n_out) float64, the forward of `oracle.qiddm_oracle.run_stage` computed gate by gate in C (OpenMP over rows)."""
    lib = _load()
    d = _Desc(desc.n_qubits, desc.n_blocks, desc.layers_per_block, desc.init, desc.n_features, desc.pad_value,
              desc.add_offset, desc.enc, desc.enc_scale, desc.imprimitive, desc.remap, desc.readout, desc.read_count,
              desc.read_stride, desc.post_scale, int(bool(desc.clamp)), desc.clamp_lo, desc.clamp_hi)
    w = weights.detach().to(torch.float64).reshape(desc.n_blocks, desc.layers_per_block, desc.n_qubits, 3).contiguous()
    xs = x.detach().to(torch.float64).contiguous() if x is not None else None
    bi = basis_index.to(torch.int64).contiguous() if basis_index is not None else None
    B = xs.shape[0] if xs is not None else (bi.shape[0] if bi is not None else int(batch))
    out = torch.empty(B, desc.n_out, dtype=torch.float64)
    if threads is not None:
        lib.qc_set_threads(int(threads))
    rc = lib.qc_forward(ctypes.byref(d), xs.data_ptr() if xs is not None else None, xs.shape[1] if xs is not None else 0,
                        bi.data_ptr() if bi is not None else None, w.data_ptr(), out.data_ptr(), B)
    if rc != 0:
        raise ValueError(f"qc_forward: invalid descriptor (rc={rc})")
    return out
