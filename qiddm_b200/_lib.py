"""ctypes binding of libqiddm_b200.so (include/qiddm.h).  No torch types cross the ABI: only raw
device pointers, sizes and the CUDA stream handle.  There is NO CPU fallback: a missing library
or a non-CUDA tensor raises."""
from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import torch

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libqiddm_b200.so"

# enums (include/qiddm.h)
INIT_ZERO, INIT_AMPLITUDE, INIT_BASIS = 0, 1, 2
ENC_NONE, ENC_RZ, ENC_RY = 0, 1, 2
IMP_CNOT, IMP_CZ = 0, 1
REMAP_NONE, REMAP_TANH, REMAP_PI_TANH = 0, 1, 2
READ_PROBS, READ_EXPVAL_Z, READ_STATE = 0, 1, 2
DTYPE_F32, DTYPE_F64 = 0, 1
PATH_AUTO, PATH_GATE, PATH_GEMM = 0, 1, 2
MAX_QUBITS = 12


class CircuitDesc(C.Structure):
    _fields_ = [
        ("n_qubits", C.c_int32), ("n_blocks", C.c_int32), ("layers_per_block", C.c_int32),
        ("init", C.c_int32), ("n_features", C.c_int32), ("pad_value", C.c_float),
        ("add_offset", C.c_float), ("enc", C.c_int32), ("enc_scale", C.c_float),
        ("imprimitive", C.c_int32), ("remap", C.c_int32), ("readout", C.c_int32),
        ("read_count", C.c_int32), ("read_stride", C.c_int32), ("post_scale", C.c_float),
        ("clamp", C.c_int32), ("clamp_lo", C.c_float), ("clamp_hi", C.c_float), ("path", C.c_int32),
    ]


class UnfoldDesc(C.Structure):
    _fields_ = [
        ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("kernel_h", C.c_int32), ("kernel_w", C.c_int32), ("pad_h", C.c_int32), ("pad_w", C.c_int32),
    ]


EXPORTS = [
    "qiddm_abi_version", "qiddm_error_string", "qiddm_n_inputs", "qiddm_n_outputs", "qiddm_n_weights",
    "qiddm_plan_create", "qiddm_plan_destroy", "qiddm_workspace_bytes", "qiddm_forward", "qiddm_backward",
    "qiddm_qconv_forward", "qiddm_qconv_backward", "qiddm_build_unitary", "qiddm_launch_count",
    "qiddm_gemm_supported", "qiddm_gemm_collapsed_bytes", "qiddm_gemm_workspace_bytes", "qiddm_gemm_prepare",
    "qiddm_gemm_forward", "qiddm_gemm_backward", "qiddm_gemm_saved_bytes", "qiddm_timing_enable",
    "qiddm_timing_collect", "qiddm_qconv_gemm_saved_bytes", "qiddm_qconv_gemm_workspace_bytes",
    "qiddm_qconv_gemm_forward", "qiddm_qconv_gemm_backward", "qiddm_stream_capture_id", "qiddm_qconv_direct_supported", "qiddm_gemm_prepare_direct", "qiddm_qconv_up_forward", "qiddm_qconv_up_backward", "qiddm_mse_ladder_loss_grad", "qiddm_linear_up_mse_workspace_bytes", "qiddm_linear_up_mse_step",
    "qiddm_sym_eigh_max_dim", "qiddm_sym_eigh_f64", "qiddm_sym_eigh_f64_batched", "qiddm_upsample_bilinear_forward",
    "qiddm_upsample_bilinear_backward", "qiddm_batchnorm_workspace_bytes", "qiddm_batchnorm_forward",
    "qiddm_batchnorm_backward", "qiddm_noise_ladder", "qiddm_mse_workspace_bytes", "qiddm_mse_loss_grad",
    "qiddm_readout_channel", "qiddm_gemm_forward_workspace_bytes", "qiddm_probe_fp32_fma",
    "qiddm_qconv_reference_map_forward", "qiddm_qconv_reference_map_backward",
    "qiddm_state_bytes", "qiddm_forward_save", "qiddm_backward_saved",
    "qiddm_noisy_workspace_bytes", "qiddm_noisy_forward",
    "qiddm_batchnorm_relu_forward", "qiddm_batchnorm_relu_backward", "qiddm_maxpool2d_forward", "qiddm_maxpool2d_backward",
    "qiddm_qconv_forward_io", "qiddm_qconv_backward_io",
    "qiddm_dense_mse_step_workspace_bytes", "qiddm_dense_mse_step",
    "qiddm_skinny_linear_workspace_bytes", "qiddm_skinny_linear_forward", "qiddm_skinny_linear_backward",
]

_lib = None
_lock = threading.Lock()


class QiddmError(RuntimeError):
    pass


def load_library(path: Optional[Path] = None) -> C.CDLL:
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        p = Path(path) if path else LIB_PATH
        if not p.exists():
            raise QiddmError(
                f"{p} is missing: build it with `python -m qiddm_b200.build` (nvcc, sm_100a). "
                "qiddm_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(str(p))
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        lib.qiddm_abi_version.restype = i32
        lib.qiddm_error_string.restype = C.c_char_p
        lib.qiddm_error_string.argtypes = [i32]
        for f in (lib.qiddm_n_inputs, lib.qiddm_n_outputs, lib.qiddm_n_weights):
            f.restype = i32
            f.argtypes = [C.POINTER(CircuitDesc)]
        lib.qiddm_plan_create.restype = i32
        lib.qiddm_plan_create.argtypes = [C.POINTER(CircuitDesc), C.POINTER(vp)]
        lib.qiddm_plan_destroy.restype = None
        lib.qiddm_plan_destroy.argtypes = [vp]
        lib.qiddm_workspace_bytes.restype = C.c_size_t
        lib.qiddm_workspace_bytes.argtypes = [vp, i64]
        lib.qiddm_forward.restype = i32
        lib.qiddm_forward.argtypes = [vp, vp, vp, vp, i32, vp, vp, i64, vp]
        lib.qiddm_backward.restype = i32
        lib.qiddm_backward.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, i64, vp]
        lib.qiddm_state_bytes.restype = C.c_size_t
        lib.qiddm_state_bytes.argtypes = [vp, i64]
        lib.qiddm_forward_save.restype = i32
        lib.qiddm_forward_save.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i64, vp]
        lib.qiddm_backward_saved.restype = i32
        lib.qiddm_backward_saved.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i64, vp]
        lib.qiddm_noisy_workspace_bytes.restype = C.c_size_t
        lib.qiddm_noisy_workspace_bytes.argtypes = [vp, i64]
        lib.qiddm_noisy_forward.restype = i32
        lib.qiddm_noisy_forward.argtypes = [vp, vp, vp, i32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp,
                                            i64, vp]
        lib.qiddm_qconv_forward.restype = i32
        lib.qiddm_qconv_forward.argtypes = [vp, C.POINTER(UnfoldDesc), vp, vp, i32, vp, vp, i64, vp]
        lib.qiddm_qconv_backward.restype = i32
        lib.qiddm_qconv_backward.argtypes = [vp, C.POINTER(UnfoldDesc), vp, vp, i32, vp, vp, vp, vp, i64, vp]
        lib.qiddm_qconv_forward_io.restype = i32
        lib.qiddm_qconv_forward_io.argtypes = [vp, C.POINTER(UnfoldDesc), i32, vp, vp, i32, vp, vp, i64, vp]
        lib.qiddm_qconv_backward_io.restype = i32
        lib.qiddm_qconv_backward_io.argtypes = [vp, C.POINTER(UnfoldDesc), i32, vp, vp, i32, vp, vp, vp, vp, i64, vp]
        lib.qiddm_build_unitary.restype = i32
        lib.qiddm_build_unitary.argtypes = [vp, vp, i32, vp, vp, vp]
        lib.qiddm_launch_count.restype = i64
        f64 = C.c_double
        for f in (lib.qiddm_upsample_bilinear_forward, lib.qiddm_upsample_bilinear_backward):
            f.restype = i32
            f.argtypes = [vp, vp, i32, i64, i32, i32, i32, i32, f64, f64, vp]
        lib.qiddm_batchnorm_workspace_bytes.restype = C.c_size_t
        lib.qiddm_batchnorm_workspace_bytes.argtypes = [i32]
        lib.qiddm_batchnorm_forward.restype = i32
        lib.qiddm_batchnorm_forward.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, f64, f64, vp, vp]
        lib.qiddm_batchnorm_backward.restype = i32
        lib.qiddm_batchnorm_backward.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        lib.qiddm_batchnorm_relu_forward.restype = i32
        lib.qiddm_batchnorm_relu_forward.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, f64, f64, i32, vp, vp]
        lib.qiddm_batchnorm_relu_backward.restype = i32
        lib.qiddm_batchnorm_relu_backward.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp]
        lib.qiddm_maxpool2d_forward.restype = i32
        lib.qiddm_maxpool2d_forward.argtypes = [vp, vp, i32, i64, i32, i32, i32, vp]
        lib.qiddm_maxpool2d_backward.restype = i32
        lib.qiddm_maxpool2d_backward.argtypes = [vp, vp, vp, i32, i64, i32, i32, i32, vp]
        lib.qiddm_dense_mse_step_workspace_bytes.restype = C.c_size_t
        lib.qiddm_dense_mse_step_workspace_bytes.argtypes = [vp, i64, i32]
        lib.qiddm_dense_mse_step.restype = i32
        lib.qiddm_dense_mse_step.argtypes = [vp, vp, vp, vp, vp, i32, i64, i32, f64, f64, f64, f64, vp, i32, vp, vp, vp, i32, i32, vp]
        lib.qiddm_skinny_linear_workspace_bytes.restype = C.c_size_t
        lib.qiddm_skinny_linear_workspace_bytes.argtypes = [i64, i32, i32]
        lib.qiddm_skinny_linear_forward.restype = i32
        lib.qiddm_skinny_linear_forward.argtypes = [vp, vp, vp, vp, i32, i64, i32, i32, vp]
        lib.qiddm_skinny_linear_backward.restype = i32
        lib.qiddm_skinny_linear_backward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, vp, vp]
        lib.qiddm_noise_ladder.restype = i32
        lib.qiddm_noise_ladder.argtypes = [vp, vp, vp, i32, i64, i32, i32, vp, vp, vp]
        lib.qiddm_mse_workspace_bytes.restype = C.c_size_t
        lib.qiddm_mse_loss_grad.restype = i32
        lib.qiddm_mse_loss_grad.argtypes = [vp, vp, vp, i32, f64, f64, i64, vp, vp, vp, vp]
        lib.qiddm_linear_up_mse_workspace_bytes.restype = C.c_size_t
        lib.qiddm_linear_up_mse_workspace_bytes.argtypes = [i32, i32]
        lib.qiddm_linear_up_mse_step.restype = i32
        lib.qiddm_linear_up_mse_step.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, i32, i32, i32, f64, f64, f64, f64, vp, vp, vp, vp,
                                                 vp, vp]
        lib.qiddm_mse_ladder_loss_grad.restype = i32
        lib.qiddm_mse_ladder_loss_grad.argtypes = [vp, vp, vp, vp, i32, i64, i32, i32, f64, f64, f64, f64, vp, vp, vp, vp]
        lib.qiddm_readout_channel.restype = i32
        lib.qiddm_readout_channel.argtypes = [vp, vp, i32, i64, i32, f64, f64, f64, f64, vp]
        lib.qiddm_sym_eigh_max_dim.restype = i32
        lib.qiddm_sym_eigh_f64.restype = i32
        lib.qiddm_sym_eigh_f64.argtypes = [vp, i32, vp, vp, vp]
        lib.qiddm_sym_eigh_f64_batched.restype = i32
        lib.qiddm_sym_eigh_f64_batched.argtypes = [vp, i32, i64, vp, vp, vp]
        lib.qiddm_stream_capture_id.restype = i64
        lib.qiddm_stream_capture_id.argtypes = [vp]
        lib.qiddm_gemm_supported.restype = i32
        lib.qiddm_gemm_supported.argtypes = [vp]
        lib.qiddm_gemm_collapsed_bytes.restype = C.c_size_t
        lib.qiddm_gemm_collapsed_bytes.argtypes = [vp]
        lib.qiddm_gemm_workspace_bytes.restype = C.c_size_t
        lib.qiddm_gemm_workspace_bytes.argtypes = [vp, i64]
        lib.qiddm_gemm_prepare.restype = i32
        lib.qiddm_gemm_prepare.argtypes = [vp, vp, i32, vp, vp, vp]
        lib.qiddm_gemm_prepare_direct.restype = i32
        lib.qiddm_gemm_prepare_direct.argtypes = [vp, vp, i32, vp, vp, vp]
        lib.qiddm_gemm_forward.restype = i32
        lib.qiddm_gemm_forward.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, vp]
        lib.qiddm_gemm_backward.restype = i32
        lib.qiddm_gemm_backward.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i64, i32, vp]
        lib.qiddm_gemm_saved_bytes.restype = C.c_size_t
        lib.qiddm_gemm_saved_bytes.argtypes = [vp, i64]
        lib.qiddm_gemm_forward_workspace_bytes.restype = C.c_size_t
        lib.qiddm_gemm_forward_workspace_bytes.argtypes = [vp, i64]
        lib.qiddm_qconv_reference_map_forward.restype = i32
        lib.qiddm_qconv_reference_map_forward.argtypes = [C.POINTER(UnfoldDesc), i32, vp, vp, i32, i64, vp]
        lib.qiddm_qconv_reference_map_backward.restype = i32
        lib.qiddm_qconv_reference_map_backward.argtypes = [C.POINTER(UnfoldDesc), i32, vp, vp, vp, i32, i64, vp]
        lib.qiddm_probe_fp32_fma.restype = i32
        lib.qiddm_probe_fp32_fma.argtypes = [i32, vp, C.POINTER(C.c_double), vp]
        for f in (lib.qiddm_qconv_gemm_saved_bytes, lib.qiddm_qconv_gemm_workspace_bytes):
            f.restype = C.c_size_t
            f.argtypes = [vp, C.POINTER(UnfoldDesc), i64]
        lib.qiddm_qconv_up_forward.restype = i32
        lib.qiddm_qconv_up_forward.argtypes = [vp, vp, C.POINTER(UnfoldDesc), i32, vp, i32, i32, f64, f64, vp, vp, i64, vp]
        lib.qiddm_qconv_up_backward.restype = i32
        lib.qiddm_qconv_up_backward.argtypes = [vp, vp, C.POINTER(UnfoldDesc), i32, vp, i32, i32, f64, f64, vp, i32, vp, vp, vp, vp,
                                                vp, i64, vp]
        lib.qiddm_qconv_direct_supported.restype = i32
        lib.qiddm_qconv_direct_supported.argtypes = [vp, C.POINTER(UnfoldDesc)]
        lib.qiddm_qconv_gemm_forward.restype = i32
        lib.qiddm_qconv_gemm_forward.argtypes = [vp, vp, C.POINTER(UnfoldDesc), i32, vp, vp, vp, vp, i64, i32, vp]
        lib.qiddm_qconv_gemm_backward.restype = i32
        lib.qiddm_qconv_gemm_backward.argtypes = [vp, vp, C.POINTER(UnfoldDesc), i32, vp, vp, i32, vp, vp, vp, vp, vp,
                                                  i64, i32, vp]
        lib.qiddm_timing_enable.restype = None
        lib.qiddm_timing_enable.argtypes = [i32]
        lib.qiddm_timing_collect.restype = i32
        lib.qiddm_timing_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
        if lib.qiddm_abi_version() != 1:
            raise QiddmError("libqiddm_b200.so ABI version mismatch")
        _lib = lib
        return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load_library().qiddm_error_string(code).decode()
        raise QiddmError(f"{what} failed: {msg} (code {code})")


def launch_count() -> int:
    return int(load_library().qiddm_launch_count())


TIMING_KINDS = ("gate_forward", "gate_backward", "gemm", "other", "gemm_forward", "gemm_dx", "gemm_dw", "prep_x",
                "transpose_x", "g_bound", "grad_y", "finish_dx", "assemble", "build_w", "conv_forward", "conv_backward")


def timing_enable(on: bool) -> None:
    load_library().qiddm_timing_enable(1 if on else 0)


def timing_collect() -> dict:
    """{kind: {"ms": .., "work": .., "launches": ..}} since the last collect (synchronises)."""
    ms, wk, n = (C.c_double * 16)(), (C.c_double * 16)(), (C.c_int64 * 16)()
    check(load_library().qiddm_timing_collect(ms, wk, n), "qiddm_timing_collect")
    return {k: {"ms": ms[i], "work": wk[i], "launches": int(n[i])} for i, k in enumerate(TIMING_KINDS)}


def fp32_fma_peak_tflops(device=None, sustained: bool = False) -> float:
    """Measured FP32 FMA-pipe rate of this GPU (packed fma.rn.f32x2 chains, no memory traffic): the roofline denominator
    of the gate-by-gate kernels (MEASURED_PEAKS.json holds HBM and tensor peaks only).  Default = burst: best of five
    ~0.7 ms launches (the figure for a kernel timed alone; ~74 TFLOP/s = 148 SMs x 128 lanes x 2 x 1.965 GHz);
    sustained=True: one ~50 ms launch, which runs into the 1 kW power cap (~57 TFLOP/s)."""
    lib = load_library()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    sink = torch.zeros(1, dtype=torch.float32, device=dev)
    flops = C.c_double()
    best = 0.0
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(lib.qiddm_probe_fp32_fma(500, _ptr(sink), C.byref(flops), st), "qiddm_probe_fp32_fma")
        for _ in range(1 if sustained else 5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            check(lib.qiddm_probe_fp32_fma(150000 if sustained else 2000, _ptr(sink), C.byref(flops), st), "qiddm_probe_fp32_fma")
            b.record()
            b.synchronize()
            best = max(best, flops.value / (a.elapsed_time(b) * 1e-3) / 1e12)
    return best


def qconv_reference_map(img: torch.Tensor, unfold: UnfoldDesc, out_channels: int, grad_out: Optional[torch.Tensor] = None):
    """The forward nn/qconv.py:71-90 literally executes (no circuit) or, with `grad_out`, its backward."""
    _require_cuda(img, "input")
    lib = load_library()
    if img.dtype not in (torch.float32, torch.float64):
        raise QiddmError("qconv_reference_map needs float32 / float64 images")
    dt = DTYPE_F64 if img.dtype == torch.float64 else DTYPE_F32
    img = img.contiguous()
    n, c, h, w = img.shape
    ho, wo = h + 2 * unfold.pad_h - unfold.kernel_h + 1, w + 2 * unfold.pad_w - unfold.kernel_w + 1
    with torch.cuda.device(img.device):
        st = C.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)
        if grad_out is None:
            out = torch.empty((n, out_channels, ho, wo), dtype=img.dtype, device=img.device)
            check(lib.qiddm_qconv_reference_map_forward(C.byref(unfold), dt, _ptr(img), _ptr(out), out_channels, n, st),
                  "qiddm_qconv_reference_map_forward")
            return out
        go = grad_out.to(img.dtype).contiguous()
        gi = torch.empty_like(img)
        check(lib.qiddm_qconv_reference_map_backward(C.byref(unfold), dt, _ptr(img), _ptr(go), _ptr(gi), out_channels, n, st),
              "qiddm_qconv_reference_map_backward")
        return gi


def sym_eigh(a: torch.Tensor):
    """Eigen-decomposition of symmetric float64 CUDA matrices (..., m, m) by the library's Jacobi kernel (one CTA per
    matrix): (evals descending (..., m), evecs (..., m, m) with matching columns).  Asynchronous, CUDA-graph capturable."""
    _require_cuda(a, "matrix")
    lib = load_library()
    if a.dim() < 2 or a.shape[-1] != a.shape[-2] or a.dtype != torch.float64:
        raise QiddmError("sym_eigh needs square float64 matrices")
    m = a.shape[-1]
    if m > lib.qiddm_sym_eigh_max_dim():
        raise QiddmError(f"sym_eigh supports m <= {lib.qiddm_sym_eigh_max_dim()}, got {m}")
    a = a.contiguous()
    count = a.numel() // (m * m)
    evals = torch.empty(a.shape[:-1], dtype=torch.float64, device=a.device)
    evecs = torch.empty(a.shape, dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.qiddm_sym_eigh_f64_batched(_ptr(a), m, count, _ptr(evals), _ptr(evecs),
                                             C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)),
              "qiddm_sym_eigh_f64_batched")
    return evals, evecs


@dataclass(frozen=True)
class StageSpec:
    """Python mirror of qiddm_circuit_desc; hashable so plans can be cached per module."""
    n_qubits: int
    n_blocks: int = 1
    layers_per_block: int = 1
    init: int = INIT_ZERO
    n_features: int = 0
    pad_value: float = 0.0
    add_offset: float = 0.0
    enc: int = ENC_NONE
    enc_scale: float = 1.0
    imprimitive: int = IMP_CNOT
    remap: int = REMAP_NONE
    readout: int = READ_PROBS
    read_count: int = 0
    read_stride: int = 1
    post_scale: float = 1.0
    clamp: bool = False
    clamp_lo: float = 0.0
    clamp_hi: float = 1.0
    path: int = PATH_AUTO
    gemm_precision: int = 3   # host-side only: 3 = fp32-grade 3-term fp16 split, 1 = single fp16 pass
    gemm_bwd_precision: int = 0   # host-side only: precision of the dX / dW GEMMs; 0 = same as gemm_precision, 1 behind a
                                  # precision-3 forward = "x3 forward, x1 gradients" (stated gradient bound: DESIGN.md 4.2)

    def to_c(self) -> CircuitDesc:
        return CircuitDesc(self.n_qubits, self.n_blocks, self.layers_per_block, self.init, self.n_features,
                           self.pad_value, self.add_offset, self.enc, self.enc_scale, self.imprimitive,
                           self.remap, self.readout, self.read_count, self.read_stride, self.post_scale,
                           int(self.clamp), self.clamp_lo, self.clamp_hi, self.path)

    @property
    def bwd_precision(self) -> int:
        return self.gemm_bwd_precision or self.gemm_precision

    @property
    def dim(self) -> int:
        return 1 << self.n_qubits

    @property
    def n_in(self) -> int:
        if self.init == INIT_AMPLITUDE:
            return self.n_features
        return self.n_qubits if self.enc != ENC_NONE else 0

    @property
    def n_out(self) -> int:
        if self.readout == READ_PROBS:
            return self.read_count
        return self.n_qubits if self.readout == READ_EXPVAL_Z else 2 * self.dim

    @property
    def n_weights(self) -> int:
        return self.n_blocks * self.layers_per_block * self.n_qubits * 3


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _wdtype(w: torch.Tensor) -> int:
    if w.dtype == torch.float32:
        return DTYPE_F32
    if w.dtype == torch.float64:
        return DTYPE_F64
    raise QiddmError(f"weights must be float32 or float64, got {w.dtype}")


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise QiddmError(f"{name} must be a CUDA tensor: qiddm_b200 runs on sm_100a only (no CPU fallback)")


class Plan:
    """Owns a qiddm_plan*.  Thread-safe for concurrent forward/backward (each call allocates its own
    workspace from torch's caching allocator)."""

    _cache: dict = {}
    _cache_lock = threading.Lock()

    def __init__(self, spec: StageSpec):
        self.spec = spec
        self.lib = load_library()
        handle = C.c_void_p()
        desc = spec.to_c()
        check(self.lib.qiddm_plan_create(C.byref(desc), C.byref(handle)), "qiddm_plan_create")
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.qiddm_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @classmethod
    def get(cls, spec: StageSpec) -> "Plan":
        with cls._cache_lock:
            p = cls._cache.get(spec)
            if p is None:
                p = cls._cache[spec] = Plan(spec)
            return p

    # ------------------------------------------------------------------ helpers
    def _workspace(self, batch: int, device) -> torch.Tensor:
        with torch.cuda.device(device):
            nbytes = int(self.lib.qiddm_workspace_bytes(self.handle, batch))
        return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)

    @staticmethod
    def _stream(device):
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)

    def _check_weights(self, w: torch.Tensor) -> torch.Tensor:
        _require_cuda(w, "weights")
        if w.numel() != self.spec.n_weights:
            raise QiddmError(f"weights has {w.numel()} elements, circuit needs {self.spec.n_weights}")
        return w.contiguous()

    # ------------------------------------------------------------------ dense rows
    # psi_final kept between forward and backward when it is at most this many bytes (QIDDM_SAVE_STATE_MB, default 4096;
    # 0 disables): the adjoint then skips its forward recomputation
    SAVE_STATE_BYTES = int(float(os.environ.get("QIDDM_SAVE_STATE_MB", "4096")) * (1 << 20))

    def state_bytes(self, batch: int) -> int:
        return int(self.lib.qiddm_state_bytes(self.handle, batch))

    def forward(self, x: Optional[torch.Tensor], weights: torch.Tensor, batch: Optional[int] = None,
                basis: Optional[torch.Tensor] = None, save_state: bool = False):
        """(B, n_out) outputs; with `save_state` also the (B, 2^n, 2) final states for `backward(state=...)` (None when they
        would exceed SAVE_STATE_BYTES)."""
        w = self._check_weights(weights)
        dev = w.device
        if self.spec.n_in > 0:
            _require_cuda(x, "input")
            if x.dim() != 2 or x.shape[1] != self.spec.n_in:
                raise QiddmError(f"input must be (B, {self.spec.n_in}), got {tuple(x.shape)}")
            x = x.to(torch.float32).contiguous()
            batch = x.shape[0]
        elif batch is None:
            raise QiddmError("batch is required for circuits without inputs")
        out = torch.empty((batch, self.spec.n_out), dtype=torch.float32, device=dev)
        ws = self._workspace(batch, dev)
        state = None
        if save_state and 0 < self.state_bytes(batch) <= self.SAVE_STATE_BYTES:
            state = torch.empty((batch, self.spec.dim, 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.qiddm_forward_save(self.handle, _ptr(x) if self.spec.n_in else None, _ptr(basis), _ptr(w),
                                              _wdtype(w), _ptr(out), _ptr(state), _ptr(ws), batch, self._stream(dev)),
                  "qiddm_forward")
        return (out, state) if save_state else out

    def backward(self, x: Optional[torch.Tensor], weights: torch.Tensor, grad_out: torch.Tensor,
                 need_grad_in: bool = True, need_grad_w: bool = True, basis: Optional[torch.Tensor] = None,
                 state: Optional[torch.Tensor] = None):
        w = self._check_weights(weights)
        dev = w.device
        go = grad_out.to(torch.float32).contiguous()
        batch = go.shape[0]
        if self.spec.n_in > 0:
            x = x.to(torch.float32).contiguous()
        grad_in = (torch.empty((batch, self.spec.n_in), dtype=torch.float32, device=dev)
                   if (need_grad_in and self.spec.n_in > 0) else None)
        grad_w = torch.empty_like(w) if need_grad_w else None
        ws = self._workspace(batch, dev)
        with torch.cuda.device(dev):
            check(self.lib.qiddm_backward_saved(self.handle, _ptr(x) if self.spec.n_in else None, _ptr(basis), _ptr(w),
                                                _wdtype(w), _ptr(go), _ptr(state), _ptr(grad_in), _ptr(grad_w), _ptr(ws),
                                                batch, self._stream(dev)), "qiddm_backward")
        return grad_in, grad_w

    # ------------------------------------------------------------------ mid-circuit noise channels (density matrix)
    NOISY_MAX_BYTES = 64 << 30

    def noisy_forward(self, x: torch.Tensor, weights: torch.Tensor, f_off: float, m) -> torch.Tensor:
        """Re-upload stage with a single-qubit channel after every RZ(a_j) (qiddm_noisy_forward): inference only."""
        w = self._check_weights(weights)
        _require_cuda(x, "input")
        dev = w.device
        if x.dim() != 2 or x.shape[1] != self.spec.n_in:
            raise QiddmError(f"input must be (B, {self.spec.n_in}), got {tuple(x.shape)}")
        x = x.to(torch.float32).contiguous()
        batch = x.shape[0]
        out = torch.empty((batch, self.spec.n_out), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nbytes = int(self.lib.qiddm_noisy_workspace_bytes(self.handle, batch))
            if nbytes == 0:
                raise QiddmError("this circuit has no mid-circuit-noise path (needs |0..0> start, RZ re-upload, probs / <Z> readout)")
            if nbytes > self.NOISY_MAX_BYTES:
                raise QiddmError(f"density-matrix workspace of {nbytes >> 20} MiB: lower the batch ({batch} x 4^{self.spec.n_qubits} amplitudes)")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            check(self.lib.qiddm_noisy_forward(self.handle, _ptr(x), _ptr(w), _wdtype(w), float(f_off), *[float(v) for v in m],
                                               _ptr(out), _ptr(ws), batch, self._stream(dev)), "qiddm_noisy_forward")
        return out

    # ------------------------------------------------------------------ fused-unfold QConv
    def qconv_forward(self, img: torch.Tensor, weights: torch.Tensor, unfold: UnfoldDesc) -> torch.Tensor:
        w = self._check_weights(weights)
        _require_cuda(img, "input")
        dev = w.device
        # float64 images (the reference's UNet) are read and written in place of a cast; the simulation is fp32
        io = torch.float64 if img.dtype == torch.float64 else torch.float32
        img = img.to(io).contiguous()
        n, c, h, wd = img.shape
        ho = h + 2 * unfold.pad_h - unfold.kernel_h + 1
        wo = wd + 2 * unfold.pad_w - unfold.kernel_w + 1
        out = torch.empty((n, self.spec.read_count, ho, wo), dtype=io, device=dev)
        ws = self._workspace(n * ho * wo, dev)
        with torch.cuda.device(dev):
            check(self.lib.qiddm_qconv_forward_io(self.handle, C.byref(unfold), DTYPE_F64 if io == torch.float64 else DTYPE_F32,
                                                  _ptr(img), _ptr(w), _wdtype(w), _ptr(out), _ptr(ws), n, self._stream(dev)),
                  "qiddm_qconv_forward")
        return out

    def qconv_backward(self, img: torch.Tensor, weights: torch.Tensor, grad_out: torch.Tensor, unfold: UnfoldDesc,
                       need_grad_in: bool = True, need_grad_w: bool = True):
        w = self._check_weights(weights)
        dev = w.device
        io = torch.float64 if img.dtype == torch.float64 else torch.float32
        img = img.to(io).contiguous()
        go = grad_out.to(io).contiguous()
        n = img.shape[0]
        grad_img = torch.empty_like(img) if need_grad_in else None
        grad_w = torch.empty_like(w) if need_grad_w else None
        ws = self._workspace(n * go.shape[2] * go.shape[3], dev)
        with torch.cuda.device(dev):
            check(self.lib.qiddm_qconv_backward_io(self.handle, C.byref(unfold), DTYPE_F64 if io == torch.float64 else DTYPE_F32,
                                                   _ptr(img), _ptr(w), _wdtype(w), _ptr(go), _ptr(grad_img), _ptr(grad_w),
                                                   _ptr(ws), n, self._stream(dev)), "qiddm_qconv_backward")
        return grad_img, grad_w

    # ------------------------------------------------------------------ unitary-collapse (GEMM) path
    def gemm_supported(self) -> bool:
        return bool(self.lib.qiddm_gemm_supported(self.handle))

    def use_gemm(self, batch: int) -> bool:
        """PATH_AUTO rule, a cost model fitted to the B200 sweep (profiles/r2_configs_sweep.jsonl, gate path with the
        psi_final hand-over): per instance the gate path costs ~70 flop per Rot and amplitude (forward + adjoint) at
        ~28 TFLOP/s effective; the collapse path costs three fp16x3 GEMMs at ~1.5 PFLOP/s executed plus its streaming
        operand traffic, and per optimizer step the collapse itself (gate kernels on the 2^n basis columns, forward +
        adjoint) plus ~0.2 ms of small launches.  Against the 144 sweep points with both paths measured it picks the
        slower one once, by 1 % (tests/test_host_logic.py::test_path_dispatch_cost_model_against_the_measured_sweep)."""
        if self.spec.path == PATH_GATE or not self.gemm_supported():
            return False
        if self.spec.path == PATH_GEMM:
            return True
        a, f, n_out = self.spec.dim, self.spec.n_features, self.spec.read_count
        if batch < 2 * a:
            return False
        n_rot = self.spec.n_blocks * self.spec.layers_per_block * self.spec.n_qubits
        kp = (f + 1 + 7) // 8 * 8
        n_layers = self.spec.n_blocks * self.spec.layers_per_block
        t_gate = 70.0 * (n_rot + 11) * a / 28e12              # + embedding / readout ~ 11 Rot-equivalents
        t_gemm = 18.0 * kp * (2 * n_out) / 1.5e15 + (24.0 * f + 20.0 * 2 * n_out) / 5e12 + 1e-9
        # the collapse runs 2^n basis columns: throughput-bound for wide states, latency-bound (serial layers) otherwise
        t_collapse = max(a * t_gate, n_layers * 3.5e-6) + 2.0e-4
        return batch * t_gemm + t_collapse < batch * t_gate

    COLLAPSED_CACHE_ENTRIES = 8      # weight tensors per plan (layers that share a StageSpec share the plan, not U)

    def invalidate(self) -> None:
        """Drop every cached collapsed operator of this plan.  Needed after writes that do not bump the weights'
        version counter (`p.data.copy_`, `dist.broadcast(p.data)`, ...): write through `p` under torch.no_grad(), call
        torch.autograd.graph.increment_version(p), or call this."""
        with self._cache_lock:
            self._collapsed = {}
            self._collapsed_capture = {}

    @classmethod
    def invalidate_all(cls) -> None:
        with cls._cache_lock:
            plans = list(cls._cache.values())
        for p in plans:
            p.invalidate()

    def gemm_prepare(self, weights: torch.Tensor, direct: bool = False) -> torch.Tensor:
        """Collapsed operator (U^T + fp16 GEMM operands; `direct`: U^T + the fp32 filter rows of the direct QConv path only)
        for the current weights; cached per weight tensor and version (modules with equal StageSpecs share the Plan but not
        the operator: a UNet's same-shaped QConv layers keep one entry each)."""
        w = self._check_weights(weights)
        # identity of the (base) tensor object + its version counter; the cache keeps a strong reference
        # to that object, so its address cannot be recycled by another tensor while the entry lives
        base = weights._base if weights._base is not None else weights
        key = (weights.storage_offset(), weights.numel(), weights._version, w.data_ptr())
        # under CUDA-graph capture the collapse must be part of the graph (replays do not bump `_version`), and the
        # buffer belongs to the graph's pool: neither read nor update the eager cache
        dev = w.device
        capture = int(self.lib.qiddm_stream_capture_id(self._stream(dev)))
        slot = (id(base), weights.storage_offset(), weights.numel(), bool(direct))
        with self._cache_lock:
            if not isinstance(getattr(self, "_collapsed", None), dict):
                self._collapsed, self._collapsed_capture = {}, {}
            cache = self._collapsed_capture if capture else self._collapsed
            if capture:
                key = key + (capture,)
            cached = cache.get(slot)
            if cached is not None and cached[0] is base and cached[1] == key:
                cache[slot] = cache.pop(slot)          # most recently used last
                return cached[2]
        # zero-filled: the fp16 operand rows are padded to 16 bytes and the padding meets zero columns of the other operand
        # (the direct form writes every byte it later reads: no fill)
        buf = (torch.empty if direct else torch.zeros)(int(self.lib.qiddm_gemm_collapsed_bytes(self.handle)), dtype=torch.uint8, device=dev)
        ws = self._workspace(self.spec.dim, dev)
        with torch.cuda.device(dev):
            prep = self.lib.qiddm_gemm_prepare_direct if direct else self.lib.qiddm_gemm_prepare
            check(prep(self.handle, _ptr(w), _wdtype(w), _ptr(buf), _ptr(ws), self._stream(dev)), "qiddm_gemm_prepare")
        with self._cache_lock:
            cache.pop(slot, None)
            cache[slot] = (base, key, buf)
            while len(cache) > self.COLLAPSED_CACHE_ENTRIES:
                cache.pop(next(iter(cache)))
        Plan.collapse_count += 1
        return buf

    collapse_count = 0       # qiddm_gemm_prepare calls issued by this process (tests: one collapse per layer and step)

    def _gemm_ws(self, batch, dev):
        with torch.cuda.device(dev):
            nbytes = int(self.lib.qiddm_gemm_workspace_bytes(self.handle, batch))
        return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)

    def gemm_forward(self, x: torch.Tensor, weights: torch.Tensor, save: bool = False):
        """Returns out, or (out, saved) when `save` (training): `saved` holds the operand splits and Y."""
        _require_cuda(x, "input")
        col = self.gemm_prepare(weights)
        dev = col.device
        x = x.to(torch.float32).contiguous()
        batch = x.shape[0]
        out = torch.empty((batch, self.spec.n_out), dtype=torch.float32, device=dev)
        saved = None
        with torch.cuda.device(dev):
            if save:
                saved = torch.empty(int(self.lib.qiddm_gemm_saved_bytes(self.handle, batch)), dtype=torch.uint8,
                                    device=dev)
                ws = torch.empty(256, dtype=torch.uint8, device=dev)
            else:
                ws = torch.empty(int(self.lib.qiddm_gemm_forward_workspace_bytes(self.handle, batch)), dtype=torch.uint8,
                                 device=dev)
            check(self.lib.qiddm_gemm_forward(self.handle, _ptr(col), _ptr(x), _ptr(out), _ptr(saved), _ptr(ws),
                                              batch, self.spec.gemm_precision, self._stream(dev)),
                  "qiddm_gemm_forward")
        return (out, saved) if save else out

    def gemm_backward(self, x: torch.Tensor, weights: torch.Tensor, grad_out: torch.Tensor,
                      need_grad_in: bool = True, need_grad_w: bool = True, saved: Optional[torch.Tensor] = None):
        w = self._check_weights(weights)
        col = self.gemm_prepare(weights)
        dev = col.device
        x = x.to(torch.float32).contiguous()
        go = grad_out.to(torch.float32).contiguous()
        batch = x.shape[0]
        grad_in = torch.empty((batch, self.spec.n_in), dtype=torch.float32, device=dev) if need_grad_in else None
        grad_w = torch.empty_like(w) if need_grad_w else None
        ws = self._gemm_ws(batch, dev)
        with torch.cuda.device(dev):
            check(self.lib.qiddm_gemm_backward(self.handle, _ptr(col), _ptr(x), _ptr(w), _wdtype(w), _ptr(go),
                                               _ptr(saved), _ptr(grad_in), _ptr(grad_w), _ptr(ws), batch,
                                               self.spec.bwd_precision, self._stream(dev)), "qiddm_gemm_backward")
        return grad_in, grad_w

    def dense_mse_step(self, x: torch.Tensor, eps: torch.Tensor, level_w: torch.Tensor, T: int, weights: torch.Tensor,
                       a: float = 1.0, b: float = 0.0, c0: float = 1.0, c1: float = 0.0):
        """Fused diffusion training step of a single amplitude-embedding layer (qiddm_dense_mse_step): images x
        (n, pixels) float32 / float64, noise draw eps (n, pixels) float32, level weights (T + 1) -> (loss 0-d tensor in
        x.dtype, d loss / d weights).  Rows (b, t): in = level_{t+1}, target = c0 level_t + c1 level_{t+1}, d = a out + b - target."""
        _require_cuda(x, "images")
        w = self._check_weights(weights)
        col = self.gemm_prepare(weights)
        dev = col.device
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float32)
        x = x.contiguous()
        eps = eps.to(device=dev, dtype=torch.float32).contiguous()
        level_w = level_w.to(device=dev, dtype=x.dtype).contiguous()
        if x.dim() != 2 or eps.shape != x.shape or level_w.numel() != T + 1 or x.shape[1] != self.spec.n_in:
            raise QiddmError(f"dense_mse_step: images {tuple(x.shape)}, eps {tuple(eps.shape)}, {level_w.numel()} level weights for T = {T}")
        n = x.shape[0]
        loss = torch.empty((), dtype=x.dtype, device=dev)
        grad_w = torch.empty_like(w)
        with torch.cuda.device(dev):
            ws = torch.empty(max(int(self.lib.qiddm_dense_mse_step_workspace_bytes(self.handle, n, T)), 256), dtype=torch.uint8, device=dev)
            check(self.lib.qiddm_dense_mse_step(self.handle, _ptr(col), _ptr(x), _ptr(eps), _ptr(level_w),
                                                DTYPE_F64 if x.dtype == torch.float64 else DTYPE_F32, n, T, float(a), float(b),
                                                float(c0), float(c1), _ptr(w), _wdtype(w), _ptr(loss), _ptr(grad_w), _ptr(ws),
                                                self.spec.gemm_precision, self.spec.bwd_precision, self._stream(dev)),
                  "qiddm_dense_mse_step")
        return loss, grad_w

    # ------------------------------------------------------------------ QConv on the unitary-collapse path
    def qconv_direct(self, unfold: UnfoldDesc) -> bool:
        """True when the layer has a direct fp32 convolution behind the qconv_gemm entry points (csrc/qiddm_conv.cu)."""
        key = (unfold.channels, unfold.height, unfold.width, unfold.kernel_h, unfold.kernel_w, unfold.pad_h, unfold.pad_w)
        cache = self.__dict__.setdefault("_direct", {})
        v = cache.get(key)
        if v is None:
            v = cache[key] = bool(self.lib.qiddm_qconv_direct_supported(self.handle, C.byref(unfold)))
        return v

    def use_collapse_qconv(self, unfold: UnfoldDesc, patches: int) -> bool:
        """QConv dispatch: the collapse path (direct convolution where the shape has one, else the tcgen05 GEMM) or gate by
        gate.  The direct form costs F * N FMAs per patch and pass plus the collapse on 2^n basis columns: it wins as soon as
        there are at least as many patches as basis columns."""
        if self.spec.path == PATH_GATE or not self.gemm_supported():
            return False
        if self.spec.path == PATH_AUTO and self.qconv_direct(unfold):
            return patches >= self.spec.dim
        return self.use_gemm(patches)

    @staticmethod
    def _out_hw(img: torch.Tensor, unfold: UnfoldDesc):
        return (img.shape[2] + 2 * unfold.pad_h - unfold.kernel_h + 1,
                img.shape[3] + 2 * unfold.pad_w - unfold.kernel_w + 1)

    def qconv_gemm_forward(self, img: torch.Tensor, weights: torch.Tensor, unfold: UnfoldDesc, save: bool = False):
        """(N,C,H,W) -> (N,read_count,H_out,W_out); with `save` also returns the operand splits + Y for the backward."""
        _require_cuda(img, "input")
        col = self.gemm_prepare(weights, direct=self.qconv_direct(unfold))
        dev = col.device
        # float64 images (the reference's UNet) are read and written in place of a cast; the simulation is fp32
        io = torch.float64 if img.dtype == torch.float64 else torch.float32
        img = img.to(io).contiguous()
        n = img.shape[0]
        ho, wo = self._out_hw(img, unfold)
        out = torch.empty((n, self.spec.read_count, ho, wo), dtype=io, device=dev)
        saved = None
        with torch.cuda.device(dev):
            if save:
                saved = torch.empty(int(self.lib.qiddm_qconv_gemm_saved_bytes(self.handle, C.byref(unfold), n)),
                                    dtype=torch.uint8, device=dev)
                ws = torch.empty(256, dtype=torch.uint8, device=dev)
            else:
                ws = torch.empty(int(self.lib.qiddm_qconv_gemm_saved_bytes(self.handle, C.byref(unfold), n)),
                                 dtype=torch.uint8, device=dev)
            check(self.lib.qiddm_qconv_gemm_forward(self.handle, _ptr(col), C.byref(unfold),
                                                    DTYPE_F64 if io == torch.float64 else DTYPE_F32, _ptr(img), _ptr(out),
                                                    _ptr(saved), _ptr(ws), n, self.spec.gemm_precision,
                                                    self._stream(dev)), "qiddm_qconv_gemm_forward")
        return (out, saved) if save else out

    def qconv_gemm_backward(self, img: torch.Tensor, weights: torch.Tensor, grad_out: torch.Tensor, unfold: UnfoldDesc,
                            need_grad_in: bool = True, need_grad_w: bool = True,
                            saved: Optional[torch.Tensor] = None):
        w = self._check_weights(weights)
        col = self.gemm_prepare(weights, direct=self.qconv_direct(unfold))
        dev = col.device
        io = torch.float64 if img.dtype == torch.float64 else torch.float32
        img = img.to(io).contiguous()
        go = grad_out.to(io).contiguous()
        n = img.shape[0]
        grad_img = torch.empty_like(img) if need_grad_in else None
        grad_w = torch.empty_like(w) if need_grad_w else None
        with torch.cuda.device(dev):
            ws = torch.empty(int(self.lib.qiddm_qconv_gemm_workspace_bytes(self.handle, C.byref(unfold), n)),
                             dtype=torch.uint8, device=dev)
            check(self.lib.qiddm_qconv_gemm_backward(self.handle, _ptr(col), C.byref(unfold),
                                                     DTYPE_F64 if io == torch.float64 else DTYPE_F32, _ptr(img), _ptr(w),
                                                     _wdtype(w), _ptr(go), _ptr(saved), _ptr(grad_img), _ptr(grad_w),
                                                     _ptr(ws), n, self.spec.bwd_precision, self._stream(dev)),
                  "qiddm_qconv_gemm_backward")
        return grad_img, grad_w

    # ------------------------------------------------------------------ bilinear Upsample -> 1 x 1 QConv in one pass
    def qconv_up_forward(self, src: torch.Tensor, weights: torch.Tensor, unfold: UnfoldDesc, scale_h: float, scale_w: float,
                         save: bool = False):
        """(N, C, h, w) source of the upsample -> (N, read_count, H, W) of the 1 x 1 QConv on the (H, W) = unfold geometry."""
        _require_cuda(src, "input")
        col = self.gemm_prepare(weights, direct=True)
        dev = col.device
        io = torch.float64 if src.dtype == torch.float64 else torch.float32
        src = src.to(io).contiguous()
        n = src.shape[0]
        out = torch.empty((n, self.spec.read_count, unfold.height, unfold.width), dtype=io, device=dev)
        saved = None
        with torch.cuda.device(dev):
            if save:
                saved = torch.empty(int(self.lib.qiddm_qconv_gemm_saved_bytes(self.handle, C.byref(unfold), n)),
                                    dtype=torch.uint8, device=dev)
            check(self.lib.qiddm_qconv_up_forward(self.handle, _ptr(col), C.byref(unfold), DTYPE_F64 if io == torch.float64 else DTYPE_F32,
                                                  _ptr(src), src.shape[2], src.shape[3], float(scale_h), float(scale_w), _ptr(out),
                                                  _ptr(saved), n, self._stream(dev)), "qiddm_qconv_up_forward")
        return (out, saved) if save else out

    def qconv_up_backward(self, src: torch.Tensor, weights: torch.Tensor, grad_out: torch.Tensor, unfold: UnfoldDesc,
                          scale_h: float, scale_w: float, saved: torch.Tensor, need_grad_in: bool = True, need_grad_w: bool = True):
        """Returns (gradient w.r.t. the UPSAMPLED image or None, gradient w.r.t. the circuit weights or None)."""
        w = self._check_weights(weights)
        col = self.gemm_prepare(weights, direct=True)
        dev = col.device
        io = torch.float64 if src.dtype == torch.float64 else torch.float32
        src = src.to(io).contiguous()
        go = grad_out.to(io).contiguous()
        n = src.shape[0]
        grad_up = torch.empty((n, src.shape[1], unfold.height, unfold.width), dtype=io, device=dev) if need_grad_in else None
        grad_w = torch.empty_like(w) if need_grad_w else None
        with torch.cuda.device(dev):
            ws = torch.empty(int(self.lib.qiddm_qconv_gemm_workspace_bytes(self.handle, C.byref(unfold), n)),
                             dtype=torch.uint8, device=dev)
            check(self.lib.qiddm_qconv_up_backward(self.handle, _ptr(col), C.byref(unfold), DTYPE_F64 if io == torch.float64 else DTYPE_F32,
                                                   _ptr(src), src.shape[2], src.shape[3], float(scale_h), float(scale_w), _ptr(w),
                                                   _wdtype(w), _ptr(go), _ptr(saved), _ptr(grad_up), _ptr(grad_w), _ptr(ws), n,
                                                   self._stream(dev)), "qiddm_qconv_up_backward")
        return grad_up, grad_w

    def build_unitary(self, weights: torch.Tensor) -> torch.Tensor:
        """Returns U as a (2^n, 2^n) complex64 tensor (the library writes U^T, row c = U|c>)."""
        w = self._check_weights(weights)
        dev = w.device
        a = self.spec.dim
        ut = torch.empty((a, a, 2), dtype=torch.float32, device=dev)
        ws = self._workspace(a, dev)
        with torch.cuda.device(dev):
            check(self.lib.qiddm_build_unitary(self.handle, _ptr(w), _wdtype(w), _ptr(ut), _ptr(ws),
                                               self._stream(dev)), "qiddm_build_unitary")
        return torch.view_as_complex(ut).transpose(0, 1)
