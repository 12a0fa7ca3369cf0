"""The C-ABI library loads without a GPU and exports every symbol include/qiddm.h declares; descriptor
validation and error behaviour work host-side (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from qiddm_b200 import build
    build.build()
    from qiddm_b200._lib import load_library
    return load_library()


def declared_symbols():
    text = (ROOT / "include" / "qiddm.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qiddm_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/qiddm.h but not exported"
    from qiddm_b200._lib import EXPORTS
    assert sorted(EXPORTS) == syms


def test_abi_version_and_error_strings(lib):
    assert lib.qiddm_abi_version() == 1
    assert lib.qiddm_error_string(0) == b"ok"
    assert b"invalid" in lib.qiddm_error_string(-1)


def test_descriptor_io_sizes(lib):
    from qiddm_b200 import _lib as L
    s = L.StageSpec(n_qubits=10, layers_per_block=60, init=L.INIT_AMPLITUDE, n_features=784, pad_value=0.1,
                    readout=L.READ_PROBS, read_count=784, post_scale=784.0, clamp=True)
    d = s.to_c()
    assert lib.qiddm_n_inputs(C.byref(d)) == 784 == s.n_in
    assert lib.qiddm_n_outputs(C.byref(d)) == 784 == s.n_out
    assert lib.qiddm_n_weights(C.byref(d)) == 1800 == s.n_weights
    s2 = L.StageSpec(n_qubits=6, n_blocks=14, layers_per_block=2, enc=L.ENC_RZ, imprimitive=L.IMP_CZ,
                     readout=L.READ_EXPVAL_Z)
    d2 = s2.to_c()
    assert (lib.qiddm_n_inputs(C.byref(d2)), lib.qiddm_n_outputs(C.byref(d2)), lib.qiddm_n_weights(C.byref(d2))) == (6, 6, 504)


@pytest.mark.parametrize("bad", [dict(n_qubits=0), dict(n_qubits=13), dict(n_qubits=4, init=1, n_features=17),
                                 dict(n_qubits=4, readout=0, read_count=9, read_stride=2),
                                 dict(n_qubits=4, readout=1, clamp=True), dict(n_qubits=4, layers_per_block=0)])
def test_invalid_descriptors_are_rejected(lib, bad):
    from qiddm_b200 import _lib as L
    kw = dict(n_qubits=4, readout=L.READ_PROBS, read_count=4)
    kw.update(bad)
    d = L.StageSpec(**kw).to_c()
    h = C.c_void_p()
    assert lib.qiddm_plan_create(C.byref(d), C.byref(h)) == -1
    assert not h.value
    assert lib.qiddm_n_outputs(C.byref(d)) == 0


def test_plan_lifecycle_and_unsupported_depth(lib):
    from qiddm_b200 import _lib as L
    p = L.Plan(L.StageSpec(n_qubits=3, readout=L.READ_EXPVAL_Z))
    assert p.handle.value
    del p
    with pytest.raises(L.QiddmError):
        L.Plan(L.StageSpec(n_qubits=12, layers_per_block=800, readout=L.READ_EXPVAL_Z))


def test_product_path_has_no_cpu_fallback():
    import torch
    from qiddm_b200 import _lib as L, nn
    m = nn.QIDDM_LL_noise(64, 4, 2, 1)
    with pytest.raises(L.QiddmError):
        m.cpu()(torch.rand(2, 1, 8, 8))
    src = "".join(p.read_text() for p in (ROOT / "qiddm_b200").rglob("*.py"))
    assert "oracle" not in src.replace("oracle/", "")  # the product never imports the test oracle


def test_host_side_argument_checks_of_the_new_entry_points(lib):
    """No compute: NULL / invalid arguments are rejected before any launch (no GPU needed)."""
    from qiddm_b200 import _lib as L
    assert lib.qiddm_sym_eigh_max_dim() >= 100
    assert lib.qiddm_sym_eigh_f64(None, 4, None, None, None) == -1
    assert lib.qiddm_batchnorm_workspace_bytes(0) == 0 and lib.qiddm_batchnorm_workspace_bytes(8) > 0
    assert lib.qiddm_batchnorm_forward(None, None, 1, 1, 1, 1, None, None, None, None, None, None, 0.1, 1e-5, None, None) == -1
    assert lib.qiddm_batchnorm_backward(None, None, None, 1, 1, 1, 1, None, None, None, None, None, None, None) == -1
    assert lib.qiddm_upsample_bilinear_forward(None, None, 1, 1, 2, 2, 4, 4, 0.5, 0.5, None) == -1
    assert lib.qiddm_stream_capture_id(None) == 0 or True      # no context on a CPU box: must not crash
    # QConv collapse path: a descriptor / unfold geometry mismatch yields 0 bytes / EINVAL
    s = L.StageSpec(n_qubits=7, layers_per_block=3, init=L.INIT_AMPLITUDE, n_features=72, pad_value=0.5, add_offset=0.1,
                    readout=L.READ_PROBS, read_count=8, read_stride=2, post_scale=64.0, clamp=True)
    plan = L.Plan(s)
    good = L.UnfoldDesc(8, 28, 28, 3, 3, 1, 1)
    bad = L.UnfoldDesc(4, 28, 28, 3, 3, 1, 1)           # 4*9 != 72 features
    assert lib.qiddm_qconv_gemm_saved_bytes(plan.handle, C.byref(good), 10) > 0
    assert lib.qiddm_qconv_gemm_saved_bytes(plan.handle, C.byref(bad), 10) == 0
    # this layer (3x3 "same" window, 8 output channels) has the direct fp32 convolution behind the same entry points ...
    assert lib.qiddm_qconv_direct_supported(plan.handle, C.byref(good)) == 1
    assert lib.qiddm_qconv_direct_supported(plan.handle, C.byref(bad)) == 0
    assert lib.qiddm_qconv_gemm_workspace_bytes(plan.handle, C.byref(good), 10) > 0
    # ... unless the tcgen05 GEMM is asked for explicitly; a valid geometry without "same" padding has no direct form either
    import dataclasses
    plan_g = L.Plan(dataclasses.replace(s, path=L.PATH_GEMM))
    assert lib.qiddm_qconv_direct_supported(plan_g.handle, C.byref(good)) == 0
    assert lib.qiddm_qconv_direct_supported(plan.handle, C.byref(L.UnfoldDesc(8, 28, 28, 3, 3, 0, 0))) == 0
    assert lib.qiddm_qconv_gemm_workspace_bytes(plan_g.handle, C.byref(good), 10) > lib.qiddm_gemm_workspace_bytes(plan_g.handle, 7840) - 1
    assert lib.qiddm_qconv_gemm_forward(plan.handle, None, C.byref(good), 0, None, None, None, None, 1, 3, None) == -1
