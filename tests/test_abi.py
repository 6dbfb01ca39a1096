"""CPU: the C-ABI library loads here (no GPU) and exports every symbol include/nfn_b200.h
declares; argument validation returns error codes without touching the device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nfn_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(nfn_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if n not in ("nfn_status", "nfn_chain_desc")))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("nfn_chain_forward", "nfn_chain_forward_backward", "nfn_mdn_forward_backward",
                 "nfn_kmn_forward_backward", "nfn_flow_forward", "nfn_chain_forward_backward_host",
                 "nfn_logmeanexp_draws", "nfn_last_error", "nfn_version"):
        assert must in names


def test_library_exports_every_declared_symbol(nfn_lib):
    from normalizingflownetwork_b200 import _lib

    for name in declared_functions():
        assert hasattr(nfn_lib, name), "libnfn_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "ctypes binding is missing %s" % name
    assert set(_lib.SIGNATURES) == set(declared_functions())
    assert nfn_lib.nfn_version() == 100


def test_integration_doc_indexes_every_declared_symbol():
    """INTEGRATION.md names every entry point of the header (and nothing the header does not declare)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = declared_functions()
    missing = [n for n in names if "`%s`" % n not in text]
    assert not missing, missing
    index = text[text.index("## Every exported symbol"):text.index("## Error and threading contract")]
    listed = set(re.findall(r"`(nfn_[a-z0-9_]+)`", index))
    assert listed == set(names), (sorted(listed - set(names)), sorted(set(names) - listed))
    assert "(%d;" % len(names) in index


def test_struct_layout_matches_header():
    from normalizingflownetwork_b200 import _lib

    assert ctypes.sizeof(_lib.ChainDesc) == 12 + 64
    assert _lib.ChainDesc.flow_type.offset == 12


def test_descriptor_validation_without_gpu(nfn_lib):
    from normalizingflownetwork_b200 import _lib

    d = _lib.make_desc(["planar", "radial", "affine"], 3, True)
    assert nfn_lib.nfn_chain_param_size(ctypes.byref(d)) == 7 + 5 + 6 + 6  # reference test_total_param_size_nf
    d1 = _lib.make_desc(["planar", "radial", "affine"], 1, False)
    assert nfn_lib.nfn_chain_param_size(ctypes.byref(d1)) == 3 + 3 + 2
    assert nfn_lib.nfn_chain_is_specialized(ctypes.byref(d)) == 1
    assert nfn_lib.nfn_chain_is_specialized(ctypes.byref(_lib.make_desc(["affine"] * 7, 5, False))) == 0
    bad = _lib.ChainDesc()
    bad.n_dims, bad.n_flows, bad.trainable_base = 0, 1, 0
    assert nfn_lib.nfn_chain_param_size(ctypes.byref(bad)) == -2  # NFN_ERR_DESC
    assert b"n_dims" in nfn_lib.nfn_last_error()
    bad.n_dims, bad.n_flows = 2, 65
    assert nfn_lib.nfn_chain_param_size(ctypes.byref(bad)) == -2
    bad.n_flows = 1
    bad.flow_type[0] = 7
    assert nfn_lib.nfn_chain_param_size(ctypes.byref(bad)) == -2
    assert nfn_lib.nfn_chain_param_size(None) == -1  # NFN_ERR_NULL
    # shape / null / alignment checks happen before any CUDA call
    assert nfn_lib.nfn_chain_forward(ctypes.byref(d), None, None, 5, None, 4, None) == -3  # y_rows != B, 1
    assert nfn_lib.nfn_chain_forward(ctypes.byref(d), None, None, 4, None, 4, None) == -1
    assert nfn_lib.nfn_chain_forward(ctypes.byref(d), None, None, 0, None, 0, None) == 0   # B = 0 is a no-op
    # density grid: empty grids are no-ops, negative sizes and NULL buffers are refused
    assert nfn_lib.nfn_chain_forward_grid(ctypes.byref(d), None, None, 0, None, 4, None) == 0
    assert nfn_lib.nfn_chain_forward_grid(ctypes.byref(d), None, None, 3, None, 0, None) == 0
    assert nfn_lib.nfn_chain_forward_grid(ctypes.byref(d), None, None, -1, None, 4, None) == -3
    assert nfn_lib.nfn_chain_forward_grid(ctypes.byref(d), None, None, 3, None, 4, None) == -1
    buf = (ctypes.c_float * 64)()
    addr = ctypes.addressof(buf)
    odd = ctypes.c_void_p(addr + 4)
    assert nfn_lib.nfn_chain_forward(ctypes.byref(d), odd, ctypes.c_void_p(addr), 1, ctypes.c_void_p(addr), 1,
                                     None) == -4  # NFN_ERR_ALIGN
    assert nfn_lib.nfn_mdn_forward(0, 2, ctypes.c_void_p(addr), ctypes.c_void_p(addr), 1, ctypes.c_void_p(addr), 1,
                                   None) == -2
    assert nfn_lib.nfn_flow_forward(3, 2, None, None, 1, None, None, 1, None) == -2
    assert nfn_lib.nfn_logmeanexp_draws(None, 0, 4, None, None) == -3


def test_hidden_layer_kernel_range(nfn_lib):
    """nfn_dense_act_supported is pure host logic: 1..64 inputs, 8/16/32/64 units, five activations."""
    ok = nfn_lib.nfn_dense_act_supported
    assert ok(1, 16, 1) and ok(16, 16, 1) and ok(64, 64, 4) and ok(3, 8, 0)
    assert not ok(0, 16, 1) and not ok(65, 16, 1) and not ok(16, 10, 1) and not ok(16, 16, 5) and not ok(16, 16, -1)
    # argument checks happen before any CUDA call
    assert nfn_lib.nfn_dense_act_forward(None, None, None, 4, 16, 10, 1, None, None) == -6   # NFN_ERR_UNSUPPORTED
    assert nfn_lib.nfn_dense_act_forward(None, None, None, 0, 16, 16, 1, None, None) == 0    # B = 0 is a no-op
    assert nfn_lib.nfn_dense_act_forward(None, None, None, 4, 16, 16, 1, None, None) == -1   # NFN_ERR_NULL
    assert nfn_lib.nfn_dense_act_backward(None, None, None, None, 4, 16, 16, 1, None, None, None, None) == -1


def test_runtime_specialiser_compiles_without_gpu(nfn_lib):
    """The embedded device headers compile under NVRTC for sm_100a (no device needed)."""
    from normalizingflownetwork_b200 import _lib

    for ft, d, tb in [(["planar", "radial"], 3, True), (["affine", "radial", "planar"], 1, False), ([], 2, True)]:
        desc = _lib.make_desc(ft, d, tb)
        n = nfn_lib.nfn_jit_compile_check(ctypes.byref(desc), 0)
        assert n > 10_000, nfn_lib.nfn_last_error()
    assert nfn_lib.nfn_jit_compile_check(ctypes.byref(_lib.make_desc(["radial"], 2, True)), 1) > 10_000
    # the fused Dense(P)+chain kernels, both implementations (warp-level mma.sync; tcgen05 / TMEM)
    desc = _lib.make_desc(["radial", "planar"], 3, True)
    assert nfn_lib.nfn_jit_dense_compile_check(ctypes.byref(desc), 32, 0) > 10_000, nfn_lib.nfn_last_error()
    assert nfn_lib.nfn_jit_dense_tc5_compile_check(ctypes.byref(desc), 32, 0) > 10_000, nfn_lib.nfn_last_error()
    desc = _lib.make_desc(["radial"] * 10, 2, True)
    assert nfn_lib.nfn_jit_dense_tc5_compile_check(ctypes.byref(desc), 16, 1) > 10_000, nfn_lib.nfn_last_error()
    # not eligible: hidden width the tensor-core tiles cannot take
    assert nfn_lib.nfn_jit_dense_tc5_compile_check(ctypes.byref(desc), 10, 0) < 0
    # the fused Dense(P)+MDN kernel (same GEMM scaffolding, mixture head), a shape without an ahead-of-time instance
    assert nfn_lib.nfn_jit_dense_mdn_compile_check(7, 3, 32, 0) > 10_000, nfn_lib.nfn_last_error()
    assert nfn_lib.nfn_jit_dense_mdn_compile_check(20, 2, 16, 1) > 10_000, nfn_lib.nfn_last_error()
    assert nfn_lib.nfn_jit_dense_mdn_compile_check(200, 2, 16, 0) < 0   # a 1000-column tile does not fit shared memory
    # ... and the fused Dense(P)+KMN kernel (head with its own shared-memory state)
    assert nfn_lib.nfn_jit_dense_kmn_compile_check(60, 1, 16, 0) > 10_000, nfn_lib.nfn_last_error()
    assert nfn_lib.nfn_jit_dense_kmn_compile_check(24, 3, 32, 1) > 10_000, nfn_lib.nfn_last_error()
    assert nfn_lib.nfn_jit_dense_kmn_compile_check(2000, 1, 16, 0) < 0


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from normalizingflownetwork_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "normalizingflownetwork_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, fn


def test_round2_entry_points_validate_before_touching_the_gpu(nfn_lib):
    """The fused mixture / folded-draw / weight-posterior entry points reject bad requests with the documented status
    codes before any CUDA call (so this runs without a GPU); NULL and shape checks of include/nfn_b200.h."""
    from normalizingflownetwork_b200 import _lib
    from normalizingflownetwork_b200 import functional as F

    null = None
    one = (ctypes.c_float * 16)()
    ptr = ctypes.cast(one, ctypes.POINTER(ctypes.c_float))
    # fused Dense(P)+MDN: descriptor-like arguments, then pointers
    assert nfn_lib.nfn_dense_mdn_forward_x(0, 2, 16, ptr, ptr, ptr, ptr, 4, ptr, 4, null, null) == -2   # n_centers
    assert nfn_lib.nfn_dense_mdn_forward_x(5, 9, 16, ptr, ptr, ptr, ptr, 4, ptr, 4, null, null) == -2   # n_dims
    assert nfn_lib.nfn_dense_mdn_forward_x(5, 1, 0, ptr, ptr, ptr, ptr, 4, ptr, 4, null, null) == -3    # hidden
    assert nfn_lib.nfn_dense_mdn_forward_x(5, 1, 16, ptr, ptr, ptr, ptr, 3, ptr, 4, null, null) == -3   # y_rows vs B
    assert nfn_lib.nfn_dense_mdn_forward_x(5, 1, 16, null, ptr, ptr, ptr, 4, ptr, 4, null, null) == -1  # h is NULL
    assert nfn_lib.nfn_dense_mdn_forward_x(5, 1, 16, ptr, ptr, ptr, ptr, 0, ptr, 0, null, null) in (0, -3)  # empty batch
    # folded draws: y is per SAMPLE
    d = _lib.make_desc(["radial"] * 5, 1, True)
    assert nfn_lib.nfn_dense_chain_forward_draws_x(ctypes.byref(d), 16, 0, 8, ptr, ptr, ptr, ptr, 8, ptr, null, null) == -3
    assert nfn_lib.nfn_dense_chain_forward_draws_x(ctypes.byref(d), 16, 4, 8, ptr, ptr, ptr, ptr, 32, ptr, null, null) == -3
    assert b"per sample" in nfn_lib.nfn_last_error()
    assert nfn_lib.nfn_dense_chain_forward_draws_x(ctypes.byref(d), 16, 4, 8, null, ptr, ptr, ptr, 8, ptr, null, null) == -1
    assert nfn_lib.nfn_dense_act_forward_draws(ptr, null, null, ptr, 4, 8, 9, 10, 16, 1, ptr, null) == -6     # 9 inputs
    assert nfn_lib.nfn_dense_act_forward_draws(ptr, null, null, ptr, 4, 8, 1, 10, 12, 1, ptr, null) == -6     # row width % 8
    assert nfn_lib.nfn_dense_act_forward_draws(ptr, ptr, null, ptr, 4, 8, 1, 10, 16, 1, ptr, null) == -1      # mean without std
    assert nfn_lib.nfn_dense_act_backward_draws(ptr, null, null, ptr, null, 4, 8, 1, 10, 16, 1, ptr, null) == -1
    # mean-field weight posterior
    assert nfn_lib.nfn_variational_sample(ptr, ptr, ctypes.c_float(1.0), ptr, 0, 4, ptr, null, null) == -3
    assert nfn_lib.nfn_variational_sample(ptr, ptr, ctypes.c_float(0.0), ptr, 8, 4, ptr, null, null) == -2    # prior scale
    assert nfn_lib.nfn_variational_sample(null, ptr, ctypes.c_float(1.0), ptr, 8, 4, ptr, null, null) == -1
    assert nfn_lib.nfn_variational_sample_backward(ptr, ptr, ctypes.c_float(1.0), null, ptr, null, 8, 4, ptr, null, null) == -1
    # the shape helpers the estimators consult mirror the kernels' limits
    assert F.dense_mdn_supported(16, 20, 2) and F.dense_mdn_supported(64, 5, 1)
    assert not F.dense_mdn_supported(16, 200, 2) and not F.dense_mdn_supported(24, 5, 1)
    assert F.dense_act_draws_supported(1, 10, 16, "tanh") and not F.dense_act_draws_supported(9, 10, 16, "tanh")
    assert not F.dense_act_draws_supported(1, 10, 12, "tanh") and not F.dense_act_draws_supported(1, 10, 16, "gelu")
