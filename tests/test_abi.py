"""The C-ABI library builds, loads, and exports every symbol include/llama_b200.h declares.
No compute calls: this runs on the CPU-only build box."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "llama_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    fns = declared_functions()
    for must in ["b200_ctx_create", "b200_ctx_upload_tensor", "b200_ctx_finalize", "b200_forward", "b200_prefill_token",
                 "b200_reset", "b200_position", "b200_ctx_destroy", "b200_last_error", "b200_op_vec_mat_q",
                 "b200_op_dequantize", "b200_op_rms_norm", "b200_op_rope", "b200_op_attention_cached"]:
        assert must in fns


def test_library_exports_every_declared_symbol(b200):
    L = ctypes.CDLL(b200.LIB_PATH)
    missing = [f for f in declared_functions() if not hasattr(L, f)]
    assert not missing, f"declared in include/llama_b200.h but not exported: {missing}"


def test_backend_name_and_block_tables(b200):
    L = b200.lib()
    assert L.b200_backend_name() == b"cuda-b200"
    for t, (be, bb) in b200.BLOCK.items():
        assert L.b200_type_block_elems(t) == be and L.b200_type_block_bytes(t) == bb
    assert L.b200_type_block_elems(99) == 0


def test_no_cpu_fallback_without_a_gpu(b200):
    """On a box without CUDA every compute entry point must fail loudly with NotAvailable."""
    import numpy as np

    if b200.device_count() > 0:
        pytest.skip("a CUDA device is present")
    be = b200.CudaB200Backend()
    assert be.name() == "cuda-b200" and not be.is_available()
    with pytest.raises(b200.NotAvailable):
        be.add(np.ones(4, np.float32), np.ones(4, np.float32))
    with pytest.raises(b200.NotAvailable):
        be.dequantize(np.zeros(34, np.uint8), b200.Q8_0, 32)
    with pytest.raises(b200.NotAvailable):
        b200.GpuOnlyInference({"hidden": 64, "n_layers": 1, "n_heads": 1, "n_kv_heads": 1, "ffn": 64, "vocab": 8,
                               "max_seq_len": 8}, {})


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under llama-gguf_b200/ may reference it."""
    pkg = os.path.join(ROOT, "llama-gguf_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "oracle.cpp" not in txt, f
