"""llama-gguf `cuda-b200` backend — host-side mirror of the reference interface.

The product is `libllama_b200.so` (C ABI: include/llama_b200.h, sources in csrc/).
The reference's host language is Rust, which this image cannot compile, so the
classes below play the role of the Rust shim of INTEGRATION.md: same names,
argument meaning and error behaviour as

  * `Backend`            -> CudaB200Backend   (src/backend/mod.rs:29-265)
  * `GpuInference`       -> GpuOnlyInference  (src/backend/mod.rs:283-296,
                                               src/backend/cuda/gpu_only.rs:426-845)
  * `BackendError`       -> BackendError + one subclass per variant
                                              (src/backend/error.rs:3-37)

There is NO CPU fallback: if the shared library is missing or no CUDA device is
present every entry point raises NotAvailable.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "libllama_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

# ggml type ids (src/gguf/constants.rs:56-89)
F32, F16, Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, Q6_K = 0, 1, 2, 6, 8, 12, 13, 14
TYPE_NAMES = {F32: "F32", F16: "F16", Q4_0: "Q4_0", Q5_0: "Q5_0", Q8_0: "Q8_0", Q4_K: "Q4_K", Q5_K: "Q5_K", Q6_K: "Q6_K"}
BLOCK = {F32: (1, 4), F16: (1, 2), Q4_0: (32, 18), Q5_0: (32, 22), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176), Q6_K: (256, 210)}


# ----------------------------------------------------------------- errors
class BackendError(Exception):
    """src/backend/error.rs:3-37"""


class NotAvailable(BackendError): pass
class ShapeMismatch(BackendError): pass
class DTypeMismatch(BackendError): pass
class UnsupportedDType(BackendError): pass
class Unsupported(BackendError): pass
class InvalidArgument(BackendError): pass
class TensorError(BackendError): pass
class InitializationFailed(BackendError): pass
class AllocationFailed(BackendError): pass
class OperationFailed(BackendError): pass


_STATUS = {-1: NotAvailable, -2: ShapeMismatch, -3: DTypeMismatch, -4: UnsupportedDType, -5: Unsupported,
           -6: InvalidArgument, -7: TensorError, -8: InitializationFailed, -9: AllocationFailed, -10: OperationFailed}


class ModelDesc(C.Structure):
    """b200_model_desc (include/llama_b200.h)"""

    _fields_ = [
        ("hidden", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("ffn", C.c_int32), ("vocab", C.c_int32), ("max_seq_len", C.c_int32),
        ("norm_eps", C.c_float), ("rope_base", C.c_float), ("rope_scale", C.c_float),
        ("rope_neox", C.c_int32), ("n_experts", C.c_int32), ("n_experts_used", C.c_int32),
        ("expert_ffn", C.c_int32), ("tied_output", C.c_int32), ("max_batch", C.c_int32),
    ]


class ParallelDesc(C.Structure):
    _fields_ = [("world_size", C.c_int32), ("rank", C.c_int32), ("device", C.c_int32)]


class LoadStats(C.Structure):
    """b200_load_stats (include/llama_b200.h)."""
    _fields_ = [("file_bytes", C.c_uint64), ("tensor_bytes", C.c_uint64), ("device_bytes", C.c_uint64),
                ("tensors_loaded", C.c_uint32), ("tensors_skipped", C.c_uint32), ("seconds", C.c_double)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["GBps"] = d["device_bytes"] / d["seconds"] / 1e9 if d["seconds"] > 0 else 0.0
        return d


class BatchConfig(C.Structure):
    """b200_batch_config = BatchedEngineConfig (src/engine_batched.rs:23-40)."""
    _fields_ = [("max_batch_size", C.c_int32), ("max_seq_len", C.c_int32), ("max_queue_depth", C.c_int32), ("eos_token_id", C.c_uint32)]


class BatchEvent(C.Structure):
    """b200_batch_event = BatchToken (src/engine_batched.rs:60-72)."""
    _fields_ = [("request_id", C.c_uint64), ("kind", C.c_int32), ("token", C.c_uint32), ("reason", C.c_int32),
                ("prompt_tokens", C.c_int32), ("completion_tokens", C.c_int32)]


def build(verbose=False):
    """Compile csrc/ into libllama_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Raises NotAvailable if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NotAvailable(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(cuda-b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    fp, vp, u64p = C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_uint64)
    L.b200_backend_name.restype = C.c_char_p
    L.b200_last_error.restype = C.c_char_p
    L.b200_device_count.argtypes = [C.POINTER(C.c_int)]
    L.b200_type_block_elems.argtypes = [C.c_uint32]
    L.b200_type_block_bytes.argtypes = [C.c_uint32]
    L.b200_ctx_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(ParallelDesc), C.POINTER(vp)]
    L.b200_ctx_upload_tensor.argtypes = [vp, C.c_char_p, C.c_uint32, u64p, C.c_int, vp, C.c_size_t]
    L.b200_ctx_finalize.argtypes = [vp]
    L.b200_ctx_set_kv_format.argtypes = [vp, C.c_int]
    L.b200_ctx_kv_format.argtypes = [vp, C.POINTER(C.c_int)]
    L.b200_ctx_destroy.argtypes = [vp]
    L.b200_ctx_destroy.restype = None
    u32p, i32p, szp = C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.POINTER(C.c_size_t)
    L.b200_ctx_set_speculation.argtypes = [vp, C.c_int]
    L.b200_ctx_speculation_stats.argtypes = [vp, i32p, u64p, u64p]
    L.b200_debug_set_position.argtypes = [vp, C.c_int, C.c_uint64]
    L.b200_debug_plan_split.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    L.b200_decode_batch_greedy.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.c_int, C.POINTER(C.c_uint32)]
    L.b200_batch_create.argtypes = [vp, C.POINTER(BatchConfig), C.POINTER(vp)]
    L.b200_batch_destroy.argtypes = [vp]
    L.b200_batch_destroy.restype = None
    L.b200_batch_submit.argtypes = [vp, C.POINTER(C.c_uint32), C.c_int, C.c_int, u64p]
    L.b200_batch_step.argtypes = [vp, C.POINTER(BatchEvent), C.c_int, i32p]
    L.b200_batch_counts.argtypes = [vp, i32p, i32p, i32p, u64p, u64p]
    L.b200_batch_last_error.argtypes = [vp]
    L.b200_batch_last_error.restype = C.c_char_p
    L.b200_gguf_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.b200_gguf_close.argtypes = [vp]
    L.b200_gguf_close.restype = None
    L.b200_gguf_info.argtypes = [vp, u32p, u64p, u64p, u64p, u64p, u64p]
    L.b200_gguf_architecture.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.b200_gguf_model_desc.argtypes = [vp, C.c_int, C.c_int, C.POINTER(ModelDesc)]
    L.b200_gguf_tensor_info.argtypes = [vp, C.c_uint64, C.POINTER(C.c_char_p), u32p, u64p, i32p, C.POINTER(vp), szp]
    L.b200_ctx_load_gguf.argtypes = [vp, vp, C.POINTER(LoadStats)]
    L.b200_ctx_create_from_gguf.argtypes = [C.c_char_p, C.POINTER(ParallelDesc), C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(LoadStats)]
    L.b200_group_load_gguf.argtypes = [vp, vp, C.POINTER(LoadStats)]
    L.b200_forward.argtypes = [vp, C.c_int, C.c_uint32, fp]
    L.b200_prefill_token.argtypes = [vp, C.c_int, C.c_uint32]
    L.b200_prefill.argtypes = [vp, C.c_int, C.POINTER(C.c_uint32), C.c_int, fp]
    L.b200_decode_batch.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.c_int, fp]
    L.b200_reset.argtypes = [vp, C.c_int]
    L.b200_position.argtypes = [vp, C.c_int, u64p]
    L.b200_decode_greedy.argtypes = [vp, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), fp]
    L.b200_get_hidden.argtypes = [vp, C.c_int, C.c_int, fp]
    L.b200_ctx_stats.argtypes = [vp, u64p, u64p, u64p]
    L.b200_ctx_tp_handle.argtypes = [vp, vp]
    L.b200_ctx_tp_set_peer.argtypes = [vp, C.c_int, vp]
    try:   # single-process group (one context + one host thread per device inside the library)
        L.b200_group_create.argtypes = [C.POINTER(ModelDesc), C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
        L.b200_group_upload_tensor.argtypes = [vp, C.c_char_p, C.c_uint32, u64p, C.c_int, vp, C.c_size_t]
        L.b200_group_finalize.argtypes = [vp]
        L.b200_group_destroy.argtypes = [vp]
        L.b200_group_destroy.restype = None
        L.b200_group_forward.argtypes = [vp, C.c_int, C.c_uint32, fp]
        L.b200_group_prefill_token.argtypes = [vp, C.c_int, C.c_uint32]
        L.b200_group_reset.argtypes = [vp, C.c_int]
        L.b200_group_position.argtypes = [vp, C.c_int, u64p]
        L.b200_group_decode_greedy.argtypes = [vp, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), fp]
        L.b200_group_size.argtypes = [vp, C.POINTER(C.c_int)]
        L.b200_group_ctx.argtypes = [vp, C.c_int, C.POINTER(vp)]
    except AttributeError:  # an older build of the library (B200_LIB override)
        pass
    L.b200_debug_mega_timeline.argtypes = [vp, u64p, C.c_int]
    L.b200_debug_err.argtypes = [vp, C.POINTER(C.c_int)]
    L.b200_ctx_path.argtypes = [vp, C.POINTER(C.c_int)]
    L.b200_debug_mega_phase.argtypes = [vp, C.c_int, u64p, C.c_int]
    try:
        L.b200_debug_read.argtypes = [vp, C.c_int, fp, C.c_int]
    except AttributeError:  # an older build of the library (B200_LIB override)
        pass
    L.b200_bench_weight_gemv.argtypes = [vp, C.c_char_p, C.c_int, fp, u64p]
    L.b200_bench_gemv_pass.argtypes = [vp, C.c_int, C.c_int, fp, u64p, u64p]
    L.b200_op_add.argtypes = [fp, fp, fp, C.c_size_t]
    L.b200_op_mul.argtypes = [fp, fp, fp, C.c_size_t]
    L.b200_op_scale.argtypes = [fp, C.c_float, fp, C.c_size_t]
    L.b200_op_silu.argtypes = [fp, fp, C.c_size_t]
    L.b200_op_gelu.argtypes = [fp, fp, C.c_size_t]
    L.b200_op_softmax.argtypes = [fp, fp, C.c_size_t]
    L.b200_op_rms_norm.argtypes = [fp, fp, C.c_float, fp, C.c_size_t, C.c_size_t]
    L.b200_op_vec_mat.argtypes = [fp, fp, fp, C.c_size_t, C.c_size_t]
    L.b200_op_matmul.argtypes = [fp, fp, fp, C.c_size_t, C.c_size_t, C.c_size_t]
    L.b200_op_matvec.argtypes = [fp, fp, fp, C.c_size_t, C.c_size_t]
    L.b200_op_matvec_q.argtypes = [vp, C.c_uint32, fp, fp, C.c_size_t, C.c_size_t]
    L.b200_op_vec_mat_q.argtypes = [fp, vp, C.c_uint32, fp, C.c_size_t, C.c_size_t]
    L.b200_op_mat_mat_q.argtypes = [fp, vp, C.c_uint32, fp, C.c_size_t, C.c_size_t, C.c_size_t]
    L.b200_op_dequantize.argtypes = [vp, C.c_uint32, fp, C.c_size_t]
    L.b200_op_rope.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
    L.b200_op_attention_cached.argtypes = [fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
    L.b200_op_attention.argtypes = [fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        msg = lib().b200_last_error().decode("utf-8", "replace")
        raise _STATUS.get(rc, OperationFailed)(msg)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a, name="tensor"):
    a = np.asarray(a)
    if a.dtype != np.float32:
        raise DTypeMismatch(f"{name}: expected F32, got {a.dtype}")
    return np.ascontiguousarray(a)


def device_count():
    n = C.c_int(0)
    rc = lib().b200_device_count(C.byref(n))
    return n.value if rc == 0 else 0


# ----------------------------------------------------------------- Backend
class CudaB200Backend:
    """`impl Backend for CudaB200Backend` (trait: src/backend/mod.rs:29-265).

    Tensors are numpy arrays (f32 activations; uint8 raw GGUF blocks for quantised
    weights, with the ggml type id passed alongside).  Like the reference's per-op CUDA
    path every call is H2D -> kernel -> D2H; performance lives behind GpuOnlyInference.
    """

    def name(self):
        return lib().b200_backend_name().decode()

    def is_available(self):
        try:
            return device_count() > 0
        except BackendError:
            return False

    def alloc(self, shape, dtype=F32):
        if dtype != F32:
            raise UnsupportedDType(f"alloc: {dtype}")
        return np.zeros(shape, dtype=np.float32)

    def copy_to(self, tensor):
        return np.array(tensor, copy=True)

    @staticmethod
    def _same_shape(a, b, what):
        if a.shape != b.shape:
            raise ShapeMismatch(f"{what}: expected {list(a.shape)}, got {list(b.shape)}")

    def add(self, a, b):
        a, b = _f32(a), _f32(b)
        self._same_shape(a, b, "add")
        out = np.empty_like(a)
        _check(lib().b200_op_add(_fp(a), _fp(b), _fp(out), a.size))
        return out

    def mul(self, a, b):
        a, b = _f32(a), _f32(b)
        self._same_shape(a, b, "mul")
        out = np.empty_like(a)
        _check(lib().b200_op_mul(_fp(a), _fp(b), _fp(out), a.size))
        return out

    def scale(self, a, scalar):
        a = _f32(a)
        out = np.empty_like(a)
        _check(lib().b200_op_scale(_fp(a), float(scalar), _fp(out), a.size))
        return out

    def silu(self, x):
        x = _f32(x)
        out = np.empty_like(x)
        _check(lib().b200_op_silu(_fp(x), _fp(out), x.size))
        return out

    def gelu(self, x):
        x = _f32(x)
        out = np.empty_like(x)
        _check(lib().b200_op_gelu(_fp(x), _fp(out), x.size))
        return out

    def softmax(self, x):
        x = _f32(x)
        out = np.empty_like(x)
        n = x.shape[-1] if x.ndim else 1
        for r in range(x.size // max(n, 1)):
            row, orow = x.reshape(-1, n)[r], out.reshape(-1, n)[r]
            _check(lib().b200_op_softmax(_fp(row), _fp(orow), n))
        return out

    def rms_norm(self, x, weight, eps):
        x, weight = _f32(x), _f32(weight, "weight")
        hidden = x.shape[-1]
        if weight.shape != (hidden,):
            raise ShapeMismatch(f"rms_norm: expected [{hidden}], got {list(weight.shape)}")
        out = np.empty_like(x)
        _check(lib().b200_op_rms_norm(_fp(x), _fp(weight), float(eps), _fp(out), x.size // hidden, hidden))
        return out

    def vec_mat(self, a, b):
        """a [k] @ b [k, n] (GGUF layout: n rows of k contiguous) -> [n]."""
        a, b = _f32(a), _f32(b, "b")
        if a.ndim != 1 or b.ndim != 2:
            raise InvalidArgument("vec_mat requires 1D vector and 2D matrix")
        # numpy shape is (n, k) for GGUF dims [k, n]
        n, k = b.shape
        if a.shape[0] != k:
            raise ShapeMismatch(f"vec_mat: expected [{a.shape[0]}], got [{k}]")
        out = np.empty(n, dtype=np.float32)
        _check(lib().b200_op_vec_mat(_fp(a), _fp(b), _fp(out), k, n))
        return out

    def mat_mat_q(self, a, raw, ggml_type, k, n):
        """a [t, k] f32, W [k, n] quantised (raw GGUF blocks) -> [t, n] through the tcgen05 dequant-GEMM (fp16 operands)."""
        a = _f32(a)
        if a.ndim != 2 or a.shape[1] != k:
            raise ShapeMismatch(f"mat_mat_q: expected [t, {k}], got {list(a.shape)}")
        if ggml_type not in BLOCK:
            raise UnsupportedDType(f"mat_mat_q: ggml type {ggml_type}")
        raw = np.ascontiguousarray(raw).view(np.uint8).ravel()
        be, bb = BLOCK[ggml_type]
        if k % be or raw.size != k // be * bb * n:
            raise ShapeMismatch(f"mat_mat_q: {raw.size} bytes do not match [{k}, {n}] of {TYPE_NAMES[ggml_type]}")
        out = np.empty((a.shape[0], n), dtype=np.float32)
        _check(lib().b200_op_mat_mat_q(_fp(a), raw.ctypes.data, ggml_type, _fp(out), a.shape[0], k, n))
        return out

    def vec_mat_q(self, a, raw, ggml_type, k, n):
        """a [k] f32, W [k, n] quantised (raw GGUF blocks, n rows of k/bs blocks) -> [n]."""
        a = _f32(a)
        if a.ndim != 1:
            raise InvalidArgument("vec_mat_q requires 1D vector and 2D quantized matrix")
        if a.shape[0] != k:
            raise ShapeMismatch(f"vec_mat_q: expected [{k}], got {list(a.shape)}")
        if ggml_type not in BLOCK:
            raise UnsupportedDType(f"vec_mat_q: ggml type {ggml_type}")
        raw = np.ascontiguousarray(raw).view(np.uint8).ravel()
        be, bb = BLOCK[ggml_type]
        if k % be or raw.size != k // be * bb * n:
            raise ShapeMismatch(f"vec_mat_q: {raw.size} bytes do not match [{k}, {n}] of {TYPE_NAMES[ggml_type]}")
        out = np.empty(n, dtype=np.float32)
        _check(lib().b200_op_vec_mat_q(_fp(a), raw.ctypes.data, ggml_type, _fp(out), k, n))
        return out

    def matvec_q(self, raw, ggml_type, x, m, k):
        """A [m rows of k] quantised @ x [k] -> [m] (Backend::matvec_q, cpu/ops.rs:922-946: the block walk of vec_mat_q)."""
        if ggml_type not in BLOCK:
            raise UnsupportedDType(f"matvec_q: ggml type {ggml_type}")
        be, bb = BLOCK[ggml_type]
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.size != k or k % be or raw.size != m * (k // be) * bb:
            raise ShapeMismatch(f"matvec_q: a {raw.size} bytes for [{m}, {k}] type {ggml_type}, b {x.size}")
        out = np.empty(m, dtype=np.float32)
        _check(lib().b200_op_matvec_q(raw.ctypes.data_as(C.c_void_p), ggml_type, _fp(x), _fp(out), m, k))
        return out

    def matmul(self, a, b):
        """Backend::matmul (cpu/ops.rs:429-487): a [m, k] @ b [k, n], row-major f32."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        if a.ndim != 2 or b.ndim != 2:
            raise InvalidArgument("matmul requires 2D tensors")
        if a.shape[1] != b.shape[0]:
            raise ShapeMismatch(f"matmul: expected {[a.shape[0], a.shape[1]]}, got {list(b.shape)}")
        out = np.empty((a.shape[0], b.shape[1]), dtype=np.float32)
        _check(lib().b200_op_matmul(_fp(a), _fp(b), _fp(out), a.shape[0], a.shape[1], b.shape[1]))
        return out

    def matvec(self, a, b):
        """Backend::matvec (cpu/ops.rs:531-575): a [m, k] @ b [k] -> [m]."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        if a.ndim != 2 or b.ndim != 1:
            raise InvalidArgument("matvec requires 2D matrix and 1D vector")
        if b.shape[0] != a.shape[1]:
            raise ShapeMismatch(f"matvec: expected {[a.shape[1]]}, got {list(b.shape)}")
        out = np.empty(a.shape[0], dtype=np.float32)
        _check(lib().b200_op_matvec(_fp(a), _fp(b), _fp(out), a.shape[0], a.shape[1]))
        return out

    def dequantize(self, raw, ggml_type, n_elems):
        if ggml_type not in BLOCK:
            raise UnsupportedDType(f"dequantize: ggml type {ggml_type}")
        raw = np.ascontiguousarray(raw).view(np.uint8).ravel()
        be, bb = BLOCK[ggml_type]
        if n_elems % be or raw.size != n_elems // be * bb:
            raise ShapeMismatch(f"dequantize: expected [{raw.size // bb * be}], got [{n_elems}]")
        out = np.empty(n_elems, dtype=np.float32)
        _check(lib().b200_op_dequantize(raw.ctypes.data, ggml_type, _fp(out), n_elems))
        return out

    def rope(self, q, k, pos, freq_base, freq_scale, use_neox):
        """q [num_heads, seq_len=1, head_dim], k [num_kv_heads, 1, head_dim]; returns rotated copies."""
        q, k = _f32(q).copy(), _f32(k).copy()
        if q.ndim != 3 or k.ndim != 3:
            raise InvalidArgument("RoPE requires 3D tensors [num_heads, seq_len, head_dim]")
        if q.shape[1] != k.shape[1] or q.shape[2] != k.shape[2]:
            raise InvalidArgument("Q and K must have same seq_len and head_dim")
        if q.shape[1] != 1:
            raise Unsupported("rope: seq_len must be 1 on the decode path")
        _check(lib().b200_op_rope(_fp(q), _fp(k), q.shape[0], k.shape[0], q.shape[2], int(pos), float(freq_base),
                                  float(freq_scale), int(bool(use_neox))))
        return q, k

    def attention(self, q, k, v, scale):
        q, k, v = _f32(q), _f32(k), _f32(v)
        if q.ndim != 3 or k.ndim != 3 or v.ndim != 3:
            raise InvalidArgument("Attention requires 3D tensors")
        if k.shape != v.shape or k.shape[2] != q.shape[2] or k.shape[1] != q.shape[1]:
            raise InvalidArgument("Attention tensor dimension mismatch")
        out = np.empty_like(q)
        _check(lib().b200_op_attention(_fp(q), _fp(k), _fp(v), _fp(out), q.shape[0], k.shape[0], q.shape[1], q.shape[2],
                                       float(scale)))
        return out

    def attention_cached(self, q, k_cache, v_cache, scale, kv_len):
        """q [nh, 1, hd]; caches [nkv, max_seq, hd]; first kv_len positions valid."""
        q, k_cache, v_cache = _f32(q), _f32(k_cache), _f32(v_cache)
        nh, hd = q.shape[0], q.shape[-1]
        nkv, max_seq = k_cache.shape[0], k_cache.shape[1]
        out = np.empty((nh, 1, hd), dtype=np.float32)
        _check(lib().b200_op_attention_cached(_fp(q), _fp(k_cache), _fp(v_cache), _fp(out), nh, nkv, hd, max_seq,
                                              float(scale), int(kv_len)))
        return out


# ----------------------------------------------------------------- GpuInference
DESC_KEYS = [f[0] for f in ModelDesc._fields_]


def plan_split(n_rows, K, T, n_sm=148, persistent=True):
    """K range per split the dequant-GEMM would use (0 = unsplit): host logic, no device needed."""
    v = C.c_int(0)
    _check(lib().b200_debug_plan_split(int(n_rows), int(K), int(T), int(n_sm), 1 if persistent else 0, C.byref(v)))
    return v.value


class BatchedEngine:
    """Continuous batching over the sequence slots of one context (b200_batch_*): the peer of BatchedEngine
    (src/engine_batched.rs:114-197) below the tokenizer and the channels.  `submit` = BatchedEngine::submit, `step` = one iteration
    of its background loop; events are ("token", request_id, token_id), ("done", request_id, reason, prompt_tokens,
    completion_tokens) with reason "stop" | "max_tokens", and ("error", request_id, message)."""

    REASONS = ("stop", "max_tokens", "error")

    def __init__(self, gpu, max_batch_size=8, max_seq_len=4096, max_queue_depth=64, eos_token_id=2):
        self._gpu = gpu          # keeps the context alive
        self._h = C.c_void_p()
        cfg = BatchConfig(max_batch_size, max_seq_len, max_queue_depth, eos_token_id)
        _check(lib().b200_batch_create(gpu._h, C.byref(cfg), C.byref(self._h)))
        self._cap = max(64, 2 * max_batch_size)
        self._buf = (BatchEvent * self._cap)()

    def close(self):
        if getattr(self, "_h", None):
            lib().b200_batch_destroy(self._h)
            self._h = None

    __del__ = close

    def submit(self, tokens, max_tokens):
        """Request id.  Raises OperationFailed("queue full") past max_queue_depth."""
        arr = (C.c_uint32 * max(1, len(tokens)))(*tokens)
        rid = C.c_uint64()
        _check(lib().b200_batch_submit(self._h, arr, len(tokens), int(max_tokens), C.byref(rid)))
        return rid.value

    def step(self):
        n = C.c_int()
        _check(lib().b200_batch_step(self._h, self._buf, self._cap, C.byref(n)))
        out = []
        for i in range(n.value):
            e = self._buf[i]
            if e.kind == 0:
                out.append(("token", e.request_id, e.token))
            elif e.kind == 1:
                out.append(("done", e.request_id, self.REASONS[e.reason], e.prompt_tokens, e.completion_tokens))
            else:
                out.append(("error", e.request_id, lib().b200_batch_last_error(self._h).decode()))
        return out

    def counts(self):
        a, p, u = C.c_int(), C.c_int(), C.c_int()
        st, rows = C.c_uint64(), C.c_uint64()
        _check(lib().b200_batch_counts(self._h, C.byref(a), C.byref(p), C.byref(u), C.byref(st), C.byref(rows)))
        return dict(active=a.value, pending=p.value, undelivered_events=u.value, steps=st.value, decode_rows=rows.value)

    def run(self, max_steps=100000):
        """Step until nothing is active or pending; all events in order."""
        out = []
        for _ in range(max_steps):
            c = self.counts()
            if c["active"] == 0 and c["pending"] == 0 and c["undelivered_events"] == 0:
                break
            out.extend(self.step())
        return out


class GgufFile:
    """A GGUF file mapped and parsed by the library (b200_gguf_*; no CUDA device needed): what GgufFile::open + GgufReader::read +
    ModelLoader::parse_config give the reference (src/gguf/mod.rs:23-54, reader.rs:28-110, model/loader.rs:62-300)."""

    def __init__(self, path):
        self._h = C.c_void_p()
        _check(lib().b200_gguf_open(os.fsencode(path), C.byref(self._h)))
        self.path = path

    def close(self):
        if getattr(self, "_h", None):
            lib().b200_gguf_close(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def info(self):
        ver = C.c_uint32()
        v = [C.c_uint64() for _ in range(5)]
        _check(lib().b200_gguf_info(self._h, C.byref(ver), *[C.byref(x) for x in v]))
        return dict(version=ver.value, n_tensors=v[0].value, n_metadata=v[1].value, alignment=v[2].value, data_offset=v[3].value,
                    file_bytes=v[4].value)

    def architecture(self):
        buf = C.create_string_buffer(128)
        _check(lib().b200_gguf_architecture(self._h, buf, 128))
        return buf.value.decode()

    def model_desc(self, max_seq_len=0, max_batch=1):
        d = ModelDesc()
        _check(lib().b200_gguf_model_desc(self._h, int(max_seq_len), int(max_batch), C.byref(d)))
        return {k: getattr(d, k) for k in DESC_KEYS}

    def tensor(self, i, with_data=False):
        """(name, ggml_type, ne list, nbytes[, uint8 copy of the bytes or None for a type the engine does not implement])."""
        name, t, ne, nd = C.c_char_p(), C.c_uint32(), (C.c_uint64 * 4)(), C.c_int()
        data, nb = C.c_void_p(), C.c_size_t()
        _check(lib().b200_gguf_tensor_info(self._h, i, C.byref(name), C.byref(t), ne, C.byref(nd), C.byref(data) if with_data else None,
                                           C.byref(nb)))
        out = (name.value.decode(), t.value, [int(ne[k]) for k in range(nd.value)], nb.value)
        if with_data:
            arr = None
            if data.value:
                arr = np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_uint8)), shape=(nb.value,)).copy()
            out = out + (arr,)
        return out

    def tensors(self, with_data=False):
        return [self.tensor(i, with_data) for i in range(self.info()["n_tensors"])]


class GpuOnlyInference:
    """`impl GpuInference for GpuOnlyInference` on cuda-b200.

    Mirrors src/backend/cuda/gpu_only.rs: from_model (:426) consumes the model's tensors
    and uploads them once; forward (:728) / prefill_token (:792) / forward_batch (:776) /
    reset (:808) / position (:845).  `tensors` maps GGUF tensor names to
    (ggml_type, ne, ndarray) with ne[0] = in_features.
    """

    def __init__(self, desc: dict, tensors: dict, device=0, taps=False, feeder=None, parallel=None, exchange=None, kv_format="f32"):
        """`feeder(upload)` may stream tensors one by one (upload(name, type, ne, data)) instead of `tensors`.

        Tensor parallel: `parallel=(world_size, rank)` makes this the context of one rank (one process per GPU);
        `exchange(handle_bytes) -> [handle_bytes of every rank]` is the all-gather the host provides
        (parallel.all_gather_bytes over torch.distributed).  Every rank uploads the FULL tensors; the library keeps
        its shard.  forward() then returns this rank's slice of the logits (see parallel.TensorParallelInference).

        `kv_format`: "f32" (KVCache, src/model/mod.rs:83-108) or "int8" (QuantizedKVCache's Int8 format,
        src/model/kv_quantized.rs:143-216: one scale per (kv head, position) row, a quarter of the bytes)."""
        L = lib()
        if device_count() == 0:
            raise NotAvailable("cuda-b200: no CUDA device (this backend has no CPU fallback)")
        d = ModelDesc()
        for k, v in desc.items():
            if k not in DESC_KEYS:
                raise InvalidArgument(f"unknown model desc field {k}")
            setattr(d, k, v)
        self.desc = {k: getattr(d, k) for k in DESC_KEYS}
        world, rank = parallel if parallel else (1, 0)
        par = ParallelDesc(world, rank, device)
        self.world, self.rank = world, rank
        h = C.c_void_p()
        old = os.environ.get("B200_TAPS")
        if taps:
            os.environ["B200_TAPS"] = "1"
        try:
            _check(L.b200_ctx_create(C.byref(d), C.byref(par), C.byref(h)))
        finally:
            if taps:
                if old is None:
                    os.environ.pop("B200_TAPS", None)
                else:
                    os.environ["B200_TAPS"] = old
        self._h = h
        try:
            if kv_format not in ("f32", "int8"):
                raise Unsupported(f"kv_format {kv_format!r}: f32 or int8")
            if kv_format == "int8":
                _check(L.b200_ctx_set_kv_format(h, 1))
            if world > 1:
                if exchange is None:
                    raise InvalidArgument("tensor parallel contexts need an `exchange` (all-gather of IPC handles)")
                mine = C.create_string_buffer(64)
                _check(L.b200_ctx_tp_handle(self._h, mine))
                handles = exchange(mine.raw)
                for r, hb in enumerate(handles):
                    if r != rank:
                        _check(L.b200_ctx_tp_set_peer(self._h, r, C.create_string_buffer(hb, 64)))
            for name, (t, ne, data) in (tensors or {}).items():
                self.upload_tensor(name, t, ne, data)
            if feeder is not None:
                feeder(self.upload_tensor)
            _check(L.b200_ctx_finalize(self._h))
        except Exception:
            self.close()
            raise
        # tensor parallel: this rank's slice of the logits; expert parallel (MoE + world > 1): the head is replicated, full logits
        self.expert_parallel = world > 1 and self.desc.get("n_experts", 0) > 0
        self.vocab = self.desc["vocab"] if self.expert_parallel else self.desc["vocab"] // world

    @classmethod
    def from_model(cls, model, max_seq_len, **kw):
        """`model` = (desc dict, tensors dict) as produced by gguf_io.load_gguf / the synthetic generators."""
        desc, tensors = model
        desc = dict(desc)
        desc["max_seq_len"] = min(int(max_seq_len), desc.get("max_seq_len", max_seq_len)) if max_seq_len else desc["max_seq_len"]
        return cls(desc, tensors, **kw)

    @classmethod
    def from_gguf(cls, path, max_seq_len=0, max_batch=1, device=0, kv_format="f32"):
        """GpuOnlyInference::from_model for a file path, entirely inside the library (b200_ctx_create_from_gguf): mmap, parse,
        pinned double-buffered upload of every tensor the engine uses, finalize.  `load_stats` holds bytes / seconds / GB/s."""
        L = lib()
        if device_count() == 0:
            raise NotAvailable("cuda-b200: no CUDA device (this backend has no CPU fallback)")
        if kv_format not in ("f32", "int8"):
            raise Unsupported(f"kv_format {kv_format!r}: f32 or int8")
        self = cls.__new__(cls)
        with GgufFile(path) as f:
            self.desc = f.model_desc(max_seq_len, max_batch)
            self.arch = f.architecture()
        par = ParallelDesc(1, 0, device)
        h, st = C.c_void_p(), LoadStats()
        self._h = None
        _check(L.b200_ctx_create_from_gguf(os.fsencode(path), C.byref(par), int(max_seq_len), int(max_batch),
                                           1 if kv_format == "int8" else 0, C.byref(h), C.byref(st)))
        self._h = h
        self.load_stats = st.as_dict()
        self.world, self.rank, self.expert_parallel = 1, 0, False
        self.vocab = self.desc["vocab"]
        return self

    def upload_tensor(self, name, ggml_type, ne, data):
        a = np.ascontiguousarray(data)
        nd = (C.c_uint64 * 4)(*(list(ne) + [1] * (4 - len(ne))))
        _check(lib().b200_ctx_upload_tensor(self._h, name.encode(), int(ggml_type), nd, len(ne), a.ctypes.data, a.nbytes))

    def close(self):
        if getattr(self, "_h", None):
            lib().b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def kv_format(self):
        v = C.c_int(0)
        _check(lib().b200_ctx_kv_format(self._h, C.byref(v)))
        return ("f32", "int8")[v.value]

    def path(self):
        """Decode path chosen at finalize: "graph" | "mega" | "stream" | "stream2"."""
        v = C.c_int(0)
        _check(lib().b200_ctx_path(self._h, C.byref(v)))
        return ("graph", "mega", "stream", "stream2")[v.value]

    def watchdog(self):
        """The eight watchdog words of the megakernel paths (all zero = no bounded wait ever gave up); clears them."""
        e = (C.c_int * 8)()
        _check(lib().b200_debug_err(self._h, e))
        return list(e)

    # -- GpuInference ------------------------------------------------------
    def forward(self, token_id, seq=0):
        logits = np.empty(self.vocab, dtype=np.float32)
        _check(lib().b200_forward(self._h, seq, int(token_id), _fp(logits)))
        return logits

    def prefill_token(self, token_id, seq=0):
        _check(lib().b200_prefill_token(self._h, seq, int(token_id)))

    def forward_batch(self, token_ids, seq=0):
        """GpuOnlyInference::forward_batch (src/backend/cuda/gpu_only.rs:776-790): the whole prompt, logits of its last
        token.  Prompts of >= 32 tokens of an eligible dense model run through the tcgen05 dequant-GEMM (fp16 operands)."""
        toks = np.ascontiguousarray(token_ids, dtype=np.uint32)
        logits = np.empty(self.vocab, dtype=np.float32)
        _check(lib().b200_prefill(self._h, seq, toks.ctypes.data_as(C.POINTER(C.c_uint32)), toks.size, _fp(logits)))
        return logits

    prefill = forward_batch

    def reset(self, seq=0):
        _check(lib().b200_reset(self._h, seq))

    def position(self, seq=0):
        p = C.c_uint64(0)
        _check(lib().b200_position(self._h, seq, C.byref(p)))
        return p.value

    # -- beyond the trait --------------------------------------------------
    def decode_batch(self, seqs, tokens):
        seqs = np.ascontiguousarray(seqs, dtype=np.int32)
        toks = np.ascontiguousarray(tokens, dtype=np.uint32)
        logits = np.empty((seqs.size, self.vocab), dtype=np.float32)
        _check(lib().b200_decode_batch(self._h, seqs.ctypes.data_as(C.POINTER(C.c_int)),
                                       toks.ctypes.data_as(C.POINTER(C.c_uint32)), seqs.size, _fp(logits)))
        return logits

    def set_speculation(self, on=True):
        """Greedy continuation behind forward() (b200_ctx_set_speculation): the device picks argmax and starts the next token
        before forward() returns; a caller that feeds that token back finds its logits already on the way."""
        _check(lib().b200_ctx_set_speculation(self._h, 1 if on else 0))

    def speculation_stats(self):
        en, h, m = C.c_int(), C.c_uint64(), C.c_uint64()
        _check(lib().b200_ctx_speculation_stats(self._h, C.byref(en), C.byref(h), C.byref(m)))
        return dict(enabled=bool(en.value), hits=h.value, misses=m.value)

    def debug_set_position(self, pos, seq=0):
        _check(lib().b200_debug_set_position(self._h, seq, int(pos)))

    def decode_batch_greedy(self, seqs, tokens):
        """b200_decode_batch_greedy: the next token of every sequence, picked on the device (last maximum wins)."""
        n = len(seqs)
        s_arr = (C.c_int * n)(*seqs)
        t_arr = (C.c_uint32 * n)(*tokens)
        out = (C.c_uint32 * n)()
        _check(lib().b200_decode_batch_greedy(self._h, s_arr, t_arr, n, out))
        return [int(x) for x in out]

    def decode_greedy(self, first_token, n_steps, seq=0):
        """n_steps device-resident greedy tokens; returns (tokens, elapsed_ms by CUDA events)."""
        out = np.empty(n_steps, dtype=np.uint32)
        ms = C.c_float(0)
        _check(lib().b200_decode_greedy(self._h, seq, int(first_token), n_steps, out.ctypes.data_as(C.POINTER(C.c_uint32)),
                                        C.byref(ms)))
        return out, ms.value

    def hidden(self, layer, seq=0):
        out = np.empty(self.desc["hidden"], dtype=np.float32)
        _check(lib().b200_get_hidden(self._h, seq, layer, _fp(out)))
        return out

    def stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().b200_ctx_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"kernel_launches": a.value, "weight_bytes_per_token": b.value, "kv_bytes_per_pos": c.value}

    def bench_gemv_pass(self, iters=20, seq=0):
        """(avg ms per pass, gemv launches per pass, weight bytes per pass) of the GEMV-only replay."""
        ms, n, nb = C.c_float(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().b200_bench_gemv_pass(self._h, seq, iters, C.byref(ms), C.byref(n), C.byref(nb)))
        return ms.value, n.value, nb.value

    def bench_weight_gemv(self, name, iters=20):
        ms, nbytes = C.c_float(0), C.c_uint64(0)
        _check(lib().b200_bench_weight_gemv(self._h, name.encode(), iters, C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value


class GroupInference:
    """GpuInference over the GPUs of ONE process (b200_group_*): tensor parallel for dense models, expert parallel for MoE models,
    one host thread per device inside the library -- the caller never sees ranks (TensorParallel trait,
    src/backend/tensor_parallel.rs:13-32).  Same surface as GpuOnlyInference: forward() returns the full logits row."""

    @classmethod
    def from_gguf(cls, path, n_devices=1, devices=None, max_seq_len=0, max_batch=1):
        """The group built straight from a GGUF file: every rank's host thread stages its own shard out of the one mapping
        (b200_group_load_gguf)."""
        f = GgufFile(path)
        try:
            return cls(f.model_desc(max_seq_len, max_batch), None, n_devices, devices, gguf=f)
        finally:
            f.close()

    def __init__(self, desc: dict, tensors: dict, n_devices=1, devices=None, feeder=None, gguf=None):
        L = lib()
        if device_count() < n_devices:
            raise NotAvailable(f"cuda-b200: {n_devices} CUDA devices needed, {device_count()} present (no CPU fallback)")
        d = ModelDesc()
        for k, v in desc.items():
            if k not in DESC_KEYS:
                raise InvalidArgument(f"unknown model desc field {k}")
            setattr(d, k, v)
        self.desc = {k: getattr(d, k) for k in DESC_KEYS}
        self.world = n_devices
        self.vocab = self.desc["vocab"]
        devs = (C.c_int * n_devices)(*devices) if devices is not None else None
        h = C.c_void_p()
        _check(L.b200_group_create(C.byref(d), n_devices, devs, C.byref(h)))
        self._h = h
        try:
            for name, (t, ne, data) in (tensors or {}).items():
                self.upload_tensor(name, t, ne, data)
            if feeder is not None:
                feeder(self.upload_tensor)
            if gguf is not None:
                st = LoadStats()
                _check(L.b200_group_load_gguf(self._h, gguf._h, C.byref(st)))
                self.load_stats = st.as_dict()
            _check(L.b200_group_finalize(self._h))
        except Exception:
            self.close()
            raise

    def upload_tensor(self, name, ggml_type, ne, data):
        a = np.ascontiguousarray(data)
        nd = (C.c_uint64 * 4)(*(list(ne) + [1] * (4 - len(ne))))
        _check(lib().b200_group_upload_tensor(self._h, name.encode(), int(ggml_type), nd, len(ne), a.ctypes.data, a.nbytes))

    def close(self):
        if getattr(self, "_h", None):
            lib().b200_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ctx(self, rank=0):
        h = C.c_void_p()
        _check(lib().b200_group_ctx(self._h, rank, C.byref(h)))
        return h

    def path(self):
        v = C.c_int(0)
        _check(lib().b200_ctx_path(self._ctx(0), C.byref(v)))
        return ("graph", "mega", "stream", "stream2")[v.value]

    def forward(self, token_id, seq=0):
        logits = np.empty(self.vocab, dtype=np.float32)
        _check(lib().b200_group_forward(self._h, int(seq), int(token_id), logits.ctypes.data_as(C.POINTER(C.c_float))))
        return logits

    def prefill_token(self, token_id, seq=0):
        _check(lib().b200_group_prefill_token(self._h, int(seq), int(token_id)))

    def reset(self, seq=0):
        _check(lib().b200_group_reset(self._h, int(seq)))

    def position(self, seq=0):
        v = C.c_uint64(0)
        _check(lib().b200_group_position(self._h, int(seq), C.byref(v)))
        return int(v.value)

    def decode_greedy(self, first_token, n_steps, seq=0):
        toks = np.empty(n_steps, dtype=np.uint32)
        ms = C.c_float(0.0)
        _check(lib().b200_group_decode_greedy(self._h, int(seq), int(first_token), int(n_steps),
                                              toks.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(ms)))
        return toks, float(ms.value)


class GpuModelWrapper:
    """GpuModelWrapper<T: GpuInference> (src/backend/mod.rs:302-364): adapts the engine to
    Model::forward(tokens, ctx) — prefill all but the last token, forward the last."""

    def __init__(self, gpu: GpuOnlyInference):
        self.gpu = gpu

    def forward(self, tokens, ctx_position):
        if ctx_position == 0 and self.gpu.position() > 0:
            self.gpu.reset()
        if len(tokens) == 0:
            raise InvalidArgument("No tokens to process")
        for t in tokens[:-1]:
            self.gpu.prefill_token(t)
        return self.gpu.forward(tokens[-1])
