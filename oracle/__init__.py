"""ctypes binding of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(llama-gguf_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

# ggml type ids (reference: src/gguf/constants.rs:56-89)
F32, F16, Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, Q6_K = 0, 1, 2, 6, 8, 12, 13, 14
TYPE_NAMES = {F32: "F32", F16: "F16", Q4_0: "Q4_0", Q5_0: "Q5_0", Q8_0: "Q8_0", Q4_K: "Q4_K", Q5_K: "Q5_K", Q6_K: "Q6_K"}


def build(force=False):
    """Compile liboracle.so with oracle/Makefile (g++ -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class OrcDesc(C.Structure):
    """Field-for-field the same as b200_model_desc in include/llama_b200.h."""

    _fields_ = [
        ("hidden", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("ffn", C.c_int32), ("vocab", C.c_int32), ("max_seq_len", C.c_int32),
        ("norm_eps", C.c_float), ("rope_base", C.c_float), ("rope_scale", C.c_float),
        ("rope_neox", C.c_int32), ("n_experts", C.c_int32), ("n_experts_used", C.c_int32),
        ("expert_ffn", C.c_int32), ("tied_output", C.c_int32), ("max_batch", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.orc_f16_to_f32.restype = C.c_float
        L.orc_f16_to_f32.argtypes = [C.c_uint16]
        L.orc_f32_to_f16.restype = C.c_uint16
        L.orc_f32_to_f16.argtypes = [C.c_float]
        L.orc_dequantize.argtypes = [C.c_int, C.c_void_p, C.c_int64, fp]
        L.orc_quantize.argtypes = [C.c_int, fp, C.c_int64, C.c_void_p]
        L.orc_dot_f32.restype = C.c_float
        L.orc_dot_f32.argtypes = [fp, fp, C.c_int64]
        L.orc_dot_q.restype = C.c_float
        L.orc_dot_q.argtypes = [C.c_int, C.c_void_p, fp, C.c_int64]
        L.orc_vec_mat_q.argtypes = [C.c_int, C.c_void_p, fp, fp, C.c_int64, C.c_int64]
        L.orc_softmax.argtypes = [fp, fp, C.c_int64]
        L.orc_rms_norm.argtypes = [fp, fp, C.c_float, fp, C.c_int64, C.c_int64]
        for n in ("orc_silu",):
            getattr(L, n).argtypes = [fp, fp, C.c_int64]
        L.orc_silu_mul_inplace.argtypes = [fp, fp, C.c_int64]
        for n in ("orc_add", "orc_mul"):
            getattr(L, n).argtypes = [fp, fp, fp, C.c_int64]
        L.orc_scale.argtypes = [fp, C.c_float, fp, C.c_int64]
        L.orc_rope.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orc_attention_cached.argtypes = [fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
        L.orc_moe_route.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), fp]
        L.orc_argmax_last.argtypes = [fp, C.c_int64]
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.POINTER(OrcDesc)]
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_model_set_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_int64]
        L.orc_model_reset.argtypes = [C.c_void_p]
        L.orc_model_position.argtypes = [C.c_void_p]
        L.orc_model_set_faithful_embedding.argtypes = [C.c_void_p, C.c_int]
        L.orc_model_set_kv_format.argtypes = [C.c_void_p, C.c_int]
        L.orc_kv_quantize_int8.restype = C.c_float
        L.orc_kv_quantize_int8.argtypes = [fp, C.c_int, C.c_void_p]
        L.orc_kv_dequantize_int8.argtypes = [C.c_void_p, C.c_int, C.c_float, fp]
        L.orc_model_get_hidden.argtypes = [C.c_void_p, C.c_int, fp]
        L.orc_model_get_kv.argtypes = [C.c_void_p, C.c_int, C.c_int, fp]
        L.orc_model_forward.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_int, fp]
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def block_elems(t):
    return lib().orc_block_elems(t)


def block_bytes(t):
    return lib().orc_block_bytes(t)


def nbytes_for(t, n_elems):
    return n_elems // block_elems(t) * block_bytes(t)


def dequantize(t, raw, n_elems):
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.empty(n_elems, dtype=np.float32)
    rc = lib().orc_dequantize(t, raw.ctypes.data, n_elems, _fp(out))
    if rc:
        raise ValueError("orc_dequantize failed")
    return out


def quantize(t, x):
    x = _f32(x).ravel()
    out = np.empty(nbytes_for(t, x.size), dtype=np.uint8)
    rc = lib().orc_quantize(t, _fp(x), x.size, out.ctypes.data)
    if rc:
        raise ValueError("orc_quantize failed")
    return out


def kv_quantize_int8(x):
    """quantize_int8 (src/model/kv_quantized.rs:366-387): (int8 ndarray, scale)."""
    x = _f32(x).ravel()
    q = np.empty(x.size, dtype=np.int8)
    s = lib().orc_kv_quantize_int8(_fp(x), x.size, q.ctypes.data)
    return q, float(s)


def kv_dequantize_int8(q, scale):
    """dequantize_int8 (src/model/kv_quantized.rs:390-392)."""
    q = np.ascontiguousarray(q, dtype=np.int8)
    out = np.empty(q.size, dtype=np.float32)
    lib().orc_kv_dequantize_int8(q.ctypes.data, q.size, scale, _fp(out))
    return out


def dot_q(t, row, x):
    row = np.ascontiguousarray(row, dtype=np.uint8)
    x = _f32(x)
    return float(lib().orc_dot_q(t, row.ctypes.data, _fp(x), x.size))


def dot_f32(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_dot_f32(_fp(a), _fp(b), a.size))


def vec_mat_q(t, w, x, n):
    """y[j] = sum_k x[k] * deq(W)[j, k]; W is n rows of k/bs blocks (GGUF [k, n])."""
    w = np.ascontiguousarray(w)
    x = _f32(x)
    out = np.empty(n, dtype=np.float32)
    rc = lib().orc_vec_mat_q(t, w.ctypes.data, _fp(x), _fp(out), x.size, n)
    if rc:
        raise ValueError("orc_vec_mat_q failed")
    return out


def rms_norm(x, w, eps):
    x, w = _f32(x), _f32(w)
    out = np.empty_like(x)
    hidden = x.shape[-1]
    lib().orc_rms_norm(_fp(x), _fp(w), eps, _fp(out), x.size // hidden, hidden)
    return out


def silu(x):
    x = _f32(x)
    out = np.empty_like(x)
    lib().orc_silu(_fp(x), _fp(out), x.size)
    return out


def silu_mul(gate, up):
    g = _f32(gate).copy()
    u = _f32(up)
    lib().orc_silu_mul_inplace(_fp(g), _fp(u), g.size)
    return g


def add(a, b):
    a, b = _f32(a), _f32(b)
    o = np.empty_like(a)
    lib().orc_add(_fp(a), _fp(b), _fp(o), a.size)
    return o


def mul(a, b):
    a, b = _f32(a), _f32(b)
    o = np.empty_like(a)
    lib().orc_mul(_fp(a), _fp(b), _fp(o), a.size)
    return o


def scale(a, s):
    a = _f32(a)
    o = np.empty_like(a)
    lib().orc_scale(_fp(a), s, _fp(o), a.size)
    return o


def softmax(x):
    x = _f32(x)
    o = np.empty_like(x)
    lib().orc_softmax(_fp(x), _fp(o), x.size)
    return o


def rope(q, k, pos, base, scale_, neox):
    """q [n_heads, 1, hd], k [n_kv, 1, hd] -> rotated copies."""
    q, k = _f32(q).copy(), _f32(k).copy()
    hd = q.shape[-1]
    lib().orc_rope(_fp(q), _fp(k), q.size // hd, k.size // hd, hd, pos, base, scale_, int(neox))
    return q, k


def attention_cached(q, kc, vc, scale_, kv_len):
    """q [nh, 1, hd]; kc/vc [nkv, max_seq, hd] -> out [nh, 1, hd]."""
    q, kc, vc = _f32(q), _f32(kc), _f32(vc)
    nh, hd = q.shape[0], q.shape[-1]
    nkv, max_seq = kc.shape[0], kc.shape[1]
    out = np.empty((nh, 1, hd), dtype=np.float32)
    lib().orc_attention_cached(_fp(q), _fp(kc), _fp(vc), _fp(out), nh, nkv, hd, max_seq, scale_, kv_len)
    return out


def moe_route(h, w_router, n_experts, top_k):
    h, w_router = _f32(h), _f32(w_router)
    idx = np.empty(top_k, dtype=np.int32)
    wts = np.empty(top_k, dtype=np.float32)
    lib().orc_moe_route(_fp(h), _fp(w_router), h.size, n_experts, top_k, idx.ctypes.data_as(C.POINTER(C.c_int)), _fp(wts))
    return idx, wts


def argmax_last(v):
    v = _f32(v)
    return int(lib().orc_argmax_last(_fp(v), v.size))


def set_num_threads(n):
    lib().orc_set_num_threads(n)


def num_threads():
    return lib().orc_num_threads()


class OracleModel:
    """CPU oracle of LlamaModel::forward (reference: src/model/llama.rs:275-362).

    `tensors` is {gguf_name: (ggml_type, ne tuple, uint8/float32 ndarray)}.
    """

    def __init__(self, desc: dict, tensors: dict, kv_format="f32"):
        d = OrcDesc()
        for k, v in desc.items():
            setattr(d, k, v)
        self.desc = dict(desc)
        self._h = lib().orc_model_create(C.byref(d))
        if kv_format == "int8":   # QuantizedKVCache Int8 (src/model/kv_quantized.rs)
            lib().orc_model_set_kv_format(self._h, 1)
        for name, (t, ne, data) in tensors.items():
            a = np.ascontiguousarray(data)
            nd = (C.c_int64 * 4)(*(list(ne) + [1] * (4 - len(ne))))
            lib().orc_model_set_tensor(self._h, name.encode(), t, nd, len(ne), a.ctypes.data, a.nbytes)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_model_destroy(self._h)
            self._h = None

    def forward(self, tokens, want_logits=True):
        toks = np.ascontiguousarray(tokens, dtype=np.uint32)
        logits = np.empty(self.desc["vocab"], dtype=np.float32) if want_logits else None
        rc = lib().orc_model_forward(self._h, toks.ctypes.data_as(C.POINTER(C.c_uint32)), toks.size,
                                     _fp(logits) if want_logits else None)
        if rc:
            raise RuntimeError(f"orc_model_forward failed: {rc}")
        return logits

    def reset(self):
        lib().orc_model_reset(self._h)

    def position(self):
        return lib().orc_model_position(self._h)

    def hidden(self, layer):
        """Hidden state of the last processed token after `layer` layers (0 = embedding)."""
        out = np.empty(self.desc["hidden"], dtype=np.float32)
        lib().orc_model_get_hidden(self._h, layer, _fp(out))
        return out

    def kv(self, layer, which):
        d = self.desc
        hd = d.get("head_dim") or d["hidden"] // d["n_heads"]
        out = np.empty((d["n_kv_heads"], d["max_seq_len"], hd), dtype=np.float32)
        lib().orc_model_get_kv(self._h, layer, which, _fp(out))
        return out

    def set_faithful_embedding(self, on):
        lib().orc_model_set_faithful_embedding(self._h, int(on))
