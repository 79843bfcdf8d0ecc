"""GGUF <-> (desc, tensors) for the cuda-b200 host shim.

Reads what ModelLoader::parse_config reads (src/model/loader.rs:62-300) with the same
keys and defaults, and hands tensors over as raw GGUF block bytes, like
GpuOnlyInference::from_model receives them from Tensor::data() (gpu_only.rs:426).
Writing uses the metadata value types the reference reader insists on
(src/gguf/types.rs:78-97: get_u32 accepts only Uint32, get_f32 only Float32).
"""
import numpy as np

NEOX_ARCHS = {"qwen2", "qwen2moe", "qwen3", "qwen3moe", "gptneox", "falcon", "phi2", "phi3", "stablelm"}
SUPPORTED_TYPES = {0, 1, 2, 6, 8, 12, 13, 14}
SUPPORTED_ARCHS = {"llama", "mistral", "mixtral", "qwen2", "qwen2moe"}


def load_gguf(path, max_batch=1):
    import gguf

    r = gguf.GGUFReader(path)

    def field(key):
        f = r.get_field(key)
        return None if f is None else f.contents()

    arch = field("general.architecture")
    if arch is None:
        raise ValueError("missing metadata general.architecture")
    if arch not in SUPPORTED_ARCHS:
        # gemma (GELU), phi / gptneox (LayerNorm, fused QKV) ... share tensor names with llama but not its arithmetic
        # (src/model/loader.rs:145-162, layers.rs uses_gelu): refuse instead of returning wrong logits
        raise ValueError(f"unsupported architecture '{arch}': cuda-b200 implements {sorted(SUPPORTED_ARCHS)} (RMSNorm + SwiGLU + RoPE)")

    def u32(key, default=None):
        v = field(f"{arch}.{key}")
        if v is None:
            if default is None:
                raise ValueError(f"missing metadata {arch}.{key}")
            return default
        return int(v)

    def f32(key, default):
        v = field(f"{arch}.{key}")
        return float(v) if v is not None else default

    hidden = u32("embedding_length")
    n_heads = u32("attention.head_count")
    desc = {
        "hidden": hidden,
        "n_layers": u32("block_count"),
        "n_heads": n_heads,
        "n_kv_heads": u32("attention.head_count_kv", n_heads),
        "head_dim": u32("attention.key_length", hidden // n_heads),
        "ffn": u32("feed_forward_length", hidden * 4 * 2 // 3),
        "max_seq_len": u32("context_length", 2048),
        "norm_eps": f32("attention.layer_norm_rms_epsilon", 1e-5),
        "rope_base": f32("rope.freq_base", 10000.0),
        "rope_scale": f32("rope.scale_linear", 1.0),
        "rope_neox": 1 if arch in NEOX_ARCHS else 0,
        "n_experts": u32("expert_count", 0),
        "n_experts_used": u32("expert_used_count", 0),
        "expert_ffn": u32("expert_feed_forward_length", 0),
        "max_batch": max_batch,
    }
    rd = field(f"{arch}.rope.dimension_count")
    if rd is not None and int(rd) != desc["head_dim"]:
        raise ValueError(f"partial RoPE (rope.dimension_count {int(rd)} != head_dim {desc['head_dim']}) is not implemented")
    tensors = {}
    for t in r.tensors:
        ttype = int(t.tensor_type)
        if ttype not in SUPPORTED_TYPES:
            raise ValueError(f"{t.name}: unsupported ggml type {ttype}")
        ne = [int(x) for x in t.shape]  # GGUF order: ne[0] contiguous
        tensors[t.name] = (ttype, ne, np.ascontiguousarray(t.data).view(np.uint8).ravel())
    v = field(f"{arch}.vocab_size")
    desc["vocab"] = int(v) if v is not None else tensors["token_embd.weight"][1][1]
    desc["tied_output"] = 0 if "output.weight" in tensors else 1
    if desc["n_experts"] > 0 and desc["expert_ffn"] == 0:
        desc["expert_ffn"] = tensors["blk.0.ffn_gate_exps.weight"][1][1]
    return arch, desc, tensors


def write_gguf(path, arch, desc, tensors):
    """Synthetic-model writer (SURVEY.md §A.2 recipe, tests/embedded_model_test.rs:107-304)."""
    import gguf

    w = gguf.GGUFWriter(path, arch)
    w.add_uint32(f"{arch}.embedding_length", desc["hidden"])
    w.add_uint32(f"{arch}.block_count", desc["n_layers"])
    w.add_uint32(f"{arch}.attention.head_count", desc["n_heads"])
    w.add_uint32(f"{arch}.attention.head_count_kv", desc["n_kv_heads"])
    w.add_uint32(f"{arch}.feed_forward_length", desc["ffn"])
    w.add_uint32(f"{arch}.context_length", desc["max_seq_len"])
    w.add_uint32(f"{arch}.vocab_size", desc["vocab"])
    w.add_float32(f"{arch}.attention.layer_norm_rms_epsilon", desc["norm_eps"])
    w.add_float32(f"{arch}.rope.freq_base", desc["rope_base"])
    if desc.get("rope_scale", 1.0) != 1.0:
        w.add_float32(f"{arch}.rope.scale_linear", desc["rope_scale"])
    if desc.get("n_experts", 0) > 0:
        w.add_uint32(f"{arch}.expert_count", desc["n_experts"])
        w.add_uint32(f"{arch}.expert_used_count", desc["n_experts_used"])
    for name, (ttype, ne, data) in tensors.items():
        qt = gguf.GGMLQuantizationType(ttype)
        if ttype == 0:
            arr = np.ascontiguousarray(data).view(np.float32).reshape(list(reversed(ne)))
            w.add_tensor(name, arr, raw_dtype=qt)
        else:
            raw = np.ascontiguousarray(data).view(np.uint8)
            be, bb = gguf.GGML_QUANT_SIZES[qt]
            shape = list(reversed(ne))
            shape[-1] = shape[-1] // be * bb
            w.add_tensor(name, raw.reshape(shape), raw_dtype=qt)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
