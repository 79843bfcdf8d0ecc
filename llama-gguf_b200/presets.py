"""Shapes of the architectures BASELINE.json names and the GGUF tensor
inventory / quant-type mix of a Q4_K_M / Q5_K_M / uniform file for them."""

F32, F16, Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, Q6_K = 0, 1, 2, 6, 8, 12, 13, 14

PRESETS = {
    # Q: Qwen2.5-0.5B (NeoX RoPE, q/k/v biases, tied embeddings)
    "qwen2.5-0.5b": dict(arch="qwen2", hidden=896, n_layers=24, n_heads=14, n_kv_heads=2, head_dim=64, ffn=4864,
                         vocab=151936, norm_eps=1e-6, rope_base=1e6, rope_neox=1, bias=True, tied=True),
    # T: TinyLlama-1.1B
    "tinyllama-1.1b": dict(arch="llama", hidden=2048, n_layers=22, n_heads=32, n_kv_heads=4, head_dim=64, ffn=5632,
                           vocab=32000, norm_eps=1e-5, rope_base=1e4, rope_neox=0, bias=False, tied=False),
    # 8B: Llama-3-8B (src/model/config.rs:239-281)
    "llama-3-8b": dict(arch="llama", hidden=4096, n_layers=32, n_heads=32, n_kv_heads=8, head_dim=128, ffn=14336,
                       vocab=128256, norm_eps=1e-5, rope_base=5e5, rope_neox=0, bias=False, tied=False),
    "llama-3-70b": dict(arch="llama", hidden=8192, n_layers=80, n_heads=64, n_kv_heads=8, head_dim=128, ffn=28672,
                        vocab=128256, norm_eps=1e-5, rope_base=5e5, rope_neox=0, bias=False, tied=False),
    # MX: Mixtral-8x7B (arch string "llama" + llama.expert_count)
    "mixtral-8x7b": dict(arch="llama", hidden=4096, n_layers=32, n_heads=32, n_kv_heads=8, head_dim=128, ffn=14336,
                         vocab=32000, norm_eps=1e-5, rope_base=1e6, rope_neox=0, bias=False, tied=False,
                         n_experts=8, n_experts_used=2),
}


def make_desc(p, max_seq_len, max_batch=1):
    return {
        "hidden": p["hidden"], "n_layers": p["n_layers"], "n_heads": p["n_heads"], "n_kv_heads": p["n_kv_heads"],
        "head_dim": p["head_dim"], "ffn": p["ffn"], "vocab": p["vocab"], "max_seq_len": max_seq_len,
        "norm_eps": p["norm_eps"], "rope_base": p["rope_base"], "rope_scale": 1.0, "rope_neox": p["rope_neox"],
        "n_experts": p.get("n_experts", 0), "n_experts_used": p.get("n_experts_used", 0),
        "expert_ffn": p["ffn"] if p.get("n_experts", 0) else 0, "tied_output": 1 if p.get("tied") else 0,
        "max_batch": max_batch,
    }


def use_more_bits(i, n_layers):
    """llama.cpp's rule for which layers get the wider type for attn_v / ffn_down in *_K_M files."""
    return i < n_layers // 8 or i >= 7 * n_layers // 8 or (i - n_layers // 8) % 3 == 2


def _fit(t, k):
    """K-quants need ne[0] % 256 == 0; real files fall back Q4_K/Q5_K -> Q5_0 and Q6_K -> Q8_0."""
    if t in (Q4_K, Q5_K, Q6_K) and k % 256:
        return Q8_0 if t == Q6_K else Q5_0
    return t


def tensor_plan(p, mix):
    """[(gguf_name, ggml_type, ne)] for a model of preset `p`.

    mix: "Q4_K_M" | "Q5_K_M" (base type + Q6_K for output and the use_more_bits layers)
         or a uniform type name "Q8_0" | "Q6_K" | "Q4_K" | "Q5_K" | "Q4_0" | "Q5_0" | "F16" | "F32".
    """
    names = {"F32": F32, "F16": F16, "Q4_0": Q4_0, "Q5_0": Q5_0, "Q8_0": Q8_0, "Q4_K": Q4_K, "Q5_K": Q5_K, "Q6_K": Q6_K}
    if mix in ("Q4_K_M", "Q5_K_M"):
        base = Q4_K if mix == "Q4_K_M" else Q5_K
        wide = lambda i: Q6_K if use_more_bits(i, p["n_layers"]) else base
        out_t = Q6_K
    else:
        base = names[mix]
        wide = lambda i: base
        out_t = base
    H, hd, nh, nkv, I, V = p["hidden"], p["head_dim"], p["n_heads"], p["n_kv_heads"], p["ffn"], p["vocab"]
    E = p.get("n_experts", 0)
    plan = [("token_embd.weight", _fit(base, H), [H, V]), ("output_norm.weight", F32, [H])]
    if not p.get("tied"):
        plan.append(("output.weight", _fit(out_t, H), [H, V]))
    for i in range(p["n_layers"]):
        b = f"blk.{i}."
        plan += [(b + "attn_norm.weight", F32, [H]), (b + "ffn_norm.weight", F32, [H]),
                 (b + "attn_q.weight", _fit(base, H), [H, nh * hd]),
                 (b + "attn_k.weight", _fit(base, H), [H, nkv * hd]),
                 (b + "attn_v.weight", _fit(wide(i), H), [H, nkv * hd]),
                 (b + "attn_output.weight", _fit(base, nh * hd), [nh * hd, H])]
        if p.get("bias"):
            plan += [(b + "attn_q.bias", F32, [nh * hd]), (b + "attn_k.bias", F32, [nkv * hd]),
                     (b + "attn_v.bias", F32, [nkv * hd])]
        if E:
            plan += [(b + "ffn_gate_inp.weight", F32, [H, E]),
                     (b + "ffn_gate_exps.weight", _fit(base, H), [H, I, E]),
                     (b + "ffn_up_exps.weight", _fit(base, H), [H, I, E]),
                     (b + "ffn_down_exps.weight", _fit(wide(i), I), [I, H, E])]
        else:
            plan += [(b + "ffn_gate.weight", _fit(base, H), [H, I]), (b + "ffn_up.weight", _fit(base, H), [H, I]),
                     (b + "ffn_down.weight", _fit(wide(i), I), [I, H])]
    return plan
