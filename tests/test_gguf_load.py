"""GGUF -> HBM direct load: the library's own GGUF reader and model-description rules (csrc/gguf_load.cuh,
b200_gguf_* / b200_ctx_create_from_gguf) against

  * gguf-py (the format's reference implementation) on synthetic files of every model family: same header fields, same tensor
    names / types / shapes / byte ranges, bit-identical tensor bytes, same model description as the Python shim
    (gguf_io.load_gguf, itself a mirror of ModelLoader::parse_config, src/model/loader.rs:62-300);
  * the reference reader's own tests: invalid magic and unsupported version (src/gguf/reader.rs:360-378);
  * hand-packed version 1 / 2 / 3 files (32-bit counts and lengths in v1, general.alignment, reader.rs:51-96, 228-330);
  * on the GPU: a context loaded straight from the file gives bit-identical logits to the same model uploaded tensor by tensor,
    within 1e-3 of the oracle, through every staging variant (pinned double buffer with several chunks per tensor, unstaged).
"""
import os
import struct

import numpy as np
import pytest

import synth
from synth import rel_err

TOL = 1e-3
FAMILIES = [("llama-tiny", "Q4_K_M"), ("qwen-tiny", "Q4_K_M"), ("tinyllama-tiny", "Q8_0"), ("mixtral-tiny", "Q4_K_M")]


def _write(tmp_path, preset, mix, ctx=32):
    from llama_gguf_b200 import gguf_io

    arch, desc, tensors = synth.synth_model(preset, mix, ctx)
    path = os.path.join(tmp_path, f"{preset}-{mix}.gguf")
    gguf_io.write_gguf(path, arch, desc, tensors)
    return path, arch, desc, tensors


@pytest.mark.parametrize("preset,mix", FAMILIES)
def test_native_reader_matches_gguf_py(b200, tmp_path, preset, mix):
    from llama_gguf_b200 import gguf_io
    import gguf

    path, arch, desc, tensors = _write(tmp_path, preset, mix)
    arch2, desc2, tensors2 = gguf_io.load_gguf(path, max_batch=3)
    r = gguf.GGUFReader(path)
    with b200.GgufFile(path) as f:
        info = f.info()
        assert info["version"] == 3 and info["n_tensors"] == len(r.tensors) == len(tensors)
        assert info["alignment"] == r.alignment and info["data_offset"] == r.data_offset
        assert info["file_bytes"] == os.path.getsize(path)
        assert f.architecture() == arch2
        assert f.model_desc(0, 3) == {k: (pytest.approx(v) if isinstance(v, float) else v) for k, v in desc2.items()}
        assert f.model_desc(16, 1)["max_seq_len"] == 16 and f.model_desc(10 ** 6, 1)["max_seq_len"] == desc2["max_seq_len"]
        for i, t in enumerate(r.tensors):
            name, ttype, ne, nbytes, data = f.tensor(i, with_data=True)
            assert (name, ttype, ne) == (t.name, int(t.tensor_type), [int(x) for x in t.shape])
            want = np.ascontiguousarray(t.data).view(np.uint8).ravel()
            assert nbytes == want.size and np.array_equal(data, want)
            assert np.array_equal(data, np.ascontiguousarray(tensors2[name][2]).view(np.uint8).ravel())


def test_reference_reader_kats(b200, tmp_path):
    """src/gguf/reader.rs:360-378: magic 0 -> InvalidMagic, version 99 -> UnsupportedVersion."""
    p = os.path.join(tmp_path, "bad_magic.gguf")
    open(p, "wb").write(struct.pack("<II", 0, 3) + b"\0" * 16)
    with pytest.raises(b200.InvalidArgument, match="magic"):
        b200.GgufFile(p)
    p = os.path.join(tmp_path, "bad_version.gguf")
    open(p, "wb").write(struct.pack("<II", 0x46554747, 99) + b"\0" * 16)
    with pytest.raises(b200.Unsupported, match="version 99"):
        b200.GgufFile(p)
    with pytest.raises(b200.InvalidArgument):
        b200.GgufFile(os.path.join(tmp_path, "does_not_exist.gguf"))
    p = os.path.join(tmp_path, "short.gguf")
    open(p, "wb").write(b"GGU")
    with pytest.raises(b200.InvalidArgument):
        b200.GgufFile(p)


def _s(version, text):
    b = text.encode()
    return struct.pack("<I" if version == 1 else "<Q", len(b)) + b


def _pack_gguf(version, kvs, tensors, alignment=None):
    """Minimal writer for v1 (32-bit counts, lengths and dims, reader.rs:51-60, 230, 319) and v2 / v3 (64-bit)."""
    cnt = "<I" if version == 1 else "<Q"
    kvs = list(kvs)
    if alignment is not None:
        kvs.append(("general.alignment", 4, alignment))
    out = struct.pack("<II", 0x46554747, version) + struct.pack(cnt, len(tensors)) + struct.pack(cnt, len(kvs))
    for key, vt, val in kvs:
        out += _s(version, key) + struct.pack("<I", vt)
        if vt == 4:
            out += struct.pack("<I", val)
        elif vt == 6:
            out += struct.pack("<f", val)
        elif vt == 10:
            out += struct.pack("<Q", val)
        elif vt == 8:
            out += _s(version, val)
        elif vt == 9:   # array of strings
            out += struct.pack("<I", 8) + struct.pack(cnt, len(val)) + b"".join(_s(version, x) for x in val)
        elif vt == 7:
            out += struct.pack("<B", val)
        else:
            raise ValueError(vt)
    al = alignment or 32
    blobs, off = [], 0
    for name, ne, arr in tensors:
        out += _s(version, name) + struct.pack("<I", len(ne)) + b"".join(struct.pack(cnt, d) for d in ne)
        out += struct.pack("<I", 0) + struct.pack("<Q", off)   # F32
        raw = np.ascontiguousarray(arr, dtype=np.float32).tobytes()
        blobs.append((off, raw))
        off = (off + len(raw) + al - 1) // al * al
    data_offset = (len(out) + al - 1) // al * al
    out += b"\0" * (data_offset - len(out))
    for o, raw in blobs:
        out += b"\0" * (o - (len(out) - data_offset)) + raw
    return out, data_offset


@pytest.mark.parametrize("version,alignment", [(1, None), (2, None), (3, None), (3, 64), (2, 256)])
def test_versions_and_alignment(b200, tmp_path, version, alignment):
    a = np.arange(64, dtype=np.float32)
    bmat = np.arange(32 * 3, dtype=np.float32) * 0.5
    kvs = [("general.architecture", 8, "llama"), ("llama.embedding_length", 4, 64), ("llama.block_count", 4, 2),
           ("llama.attention.head_count", 4, 4), ("tokenizer.ggml.tokens", 9, ["a", "bc", "def", "", "g"]), ("some.flag", 7, 1),
           ("llama.rope.freq_base", 6, 500000.0)]
    blob, data_offset = _pack_gguf(version, kvs, [("output_norm.weight", [64], a), ("x.weight", [32, 3], bmat)], alignment)
    p = os.path.join(tmp_path, f"v{version}.gguf")
    open(p, "wb").write(blob)
    with b200.GgufFile(p) as f:
        info = f.info()
        assert info["version"] == version and info["n_tensors"] == 2 and info["alignment"] == (alignment or 32)
        assert info["data_offset"] == data_offset and info["n_metadata"] == len(kvs) + (alignment is not None)
        n0, t0, ne0, nb0, d0 = f.tensor(0, with_data=True)
        n1, t1, ne1, nb1, d1 = f.tensor(1, with_data=True)
        assert (n0, t0, ne0, nb0) == ("output_norm.weight", 0, [64], 256) and np.array_equal(d0.view(np.float32), a)
        assert (n1, t1, ne1, nb1) == ("x.weight", 0, [32, 3], 384) and np.array_equal(d1.view(np.float32), bmat)
        d = f.model_desc()
        # defaults of parse_config: kv heads = heads, head_dim = hidden / heads, ffn = hidden * 8 / 3, ctx 2048, eps 1e-5, scale 1,
        # vocab from the token array's length (loader.rs:77-137), tied head when there is no output.weight
        assert (d["hidden"], d["n_layers"], d["n_heads"], d["n_kv_heads"], d["head_dim"]) == (64, 2, 4, 4, 16)
        assert d["ffn"] == 64 * 4 * 2 // 3 and d["max_seq_len"] == 2048 and d["vocab"] == 5 and d["tied_output"] == 1
        assert d["norm_eps"] == pytest.approx(1e-5) and d["rope_base"] == pytest.approx(500000.0) and d["rope_scale"] == 1.0
        assert d["rope_neox"] == 0 and d["n_experts"] == 0


def test_description_errors_and_strict_getters(b200, tmp_path):
    def make(kvs, name):
        blob, _ = _pack_gguf(3, kvs, [("output_norm.weight", [64], np.zeros(64))])
        p = os.path.join(tmp_path, name)
        open(p, "wb").write(blob)
        return p

    base = [("llama.embedding_length", 4, 64), ("llama.block_count", 4, 2), ("llama.attention.head_count", 4, 4)]
    with b200.GgufFile(make(base, "noarch.gguf")) as f:
        with pytest.raises(b200.InvalidArgument, match="general.architecture"):
            f.model_desc()
    gem = [("general.architecture", 8, "gemma")] + [(k.replace("llama.", "gemma."), t, v) for k, t, v in base]
    with b200.GgufFile(make(gem, "gemma.gguf")) as f:
        assert f.architecture() == "gemma"
        with pytest.raises(b200.Unsupported, match="gemma"):     # GELU arch: same tensor names, different arithmetic
            f.model_desc()
    with b200.GgufFile(make([("general.architecture", 8, "llama")] + base[1:], "nohidden.gguf")) as f:
        with pytest.raises(b200.InvalidArgument, match="embedding_length"):
            f.model_desc()
    # get_u32 accepts Uint32 only (gguf/types.rs:78-83): a Uint64 embedding_length reads as absent, like in the reference
    wide = [("general.architecture", 8, "llama"), ("llama.embedding_length", 10, 64)] + base[1:]
    with b200.GgufFile(make(wide, "u64.gguf")) as f:
        with pytest.raises(b200.InvalidArgument, match="embedding_length"):
            f.model_desc()
    partial = [("general.architecture", 8, "llama")] + base + [("llama.rope.dimension_count", 4, 8)]
    with b200.GgufFile(make(partial, "partial_rope.gguf")) as f:
        with pytest.raises(b200.Unsupported, match="partial RoPE"):
            f.model_desc()
    qw = [("general.architecture", 8, "qwen2")] + [(k.replace("llama.", "qwen2."), t, v) for k, t, v in base]
    with b200.GgufFile(make(qw, "qwen.gguf")) as f:
        assert f.model_desc()["rope_neox"] == 1                   # loader.rs:145-162


def test_parse_config_fallbacks(b200, tmp_path):
    """The fallback chains of ModelLoader::parse_config (loader.rs:77-137, 163-175, 349-355), one file per branch."""
    def desc_of(kvs, tensors, name):
        blob, _ = _pack_gguf(3, [("general.architecture", 8, "llama")] + kvs, tensors)
        p = os.path.join(tmp_path, name)
        open(p, "wb").write(blob)
        with b200.GgufFile(p) as f:
            return f.model_desc()

    core = [("llama.embedding_length", 4, 128), ("llama.block_count", 4, 1), ("llama.attention.head_count", 4, 4)]
    norm = ("output_norm.weight", [128], np.zeros(128))
    emb = ("token_embd.weight", [128, 77], np.zeros(128 * 77))
    # vocab: {arch}.vocab_size wins, then tokenizer.ggml.vocab_size, then the embedding's rows, then 32000
    assert desc_of(core + [("llama.vocab_size", 4, 500), ("tokenizer.ggml.vocab_size", 4, 400)], [norm, emb], "v1.gguf")["vocab"] == 500
    assert desc_of(core + [("tokenizer.ggml.vocab_size", 4, 400)], [norm, emb], "v2.gguf")["vocab"] == 400
    assert desc_of(core, [norm, emb], "v3.gguf")["vocab"] == 77
    assert desc_of(core, [norm], "v4.gguf")["vocab"] == 32000
    # eps: layer_norm_rms_epsilon, else layer_norm_epsilon, else 1e-5
    assert desc_of(core + [("llama.attention.layer_norm_epsilon", 6, 3e-6)], [norm], "e1.gguf")["norm_eps"] == pytest.approx(3e-6)
    assert desc_of(core + [("llama.attention.layer_norm_rms_epsilon", 6, 2e-6), ("llama.attention.layer_norm_epsilon", 6, 3e-6)], [norm],
                   "e2.gguf")["norm_eps"] == pytest.approx(2e-6)
    # heads / head_dim / ffn / context / rope
    d = desc_of(core + [("llama.attention.head_count_kv", 4, 2), ("llama.attention.key_length", 4, 64), ("llama.feed_forward_length", 4, 300),
                        ("llama.context_length", 4, 4096), ("llama.rope.scale_linear", 6, 4.0), ("llama.rope.freq_base", 6, 1e6),
                        ("llama.rope.dimension_count", 4, 64)], [norm], "h.gguf")
    assert (d["n_kv_heads"], d["head_dim"], d["ffn"], d["max_seq_len"], d["rope_scale"], d["rope_base"]) == (2, 64, 300, 4096, 4.0, 1e6)
    # MoE: counts from the metadata, the expert width from the metadata or from blk.0.ffn_gate_exps.weight ([in, out, experts])
    exps = ("blk.0.ffn_gate_exps.weight", [128, 96, 8], np.zeros(128 * 96 * 8))
    d = desc_of(core + [("llama.expert_count", 4, 8), ("llama.expert_used_count", 4, 2)], [norm, exps], "m1.gguf")
    assert (d["n_experts"], d["n_experts_used"], d["expert_ffn"]) == (8, 2, 96)
    d = desc_of(core + [("llama.expert_count", 4, 8), ("llama.expert_used_count", 4, 2), ("llama.expert_feed_forward_length", 4, 80)],
                [norm, exps], "m2.gguf")
    assert d["expert_ffn"] == 80
    # tied head iff there is no output.weight
    assert desc_of(core, [norm, emb], "t1.gguf")["tied_output"] == 1
    assert desc_of(core, [norm, emb, ("output.weight", [128, 77], np.zeros(128 * 77))], "t2.gguf")["tied_output"] == 0


def test_truncated_files(b200, tmp_path):
    path, arch, desc, tensors = _write(tmp_path, "tinyllama-tiny", "Q8_0")
    blob = open(path, "rb").read()
    with b200.GgufFile(path) as f:
        data_offset = f.info()["data_offset"]
        n = f.info()["n_tensors"]
    for cut, name in ((data_offset // 2, "cut_meta.gguf"), (40, "cut_header.gguf")):
        p = os.path.join(tmp_path, name)
        open(p, "wb").write(blob[:cut])
        with pytest.raises(b200.InvalidArgument, match="truncated"):
            b200.GgufFile(p)
    p = os.path.join(tmp_path, "cut_data.gguf")          # infos intact, data short: GgufFile::tensor_data -> None (mod.rs:34-42)
    open(p, "wb").write(blob[: data_offset + (len(blob) - data_offset) // 2])
    with b200.GgufFile(p) as f:
        f.tensor(0, with_data=True)
        with pytest.raises(b200.InvalidArgument, match="outside the file"):
            for i in range(n):
                f.tensor(i, with_data=True)


def test_corrupt_shapes_and_offsets_are_refused(b200, tmp_path):
    """Dimensions whose product overflows 64 bits, a shape larger than the file, an offset beyond it: errors, never a wild pointer."""
    a = np.arange(64, dtype=np.float32)
    kvs = [("general.architecture", 8, "llama")]
    blob, data_offset = _pack_gguf(3, kvs, [("output_norm.weight", [64], a)])
    # tensor info layout (v3): name (8 + 18) | n_dims u32 | ne[0] u64 | type u32 | offset u64 -- patch ne[0] / the offset in place
    info_at = blob.index(b"output_norm.weight") + len("output_norm.weight")
    def patched(ne0=None, offset=None):
        b = bytearray(blob)
        if ne0 is not None:
            b[info_at + 4: info_at + 12] = struct.pack("<Q", ne0)
        if offset is not None:
            b[info_at + 16: info_at + 24] = struct.pack("<Q", offset)
        return bytes(b)
    p = os.path.join(tmp_path, "huge_dim.gguf")
    open(p, "wb").write(patched(ne0=2 ** 63))
    with pytest.raises(b200.InvalidArgument, match="overflows"):
        b200.GgufFile(p)
    p = os.path.join(tmp_path, "big_dim.gguf")          # 4 GB of f32 in a 1 KB file: parses, but has no data
    open(p, "wb").write(patched(ne0=10 ** 9))
    with b200.GgufFile(p) as f:
        assert f.tensor(0)[3] == 4 * 10 ** 9
        with pytest.raises(b200.InvalidArgument, match="outside the file"):
            f.tensor(0, with_data=True)
    for name, off in (("far_offset.gguf", 2 ** 40), ("wrap_offset.gguf", 2 ** 64 - data_offset)):
        p = os.path.join(tmp_path, name)
        open(p, "wb").write(patched(offset=off))
        with b200.GgufFile(p) as f:
            assert f.tensor(0)[3] == 256
            with pytest.raises(b200.InvalidArgument, match="outside the file"):
                f.tensor(0, with_data=True)


# ------------------------------------------------------------------ GPU: the load itself
@pytest.mark.gpu
@pytest.mark.parametrize("preset,mix", FAMILIES)
def test_from_gguf_equals_the_tensor_by_tensor_upload_and_the_oracle(b200, oracle, tmp_path, preset, mix):
    path, arch, desc, tensors = _write(tmp_path, preset, mix)
    toks = synth.prompt_tokens(6, desc["vocab"])
    mem = b200.GpuOnlyInference(desc, tensors)
    want_gpu = mem.forward_batch(toks)
    mem.close()
    want = oracle.OracleModel(desc, tensors).forward(toks)
    gpu = b200.GpuOnlyInference.from_gguf(path, max_seq_len=32)
    st = gpu.load_stats
    assert st["tensors_loaded"] == len(tensors) and st["tensors_skipped"] == 0
    assert st["tensor_bytes"] == st["device_bytes"] == sum(np.ascontiguousarray(t[2]).nbytes for t in tensors.values())
    assert st["file_bytes"] == os.path.getsize(path) and st["seconds"] > 0
    got = gpu.forward_batch(toks)
    assert np.array_equal(got, want_gpu), "same bytes in HBM -> same logits, bit for bit"
    assert rel_err(got, want) < TOL
    gpu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"B200_LOAD_CHUNK_KB": "64", "B200_LOAD_THREADS": "3"}, {"B200_LOAD_CHUNK_KB": "4"},
                                 {"B200_LOAD_STAGED": "0"}])
def test_staging_variants_load_the_same_bytes(b200, tmp_path, env):
    """Chunks far smaller than a tensor (many trips round the double buffer), odd thread counts, and the unstaged copy."""
    path, arch, desc, tensors = _write(tmp_path, "llama-tiny", "Q4_K_M")
    toks = synth.prompt_tokens(5, desc["vocab"])
    base = b200.GpuOnlyInference.from_gguf(path, max_seq_len=32)
    want = base.forward_batch(toks)
    base.close()
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        gpu = b200.GpuOnlyInference.from_gguf(path, max_seq_len=32)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert np.array_equal(gpu.forward_batch(toks), want)
    gpu.close()


@pytest.mark.gpu
def test_from_gguf_int8_kv_and_errors(b200, oracle, tmp_path):
    path, arch, desc, tensors = _write(tmp_path, "llama-tiny", "Q4_K_M")
    toks = synth.prompt_tokens(12, desc["vocab"])
    gpu = b200.GpuOnlyInference.from_gguf(path, max_seq_len=32, kv_format="int8")
    assert gpu.kv_format() == "int8"
    assert rel_err(gpu.prefill(toks), oracle.OracleModel(desc, tensors, kv_format="int8").forward(toks)) < TOL
    gpu.close()
    # a file without one of the model's tensors: finalize reports it, nothing leaks through
    from llama_gguf_b200 import gguf_io

    bad = dict(tensors)
    del bad["blk.1.ffn_down.weight"]
    p = os.path.join(tmp_path, "missing.gguf")
    gguf_io.write_gguf(p, arch, desc, bad)
    with pytest.raises(b200.InvalidArgument, match="ffn_down"):
        b200.GpuOnlyInference.from_gguf(p)
    # an extra tensor the engine has no slot for is skipped and counted
    extra = dict(tensors)
    extra["rope_freqs.weight"] = (0, [64], np.ones(64, dtype=np.float32))
    p = os.path.join(tmp_path, "extra.gguf")
    gguf_io.write_gguf(p, arch, desc, extra)
    gpu = b200.GpuOnlyInference.from_gguf(p, max_seq_len=32)
    assert gpu.load_stats["tensors_skipped"] == 1 and gpu.load_stats["tensors_loaded"] == len(tensors)
    gpu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n_devices", [1, 2])
def test_group_from_gguf(b200, oracle, tmp_path, n_devices):
    """Single-process group (tensor parallel over n devices): every rank stages its own shard out of the one mapping."""
    if b200.device_count() < n_devices:
        pytest.skip(f"needs {n_devices} GPUs")
    from llama_gguf_b200 import gguf_io

    arch, desc, tensors = synth.synth_model("llama-stream-tiny", "Q4_K_M", 64, vocab=1024)   # (vocab / world_size: a multiple of 16)
    path = os.path.join(tmp_path, "llama-stream-tiny-v1024.gguf")
    gguf_io.write_gguf(path, arch, desc, tensors)
    toks = synth.prompt_tokens(5, desc["vocab"])
    grp = b200.GroupInference.from_gguf(path, n_devices=n_devices, max_seq_len=64)
    if n_devices > 1:
        assert grp.load_stats["device_bytes"] < grp.load_stats["tensor_bytes"]      # rank 0 copied its shard only
    for t in toks[:-1]:
        grp.prefill_token(t)
    got = grp.forward(toks[-1])
    assert rel_err(got, oracle.OracleModel(desc, tensors).forward(toks)) < TOL
    grp.close()
