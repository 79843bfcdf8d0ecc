"""GPU check of the second streamed megakernel (stream2.cuh) against the first streamed kernel / first megakernel and, for
small models, the CPU oracle.
usage: python scripts/s2_debug.py MIX [PROMPT_LEN CTX PRESET N_LAYERS VOCAB]      env MODES=s2,s1,mega  GREEDY=8"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import llama_gguf_b200 as B
import synth

a = sys.argv[1:]
mix = a[0] if len(a) > 0 else "Q4_K_M"
n_prompt = int(a[1]) if len(a) > 1 else 5
ctx = int(a[2]) if len(a) > 2 else 64
preset = a[3] if len(a) > 3 else "llama-stream-tiny"
n_layers = int(a[4]) if len(a) > 4 else None
vocab = int(a[5]) if len(a) > 5 else None
n_greedy = int(os.environ.get("GREEDY", "8"))
if preset.startswith("custom:"):   # custom:hidden,ffn,heads,kv,vocab,layers[,head_dim]
    v = [int(x) for x in preset.split(":")[1].split(",")]
    synth.TINY[preset] = dict(arch="llama", hidden=v[0], n_layers=v[5], n_heads=v[2], n_kv_heads=v[3], head_dim=v[6] if len(v) > 6 else 128,
                              ffn=v[1], vocab=v[4], norm_eps=1e-5, rope_base=5e5, rope_neox=0, bias=False, tied=False)
small = preset in synth.TINY and not preset.startswith("custom:")
if small:
    arch, desc, tensors = synth.synth_model(preset, mix, ctx, n_layers=n_layers, vocab=vocab)
else:
    from llama_gguf_b200 import randmodel
    p = dict(synth.get_preset(preset))
    if n_layers: p["n_layers"] = n_layers
    if vocab: p["vocab"] = vocab
    desc, tensors = randmodel.random_model(p, mix, ctx)
prompt = synth.prompt_tokens(n_prompt, desc["vocab"])
ENV = {"s2": dict(B200_STREAM2="1", B200_STREAM="1"), "s1": dict(B200_STREAM2="0", B200_STREAM="1"), "mega": dict(B200_STREAM2="0", B200_STREAM="0")}
outs = {}
for mode in os.environ.get("MODES", "s2,s1").split(","):
    os.environ.update(ENV[mode])
    gpu = B.GpuOnlyInference(desc, tensors)
    logits = B.GpuModelWrapper(gpu).forward(prompt, 0)
    tok = int(np.argmax(logits))
    dev, ms = gpu.decode_greedy(tok, n_greedy)
    bufs = []
    for which, n in ((0, desc["hidden"]), (1, desc["hidden"]), (3, desc["n_heads"] * desc["head_dim"]), (4, desc["ffn"])):
        o = (ctypes.c_float * n)()
        B.lib().b200_debug_read(gpu._h, which, o, n)
        bufs.append(np.array(o[:n]))
    outs[mode] = (np.array(logits), dev.tolist(), ms / n_greedy, gpu.path(), bufs)
    e = (ctypes.c_int * 8)()
    B.lib().b200_debug_err(gpu._h, e)
    if e[0]: print("mode", mode, "WATCHDOG", list(e))
    gpu.close()
tag = f"{preset} {mix} prompt {n_prompt}"
for m, o in outs.items():
    print(f"[{tag}] {m}: path {o[3]} {o[2]:.3f} ms/token, tokens {o[1]}, finite {bool(np.all(np.isfinite(o[0])))}")
modes = list(outs)
for m in modes[1:]:
    print(f"[{tag}] {modes[0]} vs {m}: logits rel err {synth.rel_err(outs[modes[0]][0], outs[m][0]):.3e}  tokens equal {outs[modes[0]][1] == outs[m][1]}"
          + "  bufs " + " ".join(f"{synth.rel_err(x, y):.1e}" for x, y in zip(outs[modes[0]][4], outs[m][4])))
if small and os.environ.get("ORACLE", "1") == "1":
    import oracle as O
    want = O.OracleModel(desc, tensors).forward(prompt)
    for m in modes:
        print(f"[{tag}] {m} vs oracle: rel err {synth.rel_err(outs[m][0], want):.3e}  argmax equal {int(np.argmax(outs[m][0])) == int(np.argmax(want))}")
