"""GPU check of the streamed megakernel (stream.cuh) against the first megakernel (B200_STREAM=0) and, for small
models, the CPU oracle.  usage: python scripts/stream_debug.py MIX [PROMPT_LEN CTX PRESET N_LAYERS VOCAB]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import llama_gguf_b200 as B
import synth

a = sys.argv[1:]
mix = a[0] if len(a) > 0 else "Q4_K_M"
n_prompt = int(a[1]) if len(a) > 1 else 5
ctx = int(a[2]) if len(a) > 2 else 64
preset = a[3] if len(a) > 3 else "stream-tiny"
n_layers = int(a[4]) if len(a) > 4 else None
vocab = int(a[5]) if len(a) > 5 else None
synth.TINY["stream-tiny"] = dict(arch="llama", hidden=2048, n_layers=2, n_heads=16, n_kv_heads=4, head_dim=128, ffn=4096,
                                 vocab=1000, norm_eps=1e-5, rope_base=5e5, rope_neox=0, bias=False, tied=False)
if preset.startswith("custom:"):   # custom:hidden,ffn,heads,kv,vocab,layers
    h, f, nh, nkv, v, nl = (int(x) for x in preset.split(":")[1].split(","))
    synth.TINY[preset] = dict(arch="llama", hidden=h, n_layers=nl, n_heads=nh, n_kv_heads=nkv, head_dim=128, ffn=f,
                              vocab=v, norm_eps=1e-5, rope_base=5e5, rope_neox=0, bias=False, tied=False)
small = preset == "stream-tiny"
if small:
    arch, desc, tensors = synth.synth_model(preset, mix, ctx)
else:
    from llama_gguf_b200 import randmodel
    p = dict(synth.get_preset(preset))
    if n_layers: p["n_layers"] = n_layers
    if vocab: p["vocab"] = vocab
    desc, tensors = randmodel.random_model(p, mix, ctx)
prompt = synth.prompt_tokens(n_prompt, desc["vocab"])
outs = {}
for mode in os.environ.get("MODES", "0,1").split(","):
    os.environ["B200_STREAM"] = mode
    gpu = B.GpuOnlyInference(desc, tensors)
    logits = B.GpuModelWrapper(gpu).forward(prompt, 0)
    tok = int(np.argmax(logits))
    dev, ms = gpu.decode_greedy(tok, 8)
    outs[mode] = (np.array(logits), dev.tolist(), ms / 8)
    import ctypes
    e = (ctypes.c_int * 8)()
    B.lib().b200_debug_err(gpu._h, e)
    if e[0]: print("mode", mode, "WATCHDOG", list(e))
    gpu.close()
if len(outs) < 2:
    print("ran modes", list(outs), "tokens", [o[1] for o in outs.values()]); sys.exit(0)
print("stream vs mega: rel err %.3e  tokens %s %s  ms/token %.3f vs %.3f" % (
    synth.rel_err(outs["1"][0], outs["0"][0]), outs["1"][1], outs["0"][1], outs["1"][2], outs["0"][2]))
if small:
    import oracle as O
    want = O.OracleModel(desc, tensors).forward(prompt)
    print("stream vs oracle: rel err %.3e" % synth.rel_err(outs["1"][0], want))
