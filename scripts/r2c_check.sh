#!/bin/bash
# quick visit: streamed-kernel parity tests, 8B bench line, timeline, 70B N=1 line
set -u
mkdir -p gpurun_out
export B200_LOG=1
timeout 600 python -m pytest tests/test_gpu_stream.py tests/test_gpu_configs.py tests/test_gpu_model.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --depth ${DEPTH:-0} --batch 0 --prefill-len 0 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
echo "bench exit $?"; grep -a "b200\]" gpurun_out/bench_q.err | grep -v "phase" | head -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print("value %.1f tok/s  %.3f ms  e2e %.1f  frac %.3f  tokens %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("greedy_tokens_head")), d.get("extras"))
PY
if [ "${TL:-1}" = "1" ]; then timeout 300 python scripts/s2_timeline.py llama-3-8b Q4_K_M 128 8192 2>/dev/null > gpurun_out/timeline.txt; head -8 gpurun_out/timeline.txt; fi
if [ "${BIG:-1}" = "1" ]; then
timeout 900 python bench.py --model llama-3-70b --steps 32 --warmup 4 --depth 0 --batch 0 --prefill-len 0 --no-cpu-baseline > gpurun_out/bench_llama-3-70b.json 2> gpurun_out/bench_llama-3-70b.err
echo "70b exit $?"; grep -a "b200\]" gpurun_out/bench_llama-3-70b.err | grep -v "phase" | head -5; cut -c1-200 gpurun_out/bench_llama-3-70b.json
fi
