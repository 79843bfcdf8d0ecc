#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench line, then (only if the plain run exited 0) the ncu launch
# list and one full capture of the per-token megakernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
BARGS="--prompt-len 8 --steps 4 --warmup 3 --ctx 2048 --no-cpu-baseline"
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${1:-}" = "ref" ]; then
  timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "bench ref exit $?"; cat gpurun_out/bench_ref.json
fi
if [ "${1:-}" = "ncu" ] || [ "${2:-}" = "ncu" ]; then
  timeout 600 python bench.py $BARGS > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
      --log-file gpurun_out/launches.csv python bench.py $BARGS > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 600 python bench.py $BARGS > gpurun_out/bench_small2.json 2> gpurun_out/bench_small2.err &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"stream_decode|mega_decode" -s 9 -c 1 \
      -o gpurun_out/mega_full -f python bench.py $BARGS > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
  ls -la gpurun_out/ | tail -8
fi
if [ "${1:-}" = "tiny" ] || [ "${2:-}" = "tiny" ] || [ "${3:-}" = "tiny" ]; then   # BASELINE configs[1]: TinyLlama-1.1B Q8_0 / Q6_K, 2K context
  for mix in Q8_0 Q6_K; do
    timeout 600 python bench.py --model tinyllama-1.1b --mix $mix --ctx 2048 --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_tinyllama_$mix.json 2> gpurun_out/bench_tinyllama_$mix.err
    echo "tinyllama $mix exit $?"; python -c "import json,sys; d=json.loads(open('gpurun_out/bench_tinyllama_$mix.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])"
  done
fi
