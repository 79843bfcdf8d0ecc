#!/bin/bash
# batched-decode lines of bench.py with eager launches (0) and the per-row-count CUDA graph (1)
for g in 0 1; do
  B200_BATCH_GRAPH=$g python bench.py --depth 0 --prefill-len 0 --no-cpu-baseline --no-speculation --steps 16 --warmup 4 > gpurun_out/bg_$g.json 2> gpurun_out/bg_$g.err
  python -c "
import json; j=json.load(open('gpurun_out/bg_$g.json')); print('batch graph $g', {k:(round(v.get('value',0)),v.get('ms_per_step')) for k,v in j['extras'].items() if isinstance(v,dict)}, j['extras'].get('batch_error'))"
done
