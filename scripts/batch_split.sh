#!/bin/bash
# batched-decode lines of bench.py with the round-1 split-K rule (0) and the persistent kernel's cost model (1)
for g in 0 1; do
  B200_GEMM2_SPLIT_PLAN=$g python bench.py --depth 0 --prefill-len 0 --no-cpu-baseline --no-speculation --steps 16 --warmup 4 > gpurun_out/bs_$g.json 2> gpurun_out/bs_$g.err
  python -c "
import json; j=json.load(open('gpurun_out/bs_$g.json')); print('split plan $g', {k:(round(v.get('value',0)),v.get('ms_per_step')) for k,v in j['extras'].items() if isinstance(v,dict)})"
done
