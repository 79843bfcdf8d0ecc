#!/bin/bash
# batched-decode lines of bench.py with 2 and 3 dequant groups in the persistent dequant-GEMM (B200_GEMM2_GROUPS)
for g in 2 3; do
  B200_GEMM2_GROUPS=$g python bench.py --depth 0 --prefill-len 0 --no-cpu-baseline --no-speculation --steps 16 --warmup 4 > gpurun_out/bb_$g.json 2> gpurun_out/bb_$g.err
  python -c "
import json; j=json.load(open('gpurun_out/bb_$g.json')); print('groups $g', {k:(round(v.get('value',0)),v.get('ms_per_step')) for k,v in j['extras'].items() if isinstance(v,dict)})"
done
