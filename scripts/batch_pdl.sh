#!/bin/bash
# batched-decode lines of bench.py without (0) and with (1) programmatic dependent launches inside the step's CUDA graph
for g in 0 1; do
  B200_BATCH_PDL=$g python bench.py --depth 0 --prefill-len 0 --no-cpu-baseline --no-speculation --steps 16 --warmup 4 > gpurun_out/bp_$g.json 2> gpurun_out/bp_$g.err
  python -c "
import json; j=json.load(open('gpurun_out/bp_$g.json')); print('batch pdl $g', {k:(round(v.get('value',0)),v.get('ms_per_step')) for k,v in j['extras'].items() if isinstance(v,dict)}, j['extras'].get('batch_error'))"
done
