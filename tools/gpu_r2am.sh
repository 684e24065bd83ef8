#!/bin/bash
# round 2, GPU call AM: LayerNorm row order (descending: freshest rows first, first rows last) A/B on the SO400M and DFN5B steps
mkdir -p gpurun_out
for rep in 1 2; do
for v in 1 0; do
  CLIPB200_LN_DESCENDING=$v timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-text | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('so400m desc', '$v', round(d['value'],1), {k:round(x,1) for k,x in d['roofline']['kernel_ms_per_step'].items() if x}, d['clocks']['sm_mhz'])"
done
done
for v in 1 0; do
  CLIPB200_LN_DESCENDING=$v timeout 300 python bench.py --workload dfn5b_text --steps 4 --warmup 3 --no-cpu-baseline --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('dfn5b text desc', '$v', round(d['value'],1), {k:round(x,1) for k,x in d['roofline']['kernel_ms_per_step'].items() if x}, d['clocks']['sm_mhz'])"
done
