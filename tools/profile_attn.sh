#!/bin/bash
# ncu evidence for the attention kernel of one round (run under gpurun, ONE GPU).  Usage: tools/profile_attn.sh <tag>
# 1) per-launch device times of one SO400M step (kernel shares)  2) --set full of one attention launch of that step.
set -e
TAG=${1:-r01g}
CMD="python bench.py --workload so400m_vision --batch 256 --steps 1 --warmup 1 --no-text --no-cpu-baseline"
KREG='regex:gemm_bf16|flash_attention|attn_fwd|layernorm|preprocess|l2_normalize|map_pool|write_cls|affine_rows'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 200 -c 200 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:flash_attention|attn_fwd' -s 28 -c 1 \
    -f -o gpurun_out/${TAG}_attn $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
