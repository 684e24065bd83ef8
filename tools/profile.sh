#!/bin/bash
# ncu evidence for one round (run under gpurun, ONE GPU).  Usage: tools/profile.sh <tag>
# 1) plain run (must exit 0)  2) per-launch device times of one timed step  3) --set full of the 4 GEMMs of one
# transformer layer, one attention and one LayerNorm launch (SO400M, one micro-batch of 256 images = the bench's
# micro-batch)  4) MobileCLIP2-S2: launch list + --set full of the TMA depthwise conv and the stem conv.
# Reports land in gpurun_out/ and are summarised into profiles/ here.
set -e
TAG=${1:-r01}
CMD="python bench.py --workload so400m_vision --batch 256 --steps 1 --warmup 1 --no-text --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1
KREG='regex:gemm_bf16|flash_attention|attn_fwd|layernorm|preprocess|l2_normalize|map_pool|write_cls|affine_rows'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 200 -c 200 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 114 -c 4 \
    -f -o gpurun_out/${TAG}_gemm $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:flash_attention|attn_fwd' -s 28 -c 1 \
    -f -o gpurun_out/${TAG}_attn $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:layernorm -s 56 -c 1 \
    -f -o gpurun_out/${TAG}_ln $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
MC="python bench.py --workload mobileclip2_vision --batch 256 --steps 1 --warmup 1 --no-text --no-cpu-baseline"
$MC > gpurun_out/${TAG}_mc_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 204 -c 204 --csv \
    --log-file gpurun_out/${TAG}_mc_launches.csv $MC > gpurun_out/${TAG}_ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwconv_tma -s 100 -c 2 \
    -f -o gpurun_out/${TAG}_mc_dwconv $MC > gpurun_out/${TAG}_ncu6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_conv -s 1 -c 1 \
    -f -o gpurun_out/${TAG}_mc_stem $MC > gpurun_out/${TAG}_ncu7.log 2>&1
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-400
