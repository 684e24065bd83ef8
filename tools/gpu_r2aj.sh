#!/bin/bash
# round 2, GPU call AJ: ncu --set full of the persistent 7x7 depthwise kernel (stage-2 launch of MobileCLIP2-S2) and of the
# small-feature-map variant (stage 4)
mkdir -p gpurun_out
CMD="python bench.py --workload mobileclip2_vision --steps 1 --warmup 1 --no-extras --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/r2aj_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dwconv_tma_kernel<\(int\)7, __nv_bfloat16' -s 8 -c 1 -o gpurun_out/r02aj_dwconv7 $CMD > gpurun_out/r2aj_ncu1.log 2>&1
tail -1 gpurun_out/r2aj_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dwconv_small_kernel<\(int\)7, __nv_bfloat16' -s 1 -c 1 -o gpurun_out/r02aj_dwconv7_small $CMD > gpurun_out/r2aj_ncu2.log 2>&1
tail -1 gpurun_out/r2aj_ncu2.log
