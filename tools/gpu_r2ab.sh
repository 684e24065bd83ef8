#!/bin/bash
# round 2, GPU call AB: MobileCLIP2-S2 launch list (one 256-image step) and ncu --set full of the persistent depthwise 7x7 kernel
mkdir -p gpurun_out
CMD="python bench.py --workload mobileclip2_vision --steps 1 --warmup 1 --no-extras --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/r2ab_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 175 --csv --log-file gpurun_out/r02ab_mobileclip2_launches.csv $CMD > gpurun_out/r2ab_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dwconv_tma_kernel<7' -s 2 -c 1 -o gpurun_out/r02ab_dwconv7 $CMD > gpurun_out/r2ab_ncu2.log 2>&1
tail -2 gpurun_out/r2ab_ncu2.log; wc -l gpurun_out/r02ab_mobileclip2_launches.csv
