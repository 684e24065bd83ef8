#!/bin/bash
# round 2, GPU call A: full -m gpu suite, default bench line, micro-batch A/B (L2-slab question), photo workload
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
for mb in 32 64 128; do
  timeout 300 python bench.py --steps 5 --warmup 3 --micro-batch $mb --no-text --no-extras --no-cpu-baseline > gpurun_out/r2a_bench_mb$mb.json 2> gpurun_out/r2a_bench_mb$mb.err
done
timeout 600 python bench.py --workload so400m_photos --steps 3 --warmup 1 > gpurun_out/r2a_photos.json 2> gpurun_out/r2a_photos.err
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_bench.json | head -c 3000
cat gpurun_out/r2a_photos.json | head -c 2000
