#!/bin/bash
# round 2, GPU call F (8-GPU box): one host process feeding 8 GPUs through the pool
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1
lscpu | head -20 > gpurun_out/r2f_lscpu.txt 2>&1
timeout 600 python bench.py --pool --gpus 8 --steps 4 --warmup 2 > gpurun_out/r02_pool_8gpu.json 2> gpurun_out/r02_pool_8gpu.err
tail -c 800 gpurun_out/r02_pool_8gpu.json; echo
timeout 600 python bench.py --pool --gpus 8 --workload dfn5b_text --steps 4 --warmup 2 > gpurun_out/r02_pool_text_8gpu.json 2> gpurun_out/r02_pool_text_8gpu.err
tail -c 500 gpurun_out/r02_pool_text_8gpu.json; echo
timeout 300 python -m pytest tests/test_pool_gpu.py -m gpu -x -q > gpurun_out/r2f_pytest_pool.log 2>&1; tail -3 gpurun_out/r2f_pytest_pool.log
