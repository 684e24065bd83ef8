#!/bin/bash
# round 2, GPU call R: FastViT real-graph binding (tiny + full-size S2 vs interpreter / OpenCV), fused ConvMlp after the GELU / ring changes
mkdir -p gpurun_out
timeout 120 tests/native/gemm_test.bin 8 > gpurun_out/r2r_fmlp.log 2>&1; echo "exit $?" >> gpurun_out/r2r_fmlp.log; tail -5 gpurun_out/r2r_fmlp.log | cut -c1-250
timeout 1200 python -m pytest tests/test_real_export_gpu.py tests/test_baseline_configs_gpu.py -m gpu -x -q -s -k "fastvit or c2 or mobileclip" > gpurun_out/r2r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2r_pytest.log
grep -h "^\[C2\|real graph\|passed\|failed\|Error\|pytest exit\|\[mobileclip" gpurun_out/r2r_pytest.log | cut -c1-300
timeout 300 python bench.py --workload mobileclip2_vision --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench_s2.json 2> gpurun_out/r2r_bench_s2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2r_bench_s2.json').read().strip().splitlines()[0]); print('S2', round(d['value']), d['roofline']['kernel_ms_per_step'])"
