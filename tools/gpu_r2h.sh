#!/bin/bash
# round 2, GPU call H: two compute lanes (micro-batches alternate between two streams): parity + A/B
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_baseline_configs_gpu.py tests/test_real_export_gpu.py tests/test_pool_gpu.py tests/test_resize_gpu.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
for lanes in 2 1 2 1; do
  CLIPB200_LANES=$lanes timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-text > gpurun_out/r2h_bench_lanes${lanes}_$RANDOM.json 2>> gpurun_out/r2h_bench.err
done
timeout 300 python bench.py --workload dfn5b_text --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2h_text_lanes2.json 2>> gpurun_out/r2h_bench.err
CLIPB200_LANES=1 timeout 300 python bench.py --workload dfn5b_text --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2h_text_lanes1.json 2>> gpurun_out/r2h_bench.err
timeout 300 python bench.py --workload mobileclip2_vision --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_mc_lanes2.json 2>> gpurun_out/r2h_bench.err
CLIPB200_LANES=1 timeout 300 python bench.py --workload mobileclip2_vision --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_mc_lanes1.json 2>> gpurun_out/r2h_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[0])
        r=d["roofline"]; print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k:round(v,1) for k,v in r["kernel_ms_per_step"].items() if v>0})
    except Exception as e: print(f, "ERR", e)
PY
