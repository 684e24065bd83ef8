#!/bin/bash
# round 2, GPU call B: transposed-V attention (stage 1): native checks, engine parity, A/B bench
set -x
mkdir -p gpurun_out
timeout 300 tests/native/gemm_test.bin 7 > gpurun_out/r2b_gemm_qkvt.log 2>&1; echo "exit $?" >> gpurun_out/r2b_gemm_qkvt.log
timeout 300 tests/native/attn_test.bin > gpurun_out/r2b_attn_vt.log 2>&1; echo "exit $?" >> gpurun_out/r2b_attn_vt.log
for c in 11 12 13 14; do ATTN_NO_VT=1 timeout 120 tests/native/attn_test.bin $c; done > gpurun_out/r2b_attn_novt.log 2>&1
tail -25 gpurun_out/r2b_gemm_qkvt.log; tail -25 gpurun_out/r2b_attn_vt.log; cat gpurun_out/r2b_attn_novt.log
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_baseline_configs_gpu.py tests/test_real_export_gpu.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_vt.json 2> gpurun_out/r2b_bench_vt.err
CLIPB200_ATTN_NO_VT=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2b_bench_novt.json 2> gpurun_out/r2b_bench_novt.err
timeout 400 python bench.py --workload so400m_photos --steps 3 --warmup 1 --photos 112 --no-cpu-baseline > gpurun_out/r2b_photos.json 2> gpurun_out/r2b_photos.err
python - <<'PY'
import json
for f in ["r2b_bench_vt","r2b_bench_novt"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        r=d["roofline"]; print(f, round(d["value"],1), d["clocks"]["sm_mhz"], {k:round(v,1) for k,v in r["kernel_ms_per_step"].items()}, "text", d["text"] and round(d["text"]["value"]))
    except Exception as e: print(f, "ERR", e)
print(open("gpurun_out/r2b_photos.json").read()[:1500])
PY
