#!/bin/bash
# round 2, GPU call C: deferred attention epilogue A/B, transposed-V fix (T % 32), double-S + transposed V
set -x
mkdir -p gpurun_out
timeout 300 tests/native/gemm_test.bin 7 > gpurun_out/r2c_gemm_qkvt.log 2>&1; echo "exit $?" >> gpurun_out/r2c_gemm_qkvt.log
tail -12 gpurun_out/r2c_gemm_qkvt.log
ATTN_NO_VT=1 timeout 300 tests/native/attn_test.bin > gpurun_out/r2c_attn_default.log 2>&1; echo "exit $?" >> gpurun_out/r2c_attn_default.log
tail -20 gpurun_out/r2c_attn_default.log
for c in 11 12 13 14 15 16; do ATTN_NO_VT=1 timeout 120 tests/native/attn_test_eager_epi.bin $c | head -1; done > gpurun_out/r2c_attn_eager.log 2>&1
cat gpurun_out/r2c_attn_eager.log
timeout 300 tests/native/attn_test.bin > gpurun_out/r2c_attn_vt.log 2>&1; echo "exit $?" >> gpurun_out/r2c_attn_vt.log
tail -8 gpurun_out/r2c_attn_vt.log
for c in 5 10 11 12 13 14; do CLIPB200_ATTN_DOUBLE_S=1 timeout 120 tests/native/attn_test.bin $c | head -1; done > gpurun_out/r2c_attn_vt_doubles.log 2>&1
cat gpurun_out/r2c_attn_vt_doubles.log
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_baseline_configs_gpu.py tests/test_real_export_gpu.py tests/test_resize_gpu.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
CLIPB200_ATTN_VT=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-text > gpurun_out/r2c_bench_vt.json 2> gpurun_out/r2c_bench_vt.err
python - <<'PY'
import json
for f in ["r2c_bench","r2c_bench_vt"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[0])
        r=d["roofline"]; print(f, round(d["value"],1), d["clocks"]["sm_mhz"], {k:round(v,1) for k,v in r["kernel_ms_per_step"].items()}, "text", d["text"] and d["text"].get("value"))
    except Exception as e: print(f, "ERR", e)
PY
