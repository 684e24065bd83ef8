#!/bin/bash
# round 2, GPU call I: idle-row warps skip the softmax arithmetic; refusal-policy test; bench
set -x
mkdir -p gpurun_out
ATTN_NO_VT=1 timeout 300 tests/native/attn_test.bin > gpurun_out/r2i_attn_natural.log 2>&1; echo "exit $?" >> gpurun_out/r2i_attn_natural.log
tail -19 gpurun_out/r2i_attn_natural.log | cut -c1-60,95-
timeout 300 tests/native/attn_test.bin > gpurun_out/r2i_attn_vt.log 2>&1; echo "exit $?" >> gpurun_out/r2i_attn_vt.log
tail -8 gpurun_out/r2i_attn_vt.log | cut -c1-60,95-
ATTN_NO_VT=1 timeout 300 tests/native/attn_test_eager_epi.bin > gpurun_out/r2i_attn_eager.log 2>&1; tail -3 gpurun_out/r2i_attn_eager.log | cut -c1-60,95-
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_baseline_configs_gpu.py tests/test_real_export_gpu.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -4 gpurun_out/r2i_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[0])
r=d["roofline"]; print("bench", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k:round(v,1) for k,v in r["kernel_ms_per_step"].items() if v>0}, "text", d["text"]["value"])
PY
