#!/bin/bash
# round 2, GPU call M: fused ConvMlp with fine-grained W rings: native check, parity, A/B, ncu of the stage-2 / stage-3 shapes
set -x
mkdir -p gpurun_out
timeout 300 tests/native/gemm_test.bin 8 > gpurun_out/r2m_fused_mlp.log 2>&1; rc=$?; echo "exit $rc" >> gpurun_out/r2m_fused_mlp.log
cut -c1-40,95- gpurun_out/r2m_fused_mlp.log
if grep -q "FUSED MLP TEST PASSED" gpurun_out/r2m_fused_mlp.log; then
  timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_baseline_configs_gpu.py -m gpu -x -q -k "mobileclip or fastvit or c2 or tiny_mobileclip" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
  tail -3 gpurun_out/r2m_pytest.log
  for fm in 1 0; do
    CLIPB200_FUSED_MLP=$fm timeout 300 python bench.py --workload mobileclip2_vision --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_mc_fused${fm}.json 2>> gpurun_out/r2m_bench.err
  done
  python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2m_mc_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[0]); r=d["roofline"]
    print(f, round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k:round(v,2) for k,v in r["kernel_ms_per_step"].items() if v>0}, "gemm TF", round(r["achieved"]))
PY
fi
