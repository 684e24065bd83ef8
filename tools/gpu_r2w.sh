#!/bin/bash
# round 2, GPU call W2: persistent depthwise-conv kernel (taps once per CTA, next tile requested under the stores): parity + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "mobileclip or fastvit or c2" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2w_pytest.log
grep -h "passed\|failed\|Error\|pytest exit" gpurun_out/r2w_pytest.log | cut -c1-220 | head
for w in mobileclip2_vision mobileclip2_s3_vision mobileclip2_s4_vision; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2w_bench_$w.json 2> gpurun_out/r2w_bench_$w.err
done
python - <<'PY'
import json
for w in ["mobileclip2_vision","mobileclip2_s3_vision","mobileclip2_s4_vision"]:
    try:
        d=json.loads(open(f"gpurun_out/r2w_bench_{w}.json").read().strip().splitlines()[0]); print(w, round(d['value']), {a:round(b,2) for a,b in d['roofline']['kernel_ms_per_step'].items() if b})
    except Exception as e: print(w, "failed", e)
PY
