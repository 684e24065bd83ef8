#!/bin/bash
# round 2, GPU call W: pixel-lane 7x7 depthwise kernel (uniform-register taps): FastViT parity + same-box A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "mobileclip or fastvit or c2" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2w_pytest.log
grep -h "^\[C2\|real graph\|passed\|failed\|Error\|pytest exit\|\[mobileclip\|cos" gpurun_out/r2w_pytest.log | cut -c1-220 | head -20
for v in 1 0; do
  CLIPB200_DWCONV_PX=$v timeout 300 python bench.py --workload mobileclip2_vision --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2w_bench_s2_px$v.json 2> gpurun_out/r2w_bench_s2_px$v.err
done
python - <<'PY'
import json
for v in (1,0):
    try:
        d=json.loads(open(f"gpurun_out/r2w_bench_s2_px{v}.json").read().strip().splitlines()[0]); print('px',v, round(d['value']), {a:round(b,2) for a,b in d['roofline']['kernel_ms_per_step'].items() if b}, d['gpu_launches'])
    except Exception as e: print(v, "failed", e)
PY
