#!/bin/bash
# round 2, GPU call AK: 16-channel blocks for widths like 80 (two tiles per iteration instead of a half-idle last block): parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mobileclip or fastvit or c2" > gpurun_out/r2ak_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ak_pytest.log
tail -3 gpurun_out/r2ak_pytest.log
for w in mobileclip2_vision; do
  for v in 1 0 1 0; do
    CLIPB200_DWCONV_HALF_BLOCKS=$v timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras | tee gpurun_out/r2ak_${w}_half$v.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('half', '$v', round(d['value']), round(d['roofline']['kernel_ms_per_step']['dwconv'],2))" 2> gpurun_out/r2ak_${w}_half$v.err
  done
done
python - <<'PY'
import json
for w in ["mobileclip2_vision"]:
    for v in (1,0):
        try:
            d=json.loads(open(f"gpurun_out/r2ak_{w}_half{v}.json").read().strip().splitlines()[0]); print(w, "5 x 16 ch" if v else "32+32+16", round(d['value']), {a:round(b,2) for a,b in d['roofline']['kernel_ms_per_step'].items() if b})
        except Exception as e: print(w, v, "failed", e)
PY
