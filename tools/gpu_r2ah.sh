#!/bin/bash
# round 2, GPU call AH: small-feature-map depthwise variants (8x8 / 4x4 images: 2 / 4 images per CTA iteration): parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mobileclip or fastvit or c2" > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ah_pytest.log
tail -3 gpurun_out/r2ah_pytest.log
for w in mobileclip2_vision mobileclip2_s3_vision mobileclip2_s4_vision; do
  for v in 1 0; do
    CLIPB200_DWCONV_SMALL=$v timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2ah_${w}_small$v.json 2> gpurun_out/r2ah_${w}_small$v.err
  done
done
python - <<'PY'
import json
for w in ["mobileclip2_vision","mobileclip2_s3_vision","mobileclip2_s4_vision"]:
    for v in (1,0):
        try:
            d=json.loads(open(f"gpurun_out/r2ah_{w}_small{v}.json").read().strip().splitlines()[0]); print(w, "small" if v else "16x16", round(d['value']), {a:round(b,2) for a,b in d['roofline']['kernel_ms_per_step'].items() if b})
        except Exception as e: print(w, v, "failed", e)
PY
