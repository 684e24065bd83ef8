#!/bin/bash
# round 2, GPU call Z5: photo path: tapered micro-batch schedule (the last, exposed tower is a small one)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resize_gpu.py tests/test_pool_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for w in so400m_photos mobileclip2_photos; do
  for mb in 0 1; do
    CLIPB200_PHOTO_TAPER=$mb timeout 400 python bench.py --workload $w --photos 112 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2z5_${w}_taper$mb.json 2> gpurun_out/r2z5_${w}_taper$mb.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z5_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[0]); print(f.split("/")[-1], round(d["value"]), "img/s", round(d["h2d_gb_per_s"],1), "GB/s", {k:round(v,1) for k,v in d["kernel_ms_per_step"].items() if v}, round(d["ms_per_step"],1))
    except Exception as e: print(f, "failed", e)
PY
