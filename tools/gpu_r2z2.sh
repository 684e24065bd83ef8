#!/bin/bash
# round 2, GPU call Z6: photo path: process-wide staging-copy workers vs threads spawned per staging group
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resize_gpu.py tests/test_pool_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
for w in so400m_photos mobileclip2_photos; do
  for pool in 1 0; do
    CLIPB200_COPY_POOL=$pool timeout 400 python bench.py --workload $w --photos 112 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2z6_${w}_pool${pool}_$rep.json 2> gpurun_out/r2z6_${w}_pool${pool}_$rep.err
  done
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z6_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[0]); print(f.split("/")[-1], round(d["value"]), "img/s", round(d["h2d_gb_per_s"],1), "GB/s", round(d["ms_per_step"],1))
    except Exception as e: print(f, "failed", e)
PY
