#!/bin/bash
# round 2, GPU call Z: column-cropped photo staging: bit-exact resize tests, photo workloads with / without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resize_gpu.py tests/test_pool_gpu.py tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2z_pytest.log
tail -3 gpurun_out/r2z_pytest.log
for w in mobileclip2_photos so400m_photos; do
  for full in 0 1; do
    CLIPB200_STAGE_FULL_ROWS=$full timeout 400 python bench.py --workload $w --photos 112 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2z_${w}_full$full.json 2> gpurun_out/r2z_${w}_full$full.err
  done
done
python - <<'PY'
import json
for w in ["mobileclip2_photos","so400m_photos"]:
    for full in (0,1):
        try:
            d=json.loads(open(f"gpurun_out/r2z_{w}_full{full}.json").read().strip().splitlines()[0])
            print(w, "full_rows" if full else "cropped  ", round(d["value"]), "img/s", round(d["h2d_gb_per_s"],1), "GB/s of source", "pcie floor", round(d["pcie_floor"]["images_per_s"]))
        except Exception as e: print(w, full, "failed", e)
PY
