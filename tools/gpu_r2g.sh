#!/bin/bash
# round 2, GPU call G: full GPU suite on the final kernels, default bench line, launch list + GEMM traffic under ncu
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2g_smoke.log
tail -5 gpurun_out/r2g_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[0])
r=d["roofline"]; print("bench", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"], {k:round(v,1) for k,v in r["kernel_ms_per_step"].items()}, "frac", round(r["frac"],3))
print("text", d["text"]["value"]); m=d["mobileclip2"]; print("mobileclip2", m["vision"]["value"], m["vision"]["roofline"]["kernel_ms_per_step"], m["text"]["value"])
PY
timeout 300 python bench.py --workload mobileclip2_s3_vision --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_s3.json 2> gpurun_out/r2g_bench_s3.err
timeout 300 python bench.py --workload mobileclip2_s4_vision --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_s4.json 2> gpurun_out/r2g_bench_s4.err
# launch list of one micro-batch of the bench (cold-cache, serialised: shares, not absolutes), same recipe as round 1
CMD="python bench.py --workload so400m_vision --batch 256 --steps 1 --warmup 1 --no-text --no-extras --no-cpu-baseline"
KREG='regex:gemm_bf16|flash_attention|attn_fwd|layernorm|preprocess|l2_normalize|map_pool|write_cls|affine_rows'
timeout 300 $CMD > gpurun_out/r2g_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 200 -c 200 --csv --log-file gpurun_out/r02g_launches.csv $CMD > gpurun_out/r2g_ncu1.log 2>&1
tail -2 gpurun_out/r2g_ncu1.log
