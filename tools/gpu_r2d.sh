#!/bin/bash
# round 2, GPU call D: cross-item QK look-ahead in the attention kernel (correctness on every case, A/B, in-step), ncu
set -x
mkdir -p gpurun_out
timeout 300 tests/native/attn_test.bin > gpurun_out/r2d_attn_vt.log 2>&1; echo "exit $?" >> gpurun_out/r2d_attn_vt.log
tail -19 gpurun_out/r2d_attn_vt.log | cut -c1-60,95-
ATTN_NO_VT=1 timeout 300 tests/native/attn_test.bin > gpurun_out/r2d_attn_natural.log 2>&1; echo "exit $?" >> gpurun_out/r2d_attn_natural.log
tail -19 gpurun_out/r2d_attn_natural.log | cut -c1-60,95-
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_baseline_configs_gpu.py tests/test_real_export_gpu.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
CLIPB200_ATTN_VT=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-text > gpurun_out/r2d_bench_novt.json 2> gpurun_out/r2d_bench_novt.err
timeout 300 python bench.py --workload gopt_vision --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2d_bench_gopt.json 2> gpurun_out/r2d_bench_gopt.err
python - <<'PY'
import json
for f in ["r2d_bench","r2d_bench_novt","r2d_bench_gopt"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[0])
        r=d["roofline"]; print(f, round(d["value"],1), d["clocks"]["sm_mhz"], {k:round(v,1) for k,v in r["kernel_ms_per_step"].items()}, "text", d.get("text") and d["text"].get("value"))
    except Exception as e: print(f, "ERR", e)
PY
# ncu: the attention kernel of the SO400M micro-batch shape (case 11: one correctness launch + 3 warm-up + 10 timed)
timeout 200 tests/native/attn_test.bin 11 > gpurun_out/r2d_ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcgen05 -s 2 -c 2 -o gpurun_out/r02d_attn tests/native/attn_test.bin 11 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
