#!/bin/bash
# round 2, GPU call S: short-sequence attention kernel (T <= 80, hd 64): native check + A/B against the long-sequence kernel,
# text-tower parity tests, text benches with both kernels
mkdir -p gpurun_out
export ATTN_NO_VT=1
for i in 6 8 17 18; do
  timeout 60 tests/native/attn_test.bin $i 2>&1 | grep -v "^$" | head -8
  CLIPB200_ATTN_SHORT=0 timeout 60 tests/native/attn_test.bin $i 2>&1 | grep "^ok\|^FAIL" | sed 's/^/   long kernel: /'
done > gpurun_out/r2s_attn_short.log 2>&1
cat gpurun_out/r2s_attn_short.log | cut -c1-260
unset ATTN_NO_VT
timeout 900 python -m pytest tests -m gpu -x -q -k "text or clip or classify or c1 or c4 or parity" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2s_pytest.log; tail -3 gpurun_out/r2s_pytest.log
for w in dfn5b_text mobileclip2_text; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2s_bench_${w}_short.json 2> gpurun_out/r2s_bench_${w}_short.err
  CLIPB200_ATTN_SHORT=0 timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2s_bench_${w}_long.json 2> gpurun_out/r2s_bench_${w}_long.err
done
python - <<'PY'
import json
for w in ["dfn5b_text","mobileclip2_text"]:
    for k in ["short","long"]:
        try:
            d=json.loads(open(f"gpurun_out/r2s_bench_{w}_{k}.json").read().strip().splitlines()[0])
            print(w,k,round(d["value"]), {a:round(b,1) for a,b in d["roofline"]["kernel_ms_per_step"].items() if b}, d["clocks"]["sm_mhz"])
        except Exception as e: print(w,k,"failed",e)
PY
