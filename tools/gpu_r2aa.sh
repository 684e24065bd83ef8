#!/bin/bash
# round 2, GPU calls AA / AE: final build: full GPU suite, smoke, default bench; then MobileCLIP2-S2 launch list and ncu --set full of
# the persistent depthwise 7x7 kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2an_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2an_pytest.log
tail -3 gpurun_out/r2an_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2an_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2an_smoke.log
tail -2 gpurun_out/r2an_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2an_bench.json 2> gpurun_out/r2an_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2an_bench.json").read().strip().splitlines()[0])
r=d["roofline"]; print("bench", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k:round(v,1) for k,v in r["kernel_ms_per_step"].items() if v}, "frac", round(r["frac"],3))
print("text", d["text"]["value"]); m=d["mobileclip2"]; print("mobileclip2", m["vision"]["value"], m["text"]["value"])
PY
timeout 120 tests/native/gemm_test.bin 8 2>&1 | tail -4 | cut -c1-200
