#!/bin/bash
# round 2, GPU call AA: final build: full GPU suite, smoke, default bench; then MobileCLIP2-S2 launch list and ncu --set full of
# the persistent depthwise 7x7 kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2aa_pytest.log
tail -3 gpurun_out/r2aa_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2aa_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2aa_smoke.log
tail -2 gpurun_out/r2aa_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2aa_bench.json").read().strip().splitlines()[0])
r=d["roofline"]; print("bench", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["clocks"]["sm_mhz"], {k:round(v,1) for k,v in r["kernel_ms_per_step"].items() if v}, "frac", round(r["frac"],3))
print("text", d["text"]["value"]); m=d["mobileclip2"]; print("mobileclip2", m["vision"]["value"], m["text"]["value"])
PY
CMD="python bench.py --workload mobileclip2_vision --steps 1 --warmup 1 --no-extras --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/r2aa_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 500 --csv --log-file gpurun_out/r02aa_mobileclip2_launches.csv $CMD > gpurun_out/r2aa_ncu1.log 2>&1
timeout 200 $CMD > gpurun_out/r2aa_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_tma_kernel -s 60 -c 1 -o gpurun_out/r02aa_dwconv7 $CMD > gpurun_out/r2aa_ncu2.log 2>&1
tail -2 gpurun_out/r2aa_ncu2.log
