#!/bin/bash
# round 2, GPU call E (multi-GPU box): in-process pool parity + one-process scaling, torchrun extras check
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_pool_gpu.py -m gpu -x -q -s > gpurun_out/r2e_pytest_pool.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest_pool.log
tail -5 gpurun_out/r2e_pytest_pool.log
for n in 1 2 4 8; do
  if [ $n -le $NG ]; then
    timeout 600 python bench.py --pool --gpus $n --steps 4 --warmup 2 > gpurun_out/r02_pool_${n}gpu.json 2> gpurun_out/r02_pool_${n}gpu.err
    tail -c 700 gpurun_out/r02_pool_${n}gpu.json; echo
  fi
done
if [ $NG -ge 2 ]; then
  timeout 600 python bench.py --pool --gpus $NG --workload dfn5b_text --steps 4 --warmup 2 > gpurun_out/r02_pool_text_${NG}gpu.json 2> gpurun_out/r02_pool_text_${NG}gpu.err
  tail -c 500 gpurun_out/r02_pool_text_${NG}gpu.json; echo
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2e_torchrun2.json 2> gpurun_out/r2e_torchrun2.err
  python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2e_torchrun2.json").read().strip().splitlines()[-1])
    print("torchrun N=2:", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "text", d["text"], "strong", d["strong"])
except Exception as e:
    print("torchrun ERR", e); print(open("gpurun_out/r2e_torchrun2.err").read()[-1500:])
PY
fi
