#!/bin/bash
# round 2, GPU call J: same-box A/B ladder of the attention item-boundary changes + packed-exponential variants
set -x
mkdir -p gpurun_out
export ATTN_NO_VT=1
for v in 111_pexp1 111_pexp2; do
  timeout 300 tests/native/attn_ab_$v.bin > gpurun_out/r2j_full_$v.log 2>&1; echo "exit $?" >> gpurun_out/r2j_full_$v.log
  grep -c "^ok" gpurun_out/r2j_full_$v.log; grep "FAIL\|exit\|PASSED\|FAILED" gpurun_out/r2j_full_$v.log | head -8
done
: > gpurun_out/r2j_ladder.log
for rep in 1 2 3; do
  for v in 000 100 110 111 111_pexp1 111_pexp2; do
    for c in 11 15 16; do
      echo -n "rep$rep $v case$c " >> gpurun_out/r2j_ladder.log
      timeout 120 tests/native/attn_ab_$v.bin $c | head -1 | sed 's/.*tcgen05//' >> gpurun_out/r2j_ladder.log
    done
  done
done
cat gpurun_out/r2j_ladder.log
