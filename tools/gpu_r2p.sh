#!/bin/bash
# round 2, GPU call P: fused ConvMlp A/B on one box: chunk groups x polling MMA issue order
mkdir -p gpurun_out
for v in "" _g1p1 _g1p0 _fmlp_timing; do
  for rep in 1 2; do
    timeout 120 tests/native/gemm_test$v.bin 8 > gpurun_out/r2p_fmlp$v.$rep.log 2>&1; echo "exit $?" >> gpurun_out/r2p_fmlp$v.$rep.log
  done
  echo "== variant '$v'"; grep -h "two launches\|CTA 0\|MMA warp\|FAIL\|exit" gpurun_out/r2p_fmlp$v.2.log | cut -c1-420
done
