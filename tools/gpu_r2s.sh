#!/bin/bash
# round 2, GPU call S: short-sequence attention kernel (T <= 80, hd 64): native check + A/B against the long-sequence kernel
mkdir -p gpurun_out
export ATTN_NO_VT=1
for i in 6 8 17 18; do
  timeout 60 tests/native/attn_test.bin $i 2>&1 | grep -v "^$" | head -8
  CLIPB200_ATTN_SHORT=0 timeout 60 tests/native/attn_test.bin $i 2>&1 | grep "^ok\|^FAIL" | sed 's/^/   long kernel: /'
done > gpurun_out/r2s_attn_short.log 2>&1
cat gpurun_out/r2s_attn_short.log | cut -c1-260
