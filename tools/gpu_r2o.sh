#!/bin/bash
# round 2, GPU call O: where the fused ConvMlp kernel spends its time (clock64 probes; GELU-free timing variant)
mkdir -p gpurun_out
timeout 300 tests/native/gemm_test_fmlp_timing.bin 8 > gpurun_out/r2o_fmlp_timing.log 2>&1; echo "exit $?" >> gpurun_out/r2o_fmlp_timing.log
timeout 300 tests/native/gemm_test_fmlp_nogelu.bin 8 > gpurun_out/r2o_fmlp_nogelu.log 2>&1; echo "exit $?" >> gpurun_out/r2o_fmlp_nogelu.log
grep -A2 "M=1048576\|M=262144\|M=65536" gpurun_out/r2o_fmlp_timing.log
grep -A2 "M=1048576\|M=262144\|M=65536" gpurun_out/r2o_fmlp_nogelu.log
