#!/bin/sh
# compute-sanitizer over the kernels added late in round 2: the STATS epilogue of conv_tc_kernel (32/64-column chunks, 8 epilogue
# warps, CTA pair), partial_reduce_kernel, pack_gather1_kernel (run under gpurun; logs -> gpurun_out/, summaries -> profiles/).
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='test_conv_stats_epilogue_and_partials_finalize or test_relu_maxpool_final_pack_kernels'
mkdir -p gpurun_out
for TOOL in memcheck synccheck; do
  timeout 420 $SAN --tool $TOOL --log-file gpurun_out/sanitizer_r2b_$TOOL.log python -m pytest tests/test_conv_gpu.py tests/test_train_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_r2b_${TOOL}_pytest.log 2>&1
  echo "== $TOOL: pytest rc $? ; $(tail -1 gpurun_out/sanitizer_r2b_${TOOL}_pytest.log)"
  grep -E "ERROR SUMMARY|Error|hazard" gpurun_out/sanitizer_r2b_$TOOL.log | sort | uniq -c | head -8
done
