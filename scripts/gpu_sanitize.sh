#!/bin/sh
# compute-sanitizer over one small convolution per kernel variant (fast, strict 3-MMA pair, strict N-concatenated, K-chunked) and
# one weight-gradient case (run under gpurun). The logs go to gpurun_out/sanitizer_*.log; summaries are committed under profiles/.
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='test_conv_case_matches_cpu_reference[3] or test_conv_cta_pair_matches_single_cta[1] or test_split_conv_case_matches_fp64_reference[0] or test_split_conv_case_matches_fp64_reference[5] or test_split_conv_case_matches_fp64_reference[15] or test_wgrad_matches_cpu_emulation[2]'
mkdir -p gpurun_out
for TOOL in memcheck synccheck racecheck; do
  timeout 420 $SAN --tool $TOOL --log-file gpurun_out/sanitizer_$TOOL.log python -m pytest tests/test_conv_gpu.py tests/test_train_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_${TOOL}_pytest.log 2>&1
  echo "== $TOOL: pytest rc $? ; $(tail -1 gpurun_out/sanitizer_${TOOL}_pytest.log)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_$TOOL.log | sort | uniq -c | head -8
done
