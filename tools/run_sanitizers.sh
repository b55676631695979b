#!/bin/bash
# compute-sanitizer over the kernels that stage through shared memory / mbarriers / TMA (round 2 additions included)
O=gpurun_out
SEL='u16_frames_match and (30 or 100) and native or long_median and 256 or long_medmad and 257 or tensormap_staging and 64 and 3.0-3.0-5 or sorted_tensormap and 64 or two_ranks'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_stack.py tests/test_gpu_badpix.py tests/test_gpu_calibrate.py -m gpu -x -q \
     -k "(u16_frames and 30 and native and average-3.0) or (long_median and 256) or (long_medmad and 257) or (tensormap_staging and 64 and 5-mean) or (sorted_tensormap and 64) or (fused_calibrate and 64 and 2-u16_fits) or (badpix and dp) or (marked_pixels and 100 and f64) or (marked_pixels and 300 and u8 and kappa) or (lane_split and 512 and 5-mean)" \
     > $O/r02_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?" >> $O/r02_sanitizer_$tool.txt
  tail -4 $O/r02_sanitizer_$tool.txt
done
