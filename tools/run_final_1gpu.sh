#!/bin/bash
# One 1-GPU session at the end of a round: full GPU test suite, the bench line, the reference arm, the launch list
# of a short bench run, and one ncu --set full launch of every kernel of the path (CSV only: the reports are large).
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02_gputest_tail.txt
python bench.py --steps 10 --warmup 3 > $O/r02_bench_1gpu.json 2>$O/bench1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2>$O/benchref.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'apgpu|stack_|calibrate|badpix|repair|flat_norm|stats_|imarith' --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-strong > $O/bench_under_ncu.log 2>&1
for what in stack frame; do
  ncu --set full --profile-from-start off --clock-control none -o /tmp/r02_$what python tools/prof_path_kernels.py $what > $O/prof_$what.log 2>&1
  ncu -i /tmp/r02_$what.ncu-rep --page raw --csv > $O/r02_${what}_raw.csv
done
cat $O/r02_gputest_tail.txt; head -c 600 $O/r02_bench_1gpu.json; echo; head -c 400 $O/r02_bench_reference_arm.json; echo; wc -l $O/r02_stack_raw.csv $O/r02_frame_raw.csv $O/r02_launches.csv
