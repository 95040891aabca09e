#!/bin/bash
# Round-2 evidence run (under gpurun, one GPU): GPU tests, the plain bench, then -- only after the plain command exited 0 --
# the ncu launch list and ONE --set full capture of the headline kernel at the benchmarked chain count.
set -u
OUT=gpurun_out
python -m pytest tests -m gpu -q > $OUT/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r2_pytest.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary"
$CMD > $OUT/r2_plain.json 2> $OUT/r2_plain.err; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_gibbs_launches.csv $CMD > $OUT/r2_ncu_l.log 2>&1; echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:sweep_v3 -c 1 -o $OUT/prof_r2_v3_gibbs -f $CMD > $OUT/r2_ncu_f.log 2>&1; echo "ncu full rc=$?"
fi
# compute-sanitizer availability probe (SURVEY 5 / VERDICT item 10)
( compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_csmc.py -q -x -k "tensor_core_kernel_shapes and 4-2-2-1" ) > $OUT/r2_sanitizer_probe.log 2>&1; echo "sanitizer rc=$?"; tail -3 $OUT/r2_sanitizer_probe.log
