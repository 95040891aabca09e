#!/bin/bash
# Round-2 evidence run (under gpurun, one GPU): GPU tests, the plain bench, then -- only after the plain command exited 0 --
# the ncu launch list and ONE --set full capture of the headline kernel at the benchmarked chain count.
set -u
OUT=gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then python -m pytest tests -m gpu -q > $OUT/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r2_pytest.log; fi
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary"
$CMD > $OUT/r2_plain.json 2> $OUT/r2_plain.err; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_gibbs_launches.csv $CMD > $OUT/r2_ncu_l.log 2>&1; echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:sweep_v3 -c 1 -o $OUT/prof_r2_v3_gibbs -f $CMD > $OUT/r2_ncu_f.log 2>&1; echo "ncu full rc=$?"
fi
