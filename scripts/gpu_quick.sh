#!/bin/bash
# One short gpurun call: memcheck of the smoke scene, GPU parity tests, bench, S3 phase timeline, ncu launch list of one contact tick.
# usage: scripts/gpu_quick.sh <tag>     (outputs under gpurun_out/<tag>/)
TAG=${1:-run}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $OUT/memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $OUT/memcheck.log
timeout 900 python -m pytest tests -m gpu -x -q $PYTEST_ARGS > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
TICKS=${TIMELINE_TICKS:-125} timeout 600 python scripts/dev_s3.py > $OUT/timeline.log 2>&1
SKIP=${PROF_SKIP:-60} TICKS=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $OUT/launches.csv python scripts/prof_ticks.py > $OUT/launches.log 2>&1
tail -3 $OUT/memcheck.log; tail -5 $OUT/pytest.log; cat $OUT/bench.json; tail -16 $OUT/timeline.log
python scripts/launch_summary.py $OUT/launches.csv | head -16
if [ -n "$PROF_KERNELS" ]; then
SKIP=${PROF_SKIP:-60} TICKS=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"$PROF_KERNELS" -c ${PROF_COUNT:-4} -f -o $OUT/prof python scripts/prof_ticks.py > $OUT/prof.log 2>&1
fi
