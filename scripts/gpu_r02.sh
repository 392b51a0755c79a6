#!/bin/bash
# One gpurun call of round 2: GPU parity tests, bench (driver's flags), S3 phase timeline, ncu launch list of one contact tick,
# optional ncu --set full capture.   usage: scripts/gpu_r02.sh <tag>     (outputs under gpurun_out/<tag>/)
TAG=${1:-run}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
timeout ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -x -q $PYTEST_ARGS > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
fi
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup ${BENCH_WARMUP:-5} $BENCH_ARGS > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
fi
if [ -z "$SKIP_TIMELINE" ]; then
TICKS=${TIMELINE_TICKS:-125} timeout 600 python scripts/dev_s3.py > $OUT/timeline.log 2>&1
fi
if [ -z "$SKIP_LAUNCHES" ]; then
SKIP=${PROF_SKIP:-60} TICKS=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $OUT/launches.csv python scripts/prof_ticks.py > $OUT/launches.log 2>&1
python scripts/launch_summary.py $OUT/launches.csv > $OUT/launches.summary.txt 2>&1
fi
if [ -n "$PROF_KERNELS" ]; then
SKIP=${PROF_SKIP:-60} TICKS=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"$PROF_KERNELS" -c ${PROF_COUNT:-4} -f -o $OUT/prof python scripts/prof_ticks.py > $OUT/prof.log 2>&1
fi
tail -8 $OUT/pytest.log 2>/dev/null; cat $OUT/bench.json 2>/dev/null; tail -3 $OUT/bench.err 2>/dev/null; tail -30 $OUT/timeline.log 2>/dev/null
head -24 $OUT/launches.summary.txt 2>/dev/null
