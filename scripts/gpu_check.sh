#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, S3 phase timeline, ncu launch list + full capture of one contact tick.
# usage: scripts/gpu_check.sh <tag>     (outputs under gpurun_out/<tag>/)
TAG=${1:-run}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
TICKS=${TIMELINE_TICKS:-150} timeout 600 python scripts/dev_s3.py > $OUT/timeline.log 2>&1
SKIP=${PROF_SKIP:-60} TICKS=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $OUT/launches.csv python scripts/prof_ticks.py > $OUT/launches.log 2>&1
SKIP=${PROF_SKIP:-60} TICKS=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:"${PROF_KERNELS:-k_tet_elems|k_gather_rhs|k_pcg_spmv|k_pcg_update}" -c ${PROF_COUNT:-8} -f -o $OUT/prof python scripts/prof_ticks.py > $OUT/prof.log 2>&1
tail -3 $OUT/pytest.log; tail -2 $OUT/smoke.log; cat $OUT/bench.json; tail -5 $OUT/timeline.log
