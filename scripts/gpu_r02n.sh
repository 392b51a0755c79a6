#!/bin/bash
# final validation of the round: smoke, full GPU suite, the driver's bench line (both arms), the other workloads
OUT=gpurun_out/r02n; mkdir -p $OUT
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 240 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference exit $?"
timeout 90 python bench.py --workload s2 --steps 12 --warmup 3 > $OUT/bench_s2.json 2> $OUT/bench_s2.err; echo "s2 exit $?"
timeout 200 python bench.py --workload s4 --steps 10 --warmup 3 --preroll 40 > $OUT/bench_s4.json 2> $OUT/bench_s4.err; echo "s4 exit $?"
TICKS=125 timeout 120 python scripts/dev_s3.py > $OUT/timeline.log 2>&1
SKIP=60 TICKS=1 timeout 180 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/prof_ticks.py > $OUT/launches.log 2>&1
python scripts/launch_summary.py $OUT/launches.csv > $OUT/launches.summary.txt 2>&1
tail -2 $OUT/smoke.log; grep -v "^$" $OUT/pytest.log | tail -8; cut -c1-300 $OUT/bench.json; cut -c1-900 $OUT/bench_reference.json; tail -2 $OUT/bench_reference.err; cut -c1-700 $OUT/bench_s2.json; tail -3 $OUT/bench_s2.err; cut -c1-500 $OUT/bench_s4.json; tail -2 $OUT/bench_s4.err; head -12 $OUT/launches.summary.txt
