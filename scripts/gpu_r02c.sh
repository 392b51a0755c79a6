#!/bin/bash
OUT=gpurun_out/r02c; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
TICKS=125 timeout 600 python scripts/dev_s3.py > $OUT/timeline.log 2>&1
SKIP=60 TICKS=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/prof_ticks.py > $OUT/launches.log 2>&1
python scripts/launch_summary.py $OUT/launches.csv > $OUT/launches.summary.txt 2>&1
SKIP=100 TICKS=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_island_pcg" -c 3 -f -o $OUT/prof_t101 python scripts/prof_ticks.py > $OUT/prof.log 2>&1
grep -v "^$" $OUT/pytest.log | tail -25; cat $OUT/bench.json; tail -3 $OUT/bench.err; tail -28 $OUT/timeline.log; head -16 $OUT/launches.summary.txt
