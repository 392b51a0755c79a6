#!/bin/bash
# bounded: every step has its own timeout; ncu captures are a handful of kernels (reports stay small)
OUT=gpurun_out/r02g; mkdir -p $OUT
timeout 420 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 240 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
TICKS=125 timeout 120 python scripts/dev_s3.py > $OUT/timeline.log 2>&1
SKIP=60 TICKS=1 timeout 180 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/prof_ticks.py > $OUT/launches.log 2>&1
python scripts/launch_summary.py $OUT/launches.csv > $OUT/launches.summary.txt 2>&1
timeout 300 python bench.py --workload s4 --steps 5 --warmup 3 --preroll 10 > $OUT/bench_s4.json 2> $OUT/bench_s4.err; echo "s4 exit $?"
timeout 120 python bench.py --workload s2 --bodies 10000 --steps 5 --warmup 3 --preroll 5 > $OUT/bench_s2_10k.json 2> $OUT/bench_s2.err; echo "s2 exit $?"
SKIP=60 TICKS=1 timeout 240 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_island_pcg|k_tet_elems|k_gather_rhs" -c 6 -f -o $OUT/prof_tick61 python scripts/prof_ticks.py > $OUT/prof.log 2>&1
grep -v "^$" $OUT/pytest.log | tail -8; cut -c1-2500 $OUT/bench.json; tail -2 $OUT/bench.err; tail -27 $OUT/timeline.log; head -14 $OUT/launches.summary.txt; cut -c1-1300 $OUT/bench_s4.json; tail -2 $OUT/bench_s4.err; cut -c1-1000 $OUT/bench_s2_10k.json; tail -2 $OUT/bench_s2.err; du -sh $OUT
