#!/bin/bash
# final validation of the round: smoke, full GPU suite, the driver's bench line (both arms), the other workloads,
# launch lists and the ncu captures VERDICT asked for
OUT=gpurun_out/r03c; mkdir -p $OUT
timeout 120 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference exit $?"
timeout 90 python bench.py --workload s2 --steps 12 --warmup 3 > $OUT/bench_s2.json 2> $OUT/bench_s2.err; echo "s2 exit $?"
timeout 200 python bench.py --workload s4 --steps 10 --warmup 3 --preroll 40 > $OUT/bench_s4.json 2> $OUT/bench_s4.err; echo "s4 exit $?"
timeout 300 python bench.py --workload s5 --bodies 128 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_s5_128_1gpu.json 2> $OUT/bench_s5.err; echo "s5 exit $?"
SKIP=60 TICKS=1 timeout 180 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/prof_ticks.py > $OUT/launches.log 2>&1
python scripts/launch_summary.py $OUT/launches.csv > $OUT/launches.summary.txt 2>&1
SKIP=60 TICKS=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_sort_scatter|k_block_factor|k_gs_mid|k_gs_cluster_friction|k_cell_rank|k_item_contacts" -c 8 -f -o $OUT/prof_misc python scripts/prof_ticks.py > $OUT/prof_misc.log 2>&1
WORKLOAD=s4 SIZE=1000 SKIP=30 TICKS=1 timeout 200 ncu --profile-from-start off --set full --clock-control none -k regex:"k_shape|k_goal" -c 3 -f -o $OUT/prof_shape python scripts/prof_other.py > $OUT/prof_shape.log 2>&1
WORKLOAD=s2 SIZE=100000 SKIP=2 TICKS=1 timeout 200 ncu --profile-from-start off --set full --clock-control none -k regex:"k_pbd" -c 8 -f -o $OUT/prof_pbd python scripts/prof_other.py > $OUT/prof_pbd.log 2>&1
ls -la $OUT | awk '{print $5, $9}'
tail -2 $OUT/smoke.log; grep -v "^$" $OUT/pytest.log | tail -5; cut -c1-300 $OUT/bench.json; cut -c1-600 $OUT/bench_reference.json; tail -2 $OUT/bench_reference.err; cut -c1-400 $OUT/bench_s2.json; tail -2 $OUT/bench_s2.err; cut -c1-400 $OUT/bench_s4.json; tail -2 $OUT/bench_s4.err; cut -c1-300 $OUT/bench_s5_128_1gpu.json; head -12 $OUT/launches.summary.txt
