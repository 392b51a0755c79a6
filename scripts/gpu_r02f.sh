#!/bin/bash
OUT=gpurun_out/r02f; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_solver_gpu.py tests/test_configs_gpu.py tests/test_abi.py -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
TICKS=125 timeout 600 python scripts/dev_s3.py > $OUT/timeline.log 2>&1
timeout 600 python bench.py --workload s2 --steps 20 --warmup 5 > $OUT/bench_s2.json 2> $OUT/bench_s2.err; echo "s2 exit $?"
timeout 900 python bench.py --workload s4 --steps 10 --warmup 3 > $OUT/bench_s4.json 2> $OUT/bench_s4.err; echo "s4 exit $?"
SKIP=60 TICKS=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"k_island_pcg|k_tet_elems|k_gather_rhs|k_sort_scatter|k_sort_hist|k_scan_tile|k_scan_add|k_pair_filter|k_ccd|k_gs_mid|k_gs_cluster|k_block_factor|k_tri_ranges|k_emit_pairs|k_isl_" \
  -c 150 -f -o $OUT/prof_s3_tick61 python scripts/prof_ticks.py > $OUT/prof_s3.log 2>&1
WORKLOAD=s4 SIZE=1000 SKIP=30 TICKS=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_shape|k_goal|k_pt_project|k_gather" -c 12 -f -o $OUT/prof_s4 python scripts/prof_other.py > $OUT/prof_s4.log 2>&1
WORKLOAD=s2 SIZE=100000 SKIP=20 TICKS=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_pbd|k_sweep|k_node" -c 24 -f -o $OUT/prof_s2 python scripts/prof_other.py > $OUT/prof_s2.log 2>&1
WORKLOAD=s2 SIZE=100000 SKIP=20 TICKS=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_s2.csv python scripts/prof_other.py > $OUT/launches_s2.log 2>&1
python scripts/launch_summary.py $OUT/launches_s2.csv > $OUT/launches_s2.summary.txt 2>&1
WORKLOAD=s4 SIZE=1000 SKIP=30 TICKS=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_s4.csv python scripts/prof_other.py > $OUT/launches_s4.log 2>&1
python scripts/launch_summary.py $OUT/launches_s4.csv > $OUT/launches_s4.summary.txt 2>&1
grep -v "^$" $OUT/pytest.log | tail -12; tail -16 $OUT/timeline.log; cut -c1-1200 $OUT/bench_s2.json; tail -2 $OUT/bench_s2.err; cut -c1-1200 $OUT/bench_s4.json; tail -2 $OUT/bench_s4.err; head -14 $OUT/launches_s2.summary.txt; head -14 $OUT/launches_s4.summary.txt; tail -2 $OUT/prof_s3.log $OUT/prof_s4.log $OUT/prof_s2.log
