#!/bin/bash
OUT=gpurun_out/r02z; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
SKIP=60 TICKS=1 timeout 180 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/prof_ticks.py > $OUT/launches.log 2>&1
python scripts/launch_summary.py $OUT/launches.csv > $OUT/launches.summary.txt 2>&1
SKIP=80 TICKS=1 timeout 180 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_t81.csv python scripts/prof_ticks.py > $OUT/launches_t81.log 2>&1
python scripts/launch_summary.py $OUT/launches_t81.csv > $OUT/launches_t81.summary.txt 2>&1
SKIP=80 TICKS=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_island_direct|k_island_dense|k_island_invert|k_gather_rhs_contacts|k_tet_elems|k_cell_rank|k_pair_filter|k_ccd" -c 10 -f -o $OUT/prof_t81 python scripts/prof_ticks.py > $OUT/prof.log 2>&1
ls -la $OUT | head -20
grep -v "^$" $OUT/pytest.log | tail -6; python - <<'PY'
import json
for f in ("bench",):
    d=json.loads([l for l in open('gpurun_out/r02z/%s.json'%f) if l.startswith('{')][0])
    print(f, {k:d[k] for k in ("value","ms_per_step","phase_ms_per_step","pcg_iterations_per_step")}, "e2e", d["e2e"]["ms_per_step"], "ff", d["free_fall"]["ms_per_step"], d["free_fall"]["phase_ms_per_step"], "isl", d["roofline"]["avg_launch_ms"])
    print(d["roofline_other_kernels"]["gather"]["avg_launch_ms"], d["roofline_local_step_plus_rhs"])
PY
head -16 $OUT/launches.summary.txt; head -16 $OUT/launches_t81.summary.txt
