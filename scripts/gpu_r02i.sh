#!/bin/bash
OUT=gpurun_out/r02i; mkdir -p $OUT
timeout 400 python -m pytest tests/test_pbd_gpu.py tests/test_solver_gpu.py -m gpu -q -k "pbd or device_vertex or svd or island or clear" > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 150 python scripts/diag_pbd_scale.py > $OUT/pbd_scale.log 2>&1; echo "pbd scale exit $?"
timeout 240 python bench.py --workload s2 --steps 10 --warmup 3 --preroll 10 > $OUT/bench_s2.json 2> $OUT/bench_s2.err; echo "s2 exit $?"
timeout 300 python bench.py --workload s5 --steps 10 --warmup 3 > $OUT/bench_s5_512_1gpu.json 2> $OUT/bench_s5.err; echo "s5 exit $?"
timeout 240 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
grep -v "^$" $OUT/pytest.log | tail -12; cat $OUT/pbd_scale.log; cut -c1-1300 $OUT/bench_s2.json; tail -2 $OUT/bench_s2.err; cut -c1-500 $OUT/bench_s5_512_1gpu.json; tail -2 $OUT/bench_s5.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02i/bench.json') if l.startswith('{')][0])
print({k:d[k] for k in ("value","ms_per_step","phase_ms_per_step")}, d["e2e"]["ms_per_step"], d["free_fall"])
PY
