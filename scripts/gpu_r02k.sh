#!/bin/bash
OUT=gpurun_out/r02k; mkdir -p $OUT
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_pbd_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest.log
timeout 200 python bench.py --workload s2 --steps 10 --warmup 3 > $OUT/bench_s2.json 2> $OUT/bench_s2.err; echo "s2 exit $?"
timeout 200 python bench.py --workload s5 --bodies 128 --steps 10 --warmup 3 > $OUT/bench_s5_128.json 2> $OUT/bench_s5.err; echo "s5 exit $?"
TICKS=125 timeout 120 python scripts/dev_s3.py > $OUT/timeline.log 2>&1
grep -v "^$" $OUT/pytest.log | tail -6; cut -c1-1500 $OUT/bench_s2.json; tail -2 $OUT/bench_s2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02k/bench_s5_128.json') if l.startswith('{')][0])
print("s5 128:", {k:d[k] for k in ("value","ms_per_step","phase_ms_per_step","pcg_iterations_per_step")})
PY
tail -3 $OUT/bench_s5.err; tail -8 $OUT/timeline.log | cut -c1-300
