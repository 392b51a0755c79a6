#!/bin/bash
# 2 GPUs: per-iteration halo exchanges skipped when no island mixes owned and ghost rows — NCCL tests + S3 weak scaling
OUT=gpurun_out/r03i; mkdir -p $OUT
timeout 500 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_s3_2gpu.json 2> $OUT/bench_s3_2gpu.err; echo "bench2 exit $?"
grep -v "^$" $OUT/pytest_multi.log | tail -5; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r03i/bench_s3_2gpu.json') if l.startswith('{')][0])
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["phase_ms_per_step"], d.get("halo"))
except Exception as e:
    print("failed", e)
PY
tail -n 3 $OUT/bench_s3_2gpu.err | cut -c1-300
