#!/bin/bash
# the driver's launch line on 4 B200s, final library: S3 weak scaling
OUT=gpurun_out/r03j; mkdir -p $OUT
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_s3_4gpu.json 2> $OUT/bench_s3_4gpu.err; echo "bench4 exit $?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r03j/bench_s3_4gpu.json') if l.startswith('{')][0])
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["phase_ms_per_step"], d.get("halo"))
except Exception as e:
    print("failed", e)
PY
tail -n 3 $OUT/bench_s3_4gpu.err
