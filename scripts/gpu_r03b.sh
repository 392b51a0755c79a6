#!/bin/bash
# 2-GPU run on the final library: NCCL tests (halo inside the library and through torch), S3 weak scaling on 2 GPUs,
# S5 (128 bodies) strong scaling 2 GPUs
OUT=gpurun_out/r03b; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_s3_2gpu.json 2> $OUT/bench_s3_2gpu.err; echo "bench2 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload s5 --bodies 128 --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_s5_128_2gpu.json 2> $OUT/bench_s5_2gpu.err; echo "s5 2gpu exit $?"
tail -6 $OUT/pytest_multi.log; python - <<'PY'
import json
for f in ("bench_s3_2gpu","bench_s5_128_2gpu"):
    try:
        d=json.loads([l for l in open('gpurun_out/r03b/%s.json'%f) if l.startswith('{')][0])
        print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["phase_ms_per_step"], d.get("halo"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 $OUT/bench_s3_2gpu.err $OUT/bench_s5_2gpu.err
