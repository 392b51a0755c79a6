#!/bin/bash
# 2-GPU run: NCCL tests (native halo + torch P2P), weak-scaling S3 bench on 2 GPUs, strong-scaling S5 on 1 and 2 GPUs
OUT=gpurun_out/r02e; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu.txt
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_kernels_gpu.py -m gpu -q -s > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_s3_2gpu.json 2> $OUT/bench_s3_2gpu.err; echo "bench2 exit $?"
timeout 900 python bench.py --workload s5 --bodies 128 --steps 10 --warmup 3 > $OUT/bench_s5_128_1gpu.json 2> $OUT/bench_s5_1gpu.err; echo "s5 1gpu exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload s5 --bodies 128 --gpus 2 --steps 10 --warmup 3 > $OUT/bench_s5_128_2gpu.json 2> $OUT/bench_s5_2gpu.err; echo "s5 2gpu exit $?"
tail -15 $OUT/pytest_multi.log; for f in bench_s3_2gpu bench_s5_128_1gpu bench_s5_128_2gpu; do echo "== $f"; cut -c1-900 $OUT/$f.json; tail -3 $OUT/${f%.json}.err 2>/dev/null; done; tail -5 $OUT/bench_s3_2gpu.err $OUT/bench_s5_1gpu.err $OUT/bench_s5_2gpu.err
