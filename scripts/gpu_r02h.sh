#!/bin/bash
# multi-GPU scaling check with the driver's launch line: N = $NGPU ranks
N=${NGPU:-4}; OUT=gpurun_out/r02h_$N; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_s3.json 2> $OUT/bench_s3.err; echo "s3 exit $?"
if [ -n "$WITH_S5" ]; then
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --workload s5 --gpus $N --steps 10 --warmup 3 > $OUT/bench_s5.json 2> $OUT/bench_s5.err; echo "s5 exit $?"
fi
grep "^{" $OUT/bench_s3.json | cut -c1-700; tail -3 $OUT/bench_s3.err; grep "^{" $OUT/bench_s5.json 2>/dev/null | cut -c1-700; tail -3 $OUT/bench_s5.err 2>/dev/null
