#!/bin/bash
OUT=gpurun_out/r02m; mkdir -p $OUT
PIES_BENCH_VERBOSE=40 timeout 70 python bench.py --workload s2 --steps 6 --warmup 3 > $OUT/bench_s2.json 2> $OUT/bench_s2.err; echo "s2 exit $?"
cut -c1-600 $OUT/bench_s2.json; tail -40 $OUT/bench_s2.err
