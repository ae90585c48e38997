#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
n=${1:-4}
B200_BENCH_HANG_DUMP=90 timeout 120 $TR --nproc-per-node $n --master-port 29911 bench.py --gpus $n --steps 30 --warmup 5 > $OUT/r02x_bench_n${n}_default.json 2> $OUT/r02x_bench_n${n}_default.err
tail -n 3 $OUT/r02x_bench_n${n}_default.err
