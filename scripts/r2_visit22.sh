#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
n=${1:-2}
B200_BENCH_HANG_DUMP=150 timeout 200 $TR --nproc-per-node $n --master-port 29801 bench.py --gpus $n --steps 40 --warmup 5 --dist G > $OUT/r02v_bench_n${n}_G.json 2> $OUT/r02v_bench_n${n}_G.err
B200_BENCH_HANG_DUMP=150 timeout 200 $TR --nproc-per-node $n --master-port 29802 bench.py --gpus $n --steps 40 --warmup 5 --dist U > $OUT/r02v_bench_n${n}_U.json 2> $OUT/r02v_bench_n${n}_U.err
tail -n 5 $OUT/r02v_bench_n${n}_G.err
