#!/bin/bash
# N=2: graph mode vs eager, G and U
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
n=2
B200_BENCH_HANG_DUMP=150 timeout 200 $TR --nproc-per-node $n --master-port 29701 bench.py --gpus $n --steps 40 --warmup 5 --dist G > $OUT/r02n_bench_n${n}_G_graph.json 2> $OUT/r02n_bench_n${n}_G_graph.err
B200_BENCH_HANG_DUMP=150 timeout 200 $TR --nproc-per-node $n --master-port 29702 bench.py --gpus $n --steps 40 --warmup 5 --dist G --no-graph > $OUT/r02n_bench_n${n}_G_eager.json 2> $OUT/r02n_bench_n${n}_G_eager.err
B200_BENCH_HANG_DUMP=150 timeout 200 $TR --nproc-per-node $n --master-port 29703 bench.py --gpus $n --steps 40 --warmup 5 --dist U > $OUT/r02n_bench_n${n}_U_graph.json 2> $OUT/r02n_bench_n${n}_U_graph.err
tail -n 5 $OUT/r02n_bench_n2_G_graph.err
