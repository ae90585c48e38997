#!/bin/bash
# N=4: peer exchange (default) U and G; then the NCCL exchange at N=2 with a hang dump
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 4 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/r02l_bench_n4_U.json 2> $OUT/r02l_bench_n4_U.err
timeout 240 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 --dist G > $OUT/r02l_bench_n4_G.json 2> $OUT/r02l_bench_n4_G.err
B200_BENCH_HANG_DUMP=100 NCCL_DEBUG=WARN timeout 150 $TR --nproc-per-node 2 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl > $OUT/r02l_bench_n2_U_nccl.json 2> $OUT/r02l_bench_n2_U_nccl.err
tail -n 3 $OUT/r02l_bench_n4_U.err
