#!/bin/bash
# N=8 / N=4 / N=2 scaling series with interleaved tiles and the unrolled peer kernel; both distributions
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  for d in G U; do
    B200_BENCH_HANG_DUMP=150 timeout 200 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 30 --warmup 5 --dist $d > $OUT/r02m_bench_n${n}_$d.json 2> $OUT/r02m_bench_n${n}_$d.err
  done
done
tail -n 2 $OUT/r02m_bench_n8_G.err
