#!/bin/bash
# visit 25: A/B of the batched MSDA launch shapes (visibility scan), their bit-identity tests, INT8 DCN epilogue
TAG=${1:-r02aa}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.csv 2>&1
timeout 120 python scripts/ab_batch.py $OUT/${TAG}_ab_batch.json > $OUT/${TAG}_ab.log 2>&1
( time timeout 150 python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "batched" 2>&1 | tail -15 ) > $OUT/${TAG}_pytest_batched.log 2>&1
( timeout 60 python -m pytest tests/test_dcn_gpu.py -m gpu -q -k "int8" 2>&1 | tail -8 ) > $OUT/${TAG}_pytest_dcn_i8.log 2>&1
timeout 90 python scripts/bench_ops.py > $OUT/${TAG}_ops.json 2> $OUT/${TAG}_ops.err
ls -la $OUT | tail -8
