#!/bin/bash
# round-2 visit 1: gather microbenchmarks (access shapes for the MSDA rework), the new parity tests, INT8 ncu capture
TAG=${1:-r02a}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_smi.csv 2>&1
./scripts/micro/gather_bw2 > $OUT/${TAG}_micro_gather_bw2.txt 2>&1
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > $OUT/${TAG}_pytest.log 2>&1
( python -m pytest tests/test_msda_gpu.py tests/test_grid_sampler_gpu.py -m gpu -q -s -k "config4 or base_shape_fp16_config4 or base_shape_int8_config4" 2>&1 | grep -E "^\[|passed|failed" ) > $OUT/${TAG}_config4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_gather -s 3 -c 1 -o $OUT/${TAG}_prof_i8_U -f \
    python bench.py --steps 3 --warmup 3 --dtype i8 --no-secondary --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_ncu_i8_U.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_gather -s 3 -c 1 -o $OUT/${TAG}_prof_i8_G -f \
    python bench.py --steps 3 --warmup 3 --dtype i8 --dist G --no-secondary --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_ncu_i8_G.log 2>&1
ls -la $OUT | tail -20
