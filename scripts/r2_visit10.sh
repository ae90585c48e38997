#!/bin/bash
TAG=${1:-r02j}
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $OUT/${TAG}_pytest.log 2>&1
AB_CAPS=128 python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2> $OUT/${TAG}_ab_msda.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"grid_sample_2d" -c 3 \
    -o $OUT/${TAG}_prof_gs -f python scripts/bench_ops.py > $OUT/${TAG}_ncu_gs.log 2>&1
ls -la $OUT | tail -6
