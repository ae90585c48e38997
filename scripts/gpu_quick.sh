#!/bin/bash
# quick GPU visit: all gpu tests + kernel timings (no ncu)
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
( python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $OUT/${TAG}_pytest.log 2>&1
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
