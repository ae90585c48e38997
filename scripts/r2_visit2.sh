#!/bin/bash
# round-2 visit 2: v2 MSDA kernels — parity, sanitizer, A/B against the round-1 kernel
TAG=${1:-r02b}
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "v2 or emits or trace or enqueue" 2>&1 | tail -30 ) > $OUT/${TAG}_pytest_v2.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "v2_fp16_and_int8 or envelope" 2>&1 | tail -25 ) > $OUT/${TAG}_sanitizer_v2.log 2>&1
python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2> $OUT/${TAG}_ab_msda.err
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $OUT/${TAG}_pytest.log 2>&1
( python -m pytest tests/test_msda_gpu.py tests/test_grid_sampler_gpu.py -m gpu -q -s -k "config4" 2>&1 | grep -E "^\[|passed|failed" ) > $OUT/${TAG}_config4.log 2>&1
ls -la $OUT | tail -8
