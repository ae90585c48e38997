#!/bin/bash
# round-2 visit 7: resident-tail FP16 kernel — parity, sanitizer, A/B, ncu
TAG=${1:-r02g}
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "resident" 2>&1 | tail -30 ) > $OUT/${TAG}_pytest_res.log 2>&1
( timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "resident and (many_pairs or three_levels or one_pixel)" 2>&1 | tail -25 ) > $OUT/${TAG}_sanitizer_res.log 2>&1
python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2> $OUT/${TAG}_ab_msda.err
( time python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $OUT/${TAG}_pytest.log 2>&1
for cfg in "f16 U" "f16 G"; do
  set -- $cfg
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"msda_res" -s 2 -c 1 \
    -o $OUT/${TAG}_prof_res_$1_$2 -f python scripts/prof_msda.py $1 $2 3 1 > $OUT/${TAG}_ncu_$1_$2.log 2>&1
done
ls -la $OUT | tail -12
