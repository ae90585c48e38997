#!/bin/bash
# round-2 visit 8: INT8 resident kernel with shared-memory records — parity, sanitizer, A/B, ncu
TAG=${1:-r02h}
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "int8 or v2 or config4" 2>&1 | tail -30 ) > $OUT/${TAG}_pytest_i8.log 2>&1
( timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "v2_int8" 2>&1 | tail -25 ) > $OUT/${TAG}_sanitizer_i8.log 2>&1
AB_CAPS=128 python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2> $OUT/${TAG}_ab_msda.err
for cfg in "i8 U" "i8 G"; do
  set -- $cfg
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"msda_i8p" -s 2 -c 1 \
    -o $OUT/${TAG}_prof_$1_$2 -f python scripts/prof_msda.py $1 $2 3 1 > $OUT/${TAG}_ncu_$1_$2.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_i8U.csv python scripts/prof_msda.py i8 U 4 1 > /dev/null 2>&1
( time python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $OUT/${TAG}_pytest.log 2>&1
ls -la $OUT | tail -12
