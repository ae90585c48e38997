#!/bin/bash
TAG=${1:-r02e}
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $OUT/${TAG}_pytest.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_i8p -s 2 -c 1 -o $OUT/${TAG}_prof_i8p_U -f python scripts/prof_msda.py i8 U 3 1 > $OUT/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_i8p -s 2 -c 1 -o $OUT/${TAG}_prof_i8p_G -f python scripts/prof_msda.py i8 G 3 1 > $OUT/${TAG}_ncu2.log 2>&1
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
ls -la $OUT | tail -6
