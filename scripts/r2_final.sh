#!/bin/bash
# round-2 closing visit: full parity suite, smoke, both bench arms, ops legs, launch list and ncu captures of the final kernels
TAG=${1:-r02z}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_smi.csv 2>&1
( time python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > $OUT/${TAG}_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
python scripts/bench_ops.py > $OUT/${TAG}_ops.json 2> $OUT/${TAG}_ops.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 60 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-ref-gpu --e2e-steps 1 > $OUT/${TAG}_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_i8G.csv python scripts/prof_msda.py i8 G 4 1 > /dev/null 2>&1
for cfg in "i8 U" "i8 G"; do
  set -- $cfg
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"msda_i8p" -s 2 -c 1 \
    -o $OUT/${TAG}_prof_$1_$2 -f python scripts/prof_msda.py $1 $2 3 1 > $OUT/${TAG}_ncu_$1_$2.log 2>&1
done
timeout 300 ncu --set full --clock-control none -k regex:"msda_pack" -s 2 -c 1 -o $OUT/${TAG}_prof_pack -f python scripts/prof_msda.py i8 U 3 1 > $OUT/${TAG}_ncu_pack.log 2>&1
ls -la $OUT | tail -12
