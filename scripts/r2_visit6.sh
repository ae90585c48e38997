#!/bin/bash
# round-2 visit 6 (re-entry): re-establish every measurement — microbenchmark, full parity suite, A/B, bench both arms,
# launch list, full ncu captures of the four headline MSDA configs.
TAG=${1:-r02f}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_smi.csv 2>&1
./scripts/micro/gather_bw2 > $OUT/${TAG}_micro_gather_bw2.txt 2>&1
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $OUT/${TAG}_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2> $OUT/${TAG}_ab_msda.err
python bench.py --steps 50 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
python scripts/bench_ops.py > $OUT/${TAG}_ops.json 2> $OUT/${TAG}_ops.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-ref-gpu --e2e-steps 1 > $OUT/${TAG}_ncu_launch.log 2>&1
for cfg in "f16 U" "f16 G" "i8 U" "i8 G"; do
  set -- $cfg
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"msda_gather|msda_i8p" -s 2 -c 1 \
    -o $OUT/${TAG}_prof_$1_$2 -f python scripts/prof_msda.py $1 $2 3 1 > $OUT/${TAG}_ncu_$1_$2.log 2>&1
done
ls -la $OUT | tail -30
