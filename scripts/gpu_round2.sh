#!/bin/bash
# Lighter GPU visit for rounds where the MSDA / DCN kernels did not change: parity tests, smoke, bench lines, op timings,
# the ncu launch list of the bench command, and full captures of the new (small) kernels only.
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_smi.csv 2>&1
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $OUT/${TAG}_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
python bench.py --steps 100 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --steps 100 --warmup 10 --dist G --no-secondary --no-cpu-baseline > $OUT/${TAG}_bench_G.json 2>> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
python scripts/bench_ops.py > $OUT/${TAG}_ops.json 2>> $OUT/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 1 > $OUT/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rotate_|point_sampling" -s 3 -c 3 -o $OUT/${TAG}_prof_prologue -f \
    python scripts/prologue_prof.py 3 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | tail -20
