#!/bin/bash
TAG=${1:-r02p}
OUT=gpurun_out; mkdir -p $OUT
( python -m pytest tests/test_grid_sampler_gpu.py tests/test_dcn_gpu.py -m gpu -q -s 2>&1 | grep -E "^\[dcn|^\.*\[dcn|passed|failed|FAILED|Error" | tail -30 ) > $OUT/${TAG}_pytest_gs_dcn.log 2>&1
python scripts/bench_ops.py > $OUT/${TAG}_ops.json 2> $OUT/${TAG}_ops.err
for v in i8w28 i8w32; do
  AB_CAPS=128 B200_BEV_OPS_LIB=$PWD/bevformer_tensorrt_b200/lib/libb200_bev_ops_${v}.so python scripts/ab_msda.py > $OUT/${TAG}_ab_msda_${v}.json 2>> $OUT/${TAG}_ab.err
done
AB_CAPS=128 python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2>> $OUT/${TAG}_ab.err
( timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_grid_sampler_gpu.py -m gpu -q -x -k "tile_path" 2>&1 | tail -8 ) > $OUT/${TAG}_sanitizer_gs.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"grid_sample_2d_tile" -c 4 \
    -o $OUT/${TAG}_prof_gs_tile -f python scripts/bench_ops.py > $OUT/${TAG}_ncu_gs.log 2>&1
( time python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > $OUT/${TAG}_pytest.log 2>&1
ls -la $OUT | tail -8
