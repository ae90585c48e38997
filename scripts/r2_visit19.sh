#!/bin/bash
TAG=${1:-r02s}
OUT=gpurun_out; mkdir -p $OUT
python scripts/exp/dbg_gs_tile.py > $OUT/${TAG}_dbg_gs.log 2>&1
if grep -q "f32 rc 0" $OUT/${TAG}_dbg_gs.log; then
  ( python -m pytest tests/test_grid_sampler_gpu.py -m gpu -q 2>&1 | tail -8 ) > $OUT/${TAG}_pytest_gs.log 2>&1
  python scripts/bench_ops.py > $OUT/${TAG}_ops.json 2> $OUT/${TAG}_ops.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"grid_sample_2d_tile" -c 4 \
      -o $OUT/${TAG}_prof_gs_tile -f python scripts/bench_ops.py > $OUT/${TAG}_ncu_gs.log 2>&1
else
  python -c "
import bevformer_tensorrt_b200 as bt
bt._lib.load().b200_grid_sample_set_tile_path(0)
" ; B200_GS_TILE=0 python scripts/bench_ops.py > $OUT/${TAG}_ops_generic.json 2> $OUT/${TAG}_ops.err
fi
AB_CAPS=128 python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2>> $OUT/${TAG}_ab.err
( time python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > $OUT/${TAG}_pytest.log 2>&1
ls -la $OUT | tail -6
