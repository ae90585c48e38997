#!/bin/bash
TAG=${1:-r02i}
OUT=gpurun_out; mkdir -p $OUT
python scripts/exp/dbg_i8.py > $OUT/${TAG}_dbg_i8.log 2>&1
( python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "head_major" 2>&1 | tail -5 ) > $OUT/${TAG}_pytest_hm.log 2>&1
AB_CAPS=128 python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2> $OUT/${TAG}_ab_msda.err
AB_CAPS=128 B200_BEV_OPS_LIB=$PWD/bevformer_tensorrt_b200/lib/libb200_bev_ops_mb4.so python scripts/ab_msda.py > $OUT/${TAG}_ab_msda_mb4.json 2>> $OUT/${TAG}_ab_msda.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"msda_gather" -s 2 -c 1 \
    -o $OUT/${TAG}_prof_f16_G_hm -f python scripts/prof_msda.py f16 G 3 1 > $OUT/${TAG}_ncu_f16_G.log 2>&1
ls -la $OUT | tail -8
