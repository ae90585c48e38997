#!/bin/bash
# round-2 visit 4: INT8 packed path + generic DCN kernel — parity, A/B, launch-bound variants of the round-1 kernel
TAG=${1:-r02d}
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $OUT/${TAG}_pytest.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_msda_gpu.py tests/test_dcn_gpu.py -m gpu -q -x -k "v2_int8 or envelope or fp32_matches_oracle or dcnv2p or packed_weight" 2>&1 | tail -25 ) > $OUT/${TAG}_sanitizer.log 2>&1
python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2> $OUT/${TAG}_ab_msda.err
for tag in mb4 mb2; do
  B200_BEV_OPS_LIB=$PWD/bevformer_tensorrt_b200/lib/libb200_bev_ops_${tag}.so python scripts/ab_msda.py > $OUT/${TAG}_ab_msda_${tag}.json 2>> $OUT/${TAG}_ab_msda.err
done
python scripts/bench_ops.py > $OUT/${TAG}_ops.json 2> $OUT/${TAG}_ops.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_i8U.csv python scripts/prof_msda.py i8 U 4 1 > /dev/null 2>&1
ls -la $OUT | tail -8
