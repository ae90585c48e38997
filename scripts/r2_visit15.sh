#!/bin/bash
# single GPU: grid sampler timing after the slice change, INT8 warp-count A/B, DCN FP16 error printout, reference-kernel
# golden vectors, full parity suite
TAG=${1:-r02o}
OUT=gpurun_out; mkdir -p $OUT
python scripts/bench_ops.py > $OUT/${TAG}_ops.json 2> $OUT/${TAG}_ops.err
for v in i8w28 i8w32; do
  AB_CAPS=128 B200_BEV_OPS_LIB=$PWD/bevformer_tensorrt_b200/lib/libb200_bev_ops_${v}.so python scripts/ab_msda.py > $OUT/${TAG}_ab_msda_${v}.json 2>> $OUT/${TAG}_ab.err
done
AB_CAPS=128 python scripts/ab_msda.py > $OUT/${TAG}_ab_msda.json 2>> $OUT/${TAG}_ab.err
python -m pytest tests/test_dcn_gpu.py -m gpu -q -s -k "fused_tcgen05" 2>&1 | grep -E "^\[dcn|passed|failed" > $OUT/${TAG}_dcn_err.log
python tests/golden/make_golden_ref_gpu.py $OUT/golden_ref_gpu > $OUT/${TAG}_golden_ref_gpu.log 2>&1
( time python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $OUT/${TAG}_pytest.log 2>&1
ls -la $OUT | tail -8
