#!/bin/bash
# visit 26 (last seconds of the round's GPU budget): smoke, INT8 / FP16 DCN device time after the epilogue change, DCN parity file
TAG=${1:-r02ab}
OUT=gpurun_out; mkdir -p $OUT
timeout 25 python scripts/dcn_time.py > $OUT/${TAG}_dcn_time.json 2> $OUT/${TAG}_dcn_time.err
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
( timeout 40 python -m pytest tests/test_dcn_gpu.py -m gpu -q 2>&1 | tail -5 ) > $OUT/${TAG}_pytest_dcn.log 2>&1
