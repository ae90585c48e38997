#!/bin/bash
TAG=${1:-r02c}
OUT=gpurun_out; mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_f16U.csv python scripts/prof_msda.py f16 U 4 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_i8U.csv python scripts/prof_msda.py i8 U 4 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_v2 -s 2 -c 1 -o $OUT/${TAG}_prof_v2_f16_U -f python scripts/prof_msda.py f16 U 3 1 > $OUT/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_v2 -s 2 -c 1 -o $OUT/${TAG}_prof_v2_i8_U -f python scripts/prof_msda.py i8 U 3 1 > $OUT/${TAG}_ncu2.log 2>&1
ls -la $OUT | tail -6
