#!/bin/bash
# A/B kernel timing of alternative builds: scripts/ab_bench.sh tagA tagB ...  ("default" = the in-tree library)
OUT=gpurun_out; mkdir -p $OUT
for tag in "$@"; do
  if [ "$tag" = "default" ]; then unset B200_BEV_OPS_LIB; else export B200_BEV_OPS_LIB=$PWD/bevformer_tensorrt_b200/lib/libb200_bev_ops_${tag}.so; fi
  for dist in U G; do
    python bench.py --steps 50 --warmup 10 --dist $dist --no-cpu-baseline --no-secondary --e2e-steps 1 2>>$OUT/ab.err | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', '$dist', 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], d['clocks'])"
  done
done
