/* examples/minimal_host.c — the smallest C host of libb200_bev_ops (plain C99, no CUDA headers needed to compile).
 *
 *   gcc -std=c99 -I include examples/minimal_host.c -L bevformer_tensorrt_b200/lib -lb200_bev_ops \
 *       -Wl,-rpath,$PWD/bevformer_tensorrt_b200/lib -o minimal_host
 *
 * It sizes the tensors of one BEVFormer-tiny spatial-cross-attention MSDA call (BASELINE configs[1]: 6 cameras,
 * 2500 queries, 1 level 15x25, 8 heads x 32 channels, 8 points, 4 Z-anchors), asks the library what it would do with
 * them (format negotiation, argument validation) and — when device buffers are supplied by the embedding application —
 * would launch with b200_msda_f16. Without a GPU it stops after the host-side checks, which is what the CPU test runs.
 */
#include <stdio.h>
#include <string.h>

#include "b200_bev_ops.h"

int main(void) {
  const int batch = 6, heads = 8, channels = 32, levels = 1, queries = 2500, points = 8, groups = 4;
  const int spatial = 15 * 25;
  printf("library: %s\n", b200_bev_ops_version());

  /* the five inputs + one output as TensorRT would describe them to supportsFormatCombination */
  b200_tensor_desc io[6];
  memset(io, 0, sizeof(io));
  const int dims[6][4] = {{batch, spatial, heads, channels},      {levels, 2, 0, 0},
                          {batch, queries, 1, 2 * groups},        {batch, queries, heads, levels * points * 2},
                          {batch, queries, heads, levels * points}, {batch, queries, heads, channels}};
  const int nb[6] = {4, 2, 4, 4, 4, 4};
  for (int t = 0; t < 6; ++t) {
    io[t].dims.nbDims = nb[t];
    for (int d = 0; d < nb[t]; ++d) io[t].dims.d[d] = dims[t][d];
    io[t].type = t == 1 ? 3 /* kINT32 */ : 1 /* kHALF */;
    io[t].format = 0; /* kLINEAR */
    io[t].scale = 1.0f;
  }
  for (int pos = 0; pos < 6; ++pos)
    if (!b200_msda_supports_format(pos, io, 5, 1)) {
      printf("format combination rejected at position %d\n", pos);
      return 1;
    }
  printf("FP16 / kLINEAR accepted for all six tensors\n");

  /* every entry validates its arguments and reports a status instead of aborting */
  const int st = b200_msda_f16(NULL, NULL, NULL, NULL, NULL, batch, spatial, heads, channels, levels, queries, points,
                               groups, NULL, NULL);
  printf("b200_msda_f16 with no buffers -> status %d (%s)\n", st, b200_status_string(st));
  if (st != B200_ERR_BAD_PARAM) return 1;

  /* launch-shape switches: same bits for every setting; a host picks one by timing them once on its own tensors */
  const int shape_before = b200_msda_set_batch_units(0, 0); /* 0 = query */
  const int variant_before = b200_msda_set_gather_variant(-1);
  b200_msda_set_gather_variant(1);
  printf("MSDA launch shape: units %d%s, gather variant %d -> %d\n", shape_before & 0xff, (shape_before >> 8) ? " (strided)" : "",
         variant_before, b200_msda_set_gather_variant(variant_before));
  if (b200_msda_set_gather_variant(-1) != variant_before) return 1;

  printf("DCN workspace for the R101 stage-3 layer: %zu bytes\n",
         b200_dcn_workspace_size(1, 6, 256, 58, 100, 3, 3, 1, 1, 1, 1, 1, 1));
  printf("kernels launched by this process so far: %llu\n", b200_launch_count());
  return 0;
}
