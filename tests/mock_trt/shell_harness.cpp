// tests/mock_trt/shell_harness.cpp — TEST-ONLY. Drives the TensorRT plugin shells (csrc/trt_plugin/b200_trt_plugins.cpp)
// through the mock plugin API on a machine without TensorRT and without a GPU: creator fields, create / serialise /
// deserialise / clone round trips, output dimensions, format negotiation tables (checked against the reference's
// rules, file:line in the comments), workspace sizes, and that enqueue() forwards to the C ABI (bad arguments come back
// as a status, nothing is launched). Prints "OK <n checks>" on success.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../bevformer_tensorrt_b200/csrc/trt_plugin/b200_trt_plugins.cpp"

using namespace nvinfer1;
using b200_trt::Creator;
using b200_trt::Op;

static int g_checks = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    ++g_checks;                                                                  \
    if (!(cond)) {                                                               \
      std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                              \
    }                                                                            \
  } while (0)

static PluginTensorDesc desc(DataType t, TensorFormat f, std::initializer_list<int> dims, float scale = 1.f) {
  PluginTensorDesc d{};
  d.dims.nbDims = static_cast<int32_t>(dims.size());
  int i = 0;
  for (int v : dims) d.dims.d[i++] = v;
  d.type = t, d.format = f, d.scale = scale;
  return d;
}

static IPluginV2DynamicExt *make(Op op, bool v2, std::vector<PluginField> fields = {}) {
  Creator c(op, v2);
  PluginFieldCollection fc{static_cast<int32_t>(fields.size()), fields.data()};
  return static_cast<IPluginV2DynamicExt *>(c.createPlugin("layer", &fc));
}

static std::vector<char> blob(const IPluginV2 *p) {
  std::vector<char> b(p->getSerializationSize());
  if (!b.empty()) p->serialize(b.data());
  return b;
}

int main() {
  const auto F = DataType::kFLOAT, H = DataType::kHALF, I8 = DataType::kINT8, I32 = DataType::kINT32;
  const auto LIN = TensorFormat::kLINEAR, CHW2 = TensorFormat::kCHW2, CHW4 = TensorFormat::kCHW4;

  // ---- names / versions: the lookup keys of the reference plugins
  struct { Op op; bool v2; const char *name; } names[] = {
      {Op::kMSDA, false, "MultiScaleDeformableAttnTRT"},       {Op::kMSDA, true, "MultiScaleDeformableAttnTRT2"},
      {Op::kGridSampler2D, false, "GridSampler2DTRT"},         {Op::kGridSampler2D, true, "GridSampler2DTRT2"},
      {Op::kGridSampler3D, false, "GridSampler3DTRT"},         {Op::kGridSampler3D, true, "GridSampler3DTRT2"},
      {Op::kDCN, false, "ModulatedDeformableConv2dTRT"},       {Op::kDCN, true, "ModulatedDeformableConv2dTRT2"},
      {Op::kRotate, false, "RotateTRT"},                       {Op::kRotate, true, "RotateTRT2"}};
  for (auto &n : names) {
    Creator c(n.op, n.v2);
    CHECK(std::string(c.getPluginName()) == n.name);
    CHECK(std::string(c.getPluginVersion()) == "1");
    IPluginV2DynamicExt *p = make(n.op, n.v2);
    CHECK(p && std::string(p->getPluginType()) == n.name && p->getNbOutputs() == 1);
    IPluginV2DynamicExt *q = p->clone();
    CHECK(q && std::string(q->getPluginType()) == n.name && blob(p) == blob(q));
    DataType in_types[5] = {H, H, H, H, H};
    CHECK(p->getOutputDataType(0, in_types, 5) == H);
    q->destroy();
    p->destroy();
  }

  // ---- attributes survive create -> serialise -> deserialise (gridSamplerPlugin.cpp:157-166, …Conv2dPlugin.cpp:200-211)
  {
    const int32_t interp = 2, pad = 1, align = 1;
    IPluginV2DynamicExt *p = make(Op::kGridSampler2D, false, {{"interpolation_mode", &interp, PluginFieldType::kINT32, 1},
                                                               {"padding_mode", &pad, PluginFieldType::kINT32, 1},
                                                               {"align_corners", &align, PluginFieldType::kINT32, 1}});
    std::vector<char> b = blob(p);
    CHECK(!b.empty());
    Creator c(Op::kGridSampler2D, false);
    IPluginV2 *r = c.deserializePlugin("layer", b.data(), b.size());
    CHECK(r && blob(r) == b);
    IPluginV2DynamicExt *d0 = make(Op::kGridSampler2D, false);
    CHECK(blob(d0) != b);  // the attributes are really in there
    d0->destroy(), r->destroy(), p->destroy();
    const int32_t stride[2] = {2, 2}, padding[2] = {1, 1}, dil[2] = {1, 1}, groups = 2, dg = 4;
    IPluginV2DynamicExt *dcn = make(Op::kDCN, false, {{"stride", stride, PluginFieldType::kINT32, 2},
                                                       {"padding", padding, PluginFieldType::kINT32, 2},
                                                       {"dilation", dil, PluginFieldType::kINT32, 2},
                                                       {"groups", &groups, PluginFieldType::kINT32, 1},
                                                       {"deform_groups", &dg, PluginFieldType::kINT32, 1}});
    Creator cd(Op::kDCN, false);
    std::vector<char> bd = blob(dcn);
    IPluginV2 *rd = cd.deserializePlugin("layer", bd.data(), bd.size());
    CHECK(rd && blob(rd) == bd);
    rd->destroy(), dcn->destroy();
    CHECK(make(Op::kMSDA, false)->getSerializationSize() == 0);  // …Plugin.cpp:142-146
  }

  // ---- output dimensions are the reference's expressions
  {
    const IDimensionExpr *e[5][8];
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 8; ++j) e[i][j] = reinterpret_cast<const IDimensionExpr *>(static_cast<uintptr_t>(0x100 * (i + 1) + j));
    DimsExprs in[5];
    for (int i = 0; i < 5; ++i) {
      in[i].nbDims = 4;
      for (int j = 0; j < 8; ++j) in[i].d[j] = e[i][j];
    }
    IExprBuilder *eb = nullptr;
    IPluginV2DynamicExt *m = make(Op::kMSDA, false);  // [value.d0, offsets.d1, value.d2, value.d3] (…Plugin.cpp:48-58)
    DimsExprs o = m->getOutputDimensions(0, in, 5, *eb);
    CHECK(o.nbDims == 4 && o.d[0] == e[0][0] && o.d[1] == e[3][1] && o.d[2] == e[0][2] && o.d[3] == e[0][3]);
    IPluginV2DynamicExt *g = make(Op::kGridSampler2D, false);  // [in.d0, in.d1, grid.d2, grid.d3]
    o = g->getOutputDimensions(0, in, 2, *eb);
    CHECK(o.nbDims == 4 && o.d[0] == e[0][0] && o.d[1] == e[0][1] && o.d[2] == e[1][2] && o.d[3] == e[1][3]);
    IPluginV2DynamicExt *g3 = make(Op::kGridSampler3D, false);
    o = g3->getOutputDimensions(0, in, 2, *eb);
    CHECK(o.nbDims == 5 && o.d[1] == e[0][1] && o.d[4] == e[1][4]);
    IPluginV2DynamicExt *c = make(Op::kDCN, false);  // [in.d0, weight.d0, offset.d2, offset.d3]
    o = c->getOutputDimensions(0, in, 5, *eb);
    CHECK(o.nbDims == 4 && o.d[0] == e[0][0] && o.d[1] == e[3][0] && o.d[2] == e[1][2] && o.d[3] == e[1][3]);
    IPluginV2DynamicExt *r = make(Op::kRotate, false);  // img dims (rotatePlugin.cpp:52-62)
    o = r->getOutputDimensions(0, in, 3, *eb);
    CHECK(o.nbDims == 3 && o.d[0] == e[0][0] && o.d[1] == e[0][1] && o.d[2] == e[0][2]);
  }

  // ---- format negotiation
  {
    // grid sampler (gridSamplerPlugin.cpp:168-194): pos 0 decides, the rest must copy type and format
    IPluginV2DynamicExt *g = make(Op::kGridSampler2D, false), *g2 = make(Op::kGridSampler2D, true);
    PluginTensorDesc io[3] = {desc(H, LIN, {1, 8, 4, 4}), desc(H, LIN, {1, 2, 4, 4}), desc(H, LIN, {1, 8, 4, 4})};
    CHECK(g->supportsFormatCombination(0, io, 2, 1) && g->supportsFormatCombination(1, io, 2, 1) &&
          g->supportsFormatCombination(2, io, 2, 1));
    CHECK(!g2->supportsFormatCombination(0, io, 2, 1));  // …TRT2 wants kCHW2 for FP16
    io[0].format = io[1].format = io[2].format = CHW2;
    CHECK(g2->supportsFormatCombination(0, io, 2, 1) && g2->supportsFormatCombination(2, io, 2, 1));
    CHECK(!g->supportsFormatCombination(0, io, 2, 1));
    io[0] = desc(I8, CHW4, {1, 8, 4, 4}), io[1] = desc(I8, CHW4, {1, 2, 4, 4}), io[2] = desc(I8, LIN, {1, 8, 4, 4});
    CHECK(g->supportsFormatCombination(0, io, 2, 1) && g->supportsFormatCombination(1, io, 2, 1));
    CHECK(!g->supportsFormatCombination(2, io, 2, 1));  // output must be kCHW4 like the input
    io[0] = desc(F, LIN, {1, 8, 4, 4}), io[1] = desc(H, LIN, {1, 2, 4, 4});
    CHECK(g->supportsFormatCombination(0, io, 2, 1) && !g->supportsFormatCombination(1, io, 2, 1));
    IPluginV2DynamicExt *g3 = make(Op::kGridSampler3D, true);
    PluginTensorDesc io3[3] = {desc(H, CHW2, {1, 4, 3, 4, 4}), desc(H, CHW2, {1, 3, 3, 4, 4}), desc(H, CHW2, {1, 4, 3, 4, 4})};
    CHECK(!g3->supportsFormatCombination(0, io3, 2, 1));  // 3-D: linear only
    io3[0].format = LIN;
    CHECK(g3->supportsFormatCombination(0, io3, 2, 1));

    // rotate (rotatePlugin.cpp:122-153)
    IPluginV2DynamicExt *r = make(Op::kRotate, false), *r2 = make(Op::kRotate, true);
    PluginTensorDesc ro[4] = {desc(H, LIN, {8, 4, 4}), desc(H, LIN, {1}), desc(H, LIN, {2}), desc(H, LIN, {8, 4, 4})};
    for (int pos = 0; pos < 4; ++pos) CHECK(r->supportsFormatCombination(pos, ro, 3, 1));
    CHECK(!r2->supportsFormatCombination(0, ro, 3, 1));
    ro[0].format = ro[3].format = CHW2;
    CHECK(r2->supportsFormatCombination(0, ro, 3, 1) && r2->supportsFormatCombination(1, ro, 3, 1) &&
          r2->supportsFormatCombination(3, ro, 3, 1));
    ro[1].type = F;  // angle must have the image's type unless the image is INT8
    CHECK(!r2->supportsFormatCombination(1, ro, 3, 1));
    ro[0] = desc(I8, CHW4, {8, 4, 4}), ro[1] = desc(F, LIN, {1}), ro[2] = desc(F, LIN, {2}), ro[3] = desc(I8, CHW4, {8, 4, 4});
    for (int pos = 0; pos < 4; ++pos) CHECK(r->supportsFormatCombination(pos, ro, 3, 1));
    ro[2].type = H;  // center must match angle
    CHECK(!r->supportsFormatCombination(2, ro, 3, 1));

    // DCN (…Conv2dPlugin.cpp:213-250): INT8 = input & weight kCHW4, offset / mask / output int8 linear, bias fp32|fp16
    IPluginV2DynamicExt *c = make(Op::kDCN, false);
    PluginTensorDesc co[6] = {desc(I8, CHW4, {2, 64, 8, 8}), desc(I8, LIN, {2, 18, 8, 8}), desc(I8, LIN, {2, 9, 8, 8}),
                              desc(I8, CHW4, {128, 64, 3, 3}), desc(F, LIN, {128}), desc(I8, LIN, {2, 128, 8, 8})};
    for (int pos = 0; pos < 6; ++pos) CHECK(c->supportsFormatCombination(pos, co, 5, 1));
    co[4].type = H;
    CHECK(c->supportsFormatCombination(4, co, 5, 1));
    co[3].format = LIN;
    CHECK(!c->supportsFormatCombination(3, co, 5, 1));
    co[0] = desc(I8, CHW4, {2, 6, 8, 8});  // channels % 4 != 0: no INT8 (use_int8, :218-220)
    CHECK(!c->supportsFormatCombination(0, co, 5, 1));
    PluginTensorDesc cf[6] = {desc(H, LIN, {2, 64, 8, 8}), desc(H, LIN, {2, 18, 8, 8}), desc(H, LIN, {2, 9, 8, 8}),
                              desc(H, LIN, {128, 64, 3, 3}), desc(H, LIN, {128}), desc(H, LIN, {2, 128, 8, 8})};
    for (int pos = 0; pos < 6; ++pos) CHECK(c->supportsFormatCombination(pos, cf, 5, 1));
    cf[1].type = F;
    CHECK(!c->supportsFormatCombination(1, cf, 5, 1));
    // …TRT2 (use_h2, :222-250): FP16 input, offset and weight as kCHW2 packets; mask, bias and output stay linear
    IPluginV2DynamicExt *c2 = make(Op::kDCN, true);
    PluginTensorDesc ch[6] = {desc(H, CHW2, {2, 64, 8, 8}), desc(H, CHW2, {2, 18, 8, 8}), desc(H, LIN, {2, 9, 8, 8}),
                              desc(H, CHW2, {128, 64, 3, 3}), desc(H, LIN, {128}), desc(H, LIN, {2, 128, 8, 8})};
    for (int pos = 0; pos < 6; ++pos) CHECK(c2->supportsFormatCombination(pos, ch, 5, 1));
    ch[0].format = LIN;
    CHECK(!c2->supportsFormatCombination(0, ch, 5, 1));  // …TRT2 does not take linear FP16 input
    ch[0].format = CHW2, ch[2].format = CHW2;
    CHECK(!c2->supportsFormatCombination(2, ch, 5, 1));  // mask is linear
    ch[2].format = LIN, ch[5].format = CHW2;
    CHECK(!c2->supportsFormatCombination(5, ch, 5, 1));  // output is linear
    CHECK(!c->supportsFormatCombination(0, ch, 5, 1));    // …TRT (not 2) does not take kCHW2

    // MSDA goes through b200_msda_supports_format (…Plugin.cpp:148-189): shapes are int32, everything linear
    IPluginV2DynamicExt *m = make(Op::kMSDA, false);
    PluginTensorDesc mo[6] = {desc(H, LIN, {6, 375, 8, 32}), desc(I32, LIN, {1, 2}), desc(H, LIN, {6, 2500, 1, 8}),
                              desc(H, LIN, {6, 2500, 8, 16}), desc(H, LIN, {6, 2500, 8, 8}), desc(H, LIN, {6, 2500, 8, 32})};
    for (int pos = 0; pos < 6; ++pos) CHECK(m->supportsFormatCombination(pos, mo, 5, 1));
    mo[1].type = F;
    CHECK(!m->supportsFormatCombination(1, mo, 5, 1));
  }

  // ---- workspace: 0 for MSDA / grid sampler / rotate, the library's figure for DCN
  {
    PluginTensorDesc in[5] = {desc(H, LIN, {6, 256, 58, 100}), desc(H, LIN, {6, 18, 58, 100}), desc(H, LIN, {6, 9, 58, 100}),
                              desc(H, LIN, {256, 256, 3, 3}), desc(H, LIN, {256})};
    PluginTensorDesc out = desc(H, LIN, {6, 256, 58, 100});
    const int32_t one[2] = {1, 1};
    IPluginV2DynamicExt *c = make(Op::kDCN, false, {{"stride", one, PluginFieldType::kINT32, 2},
                                                     {"padding", one, PluginFieldType::kINT32, 2},
                                                     {"dilation", one, PluginFieldType::kINT32, 2}});
    CHECK(c->getWorkspaceSize(in, 5, &out, 1) == b200_dcn_workspace_size(1, 6, 256, 58, 100, 3, 3, 1, 1, 1, 1, 1, 1));
    in[0].format = CHW2;
    CHECK(c->getWorkspaceSize(in, 5, &out, 1) ==
          b200_dcn_f16_chw2_workspace_size(6, 256, 58, 100, 256, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1));
    in[0].format = LIN;
    in[0].type = I8;
    CHECK(c->getWorkspaceSize(in, 5, &out, 1) ==
          b200_dcn_i8_workspace_size(6, 256, 58, 100, 256, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1));
    {  // MSDA: room for the packed value stack of the second-generation INT8 kernels; 0 where they do not apply
      PluginTensorDesc base[5] = {desc(I8, LIN, {6, 30825, 8, 32}), desc(I32, LIN, {4, 2}), desc(H, LIN, {6, 40000, 1, 8}),
                                  desc(I8, LIN, {6, 40000, 8, 64}), desc(I8, LIN, {6, 40000, 8, 32})};
      IPluginV2DynamicExt *m = make(Op::kMSDA, false);
      CHECK(m->getWorkspaceSize(base, 5, &out, 1) == b200_msda_i8_workspace_size(6, 30825, 8, 32, 4, 8, 4));
      CHECK(m->getWorkspaceSize(base, 5, &out, 1) == size_t(3) * 30825 * 64 * 6 * 8);
      base[0].type = H;
      CHECK(m->getWorkspaceSize(base, 5, &out, 1) == 0);  // FP16 / FP32: round-1 kernel, no workspace (…Plugin.cpp:64-69)
      base[0].type = F;
      CHECK(m->getWorkspaceSize(base, 5, &out, 1) == 0);
      PluginTensorDesc tsa[5] = {desc(I8, LIN, {2, 40000, 8, 32}), desc(I32, LIN, {1, 2}), desc(H, LIN, {2, 40000, 1, 2}),
                                 desc(I8, LIN, {2, 40000, 8, 8}), desc(I8, LIN, {2, 40000, 8, 4})};
      CHECK(m->getWorkspaceSize(tsa, 5, &out, 1) == 0);  // 4 points: outside the v2 envelope
    }
    CHECK(make(Op::kGridSampler2D, false)->getWorkspaceSize(in, 2, &out, 1) == 0);
    CHECK(make(Op::kRotate, false)->getWorkspaceSize(in, 3, &out, 1) == 0);
  }

  // ---- enqueue forwards to the C ABI: missing buffers come back as a status (B200_ERR_BAD_PARAM), never a crash
  {
    const void *inputs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void *outputs[1] = {nullptr};
    PluginTensorDesc ro[3] = {desc(F, LIN, {8, 4, 4}), desc(F, LIN, {1}), desc(F, LIN, {2})};
    PluginTensorDesc rout = desc(F, LIN, {8, 4, 4});
    CHECK(make(Op::kRotate, false)->enqueue(ro, &rout, inputs, outputs, nullptr, nullptr) == B200_ERR_BAD_PARAM);
    PluginTensorDesc go[2] = {desc(F, LIN, {1, 8, 4, 4}), desc(F, LIN, {1, 2, 4, 4})};
    PluginTensorDesc gout = desc(F, LIN, {1, 8, 4, 4});
    CHECK(make(Op::kGridSampler2D, false)->enqueue(go, &gout, inputs, outputs, nullptr, nullptr) == B200_ERR_BAD_PARAM);
    PluginTensorDesc mo[5] = {desc(H, LIN, {6, 375, 8, 32}), desc(I32, LIN, {1, 2}), desc(H, LIN, {6, 2500, 1, 8}),
                              desc(H, LIN, {6, 2500, 8, 16}), desc(H, LIN, {6, 2500, 8, 8})};
    PluginTensorDesc mout = desc(H, LIN, {6, 2500, 8, 32});
    CHECK(make(Op::kMSDA, false)->enqueue(mo, &mout, inputs, outputs, nullptr, nullptr) != 0);
    PluginTensorDesc co[5] = {desc(H, LIN, {2, 64, 8, 8}), desc(H, LIN, {2, 18, 8, 8}), desc(H, LIN, {2, 9, 8, 8}),
                              desc(H, LIN, {128, 64, 3, 3}), desc(H, LIN, {128})};
    PluginTensorDesc cout = desc(H, LIN, {2, 128, 8, 8});
    IPluginV2DynamicExt *c = make(Op::kDCN, false);
    c->configurePlugin(nullptr, 4, nullptr, 1);  // no bias input: inputs[4] must not be touched
    CHECK(c->enqueue(co, &cout, inputs, outputs, nullptr, nullptr) == B200_ERR_BAD_PARAM);
  }
  std::printf("OK %d checks\n", g_checks);
  return 0;
}
