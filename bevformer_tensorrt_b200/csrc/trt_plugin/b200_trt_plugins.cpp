// b200_trt_plugins.cpp — TensorRT plugin shells over libb200_bev_ops (compiled only where TensorRT's NvInfer.h exists;
// this image has no TensorRT, so the build in bevformer_tensorrt_b200/build.py skips it and tests/test_host_logic.py
// compiles it against tests/mock_trt/NvInfer.h, a declaration-only stand-in, to keep it honest).
//
// Same plugin names / versions / input orders / attributes as the reference, so an ONNX graph exported by the
// reference's symbolic() functions resolves to these plugins unchanged:
//   MultiScaleDeformableAttnTRT, MultiScaleDeformableAttnTRT2   (multiScaleDeformableAttnPlugin.cpp:19-23, :345-346)
//   GridSampler2DTRT, GridSampler2DTRT2, GridSampler3DTRT, GridSampler3DTRT2   (gridSamplerPlugin.cpp:20-26, :558-561)
//   ModulatedDeformableConv2dTRT, ModulatedDeformableConv2dTRT2   (modulatedDeformableConv2dPlugin.cpp:19-23, :515-516)
//   RotateTRT, RotateTRT2                                         (rotatePlugin.cpp:19-21, :216-330)
// enqueue() forwards to the C ABI (include/b200_bev_ops.h); nothing is computed here.
#if __has_include(<NvInfer.h>)
#include <NvInfer.h>

#include <cstring>
#include <string>
#include <vector>

#include "b200_bev_ops.h"

namespace b200_trt {
using namespace nvinfer1;

static_assert(sizeof(b200_tensor_desc) == sizeof(PluginTensorDesc), "b200_tensor_desc must mirror PluginTensorDesc");

enum class Op { kMSDA, kGridSampler2D, kGridSampler3D, kDCN, kRotate };

struct Attrs {  // serialised verbatim (the reference serialises the same fields: gridSamplerPlugin.cpp:157-166,
                // modulatedDeformableConv2dPlugin.cpp:200-211; MSDA serialises nothing, …Plugin.cpp:142-146)
  int32_t interp = 0, padding = 0, align = 0;                           // grid sampler; rotate uses interp only
  int32_t stride[2] = {1, 1}, pad[2] = {0, 0}, dil[2] = {1, 1}, groups = 1, deform_groups = 1;  // DCN
};

class Plugin final : public IPluginV2DynamicExt {
 public:
  Plugin(Op op, bool v2, const Attrs &a) : op_(op), v2_(v2), a_(a) {}

  // ---- IPluginV2DynamicExt
  IPluginV2DynamicExt *clone() const noexcept override {
    auto *p = new (std::nothrow) Plugin(op_, v2_, a_);
    if (p) p->setPluginNamespace(ns_.c_str());
    return p;
  }
  DimsExprs getOutputDimensions(int32_t, const DimsExprs *in, int32_t nb, IExprBuilder &) noexcept override {
    DimsExprs o{};
    o.nbDims = 4;
    if (op_ == Op::kRotate) {  // output = img dims [C, H, W] (rotatePlugin.cpp:52-62)
      o.nbDims = 3;
      o.d[0] = in[0].d[0], o.d[1] = in[0].d[1], o.d[2] = in[0].d[2];
    } else if (op_ == Op::kMSDA) {  // [value.d0, offsets.d1, value.d2, value.d3] (…Plugin.cpp:48-58)
      o.d[0] = in[0].d[0], o.d[1] = in[3].d[1], o.d[2] = in[0].d[2], o.d[3] = in[0].d[3];
    } else if (op_ == Op::kGridSampler2D) {  // [in.d0, in.d1, grid.d2, grid.d3] (gridSamplerPlugin.cpp:85-96)
      o.d[0] = in[0].d[0], o.d[1] = in[0].d[1], o.d[2] = in[1].d[2], o.d[3] = in[1].d[3];
    } else if (op_ == Op::kGridSampler3D) {  // [in.d0, in.d1, grid.d2, grid.d3, grid.d4]
      o.nbDims = 5;
      o.d[0] = in[0].d[0], o.d[1] = in[0].d[1], o.d[2] = in[1].d[2], o.d[3] = in[1].d[3], o.d[4] = in[1].d[4];
    } else {  // [in.d0, weight.d0, offset.d2, offset.d3] (…Conv2dPlugin.cpp:60-66)
      o.d[0] = in[0].d[0], o.d[1] = in[3].d[0], o.d[2] = in[1].d[2], o.d[3] = in[1].d[3];
    }
    (void)nb;
    return o;
  }
  bool supportsFormatCombination(int32_t pos, const PluginTensorDesc *io, int32_t nbIn, int32_t nbOut) noexcept override {
    if (op_ == Op::kMSDA)
      return b200_msda_supports_format(pos, reinterpret_cast<const b200_tensor_desc *>(io), nbIn, nbOut) != 0;
    if (op_ == Op::kRotate) {  // rotatePlugin.cpp:122-153: img fp32/fp16 linear (kCHW2 for …TRT2) or int8 kCHW4
      const PluginTensorDesc &d = io[pos], &img = io[0];
      if (pos == 0) {
        if (d.type == DataType::kINT8) return d.format == TensorFormat::kCHW4;
        if (d.type == DataType::kHALF) return d.format == (v2_ ? TensorFormat::kCHW2 : TensorFormat::kLINEAR);
        return d.type == DataType::kFLOAT && d.format == TensorFormat::kLINEAR;
      }
      if (pos == nbIn) return d.type == img.type && d.format == img.format;
      if (d.format != TensorFormat::kLINEAR) return false;  // angle, center
      if (img.type == DataType::kINT8)
        return (d.type == DataType::kFLOAT || d.type == DataType::kHALF) && d.type == io[1].type;
      return d.type == img.type;
    }
    const PluginTensorDesc &d = io[pos], &x = io[0];
    if (op_ == Op::kGridSampler2D || op_ == Op::kGridSampler3D)  // gridSamplerPlugin.cpp:168-194
      return b200_grid_sampler_supports_format(pos, reinterpret_cast<const b200_tensor_desc *>(io), nbIn, nbOut, v2_) != 0;
    // DCN (…Conv2dPlugin.cpp:213-250). …TRT: fp32 / fp16 linear everywhere. …TRT2 (use_h2): FP16 input, offset and weight
    // as kCHW2 packets, mask / bias / output linear. INT8 (both): input and weight kCHW4, offset / mask / output int8
    // linear, bias fp32 or fp16 linear.
    const bool int8_ok = x.dims.d[1] % 4 == 0 && (io[nbIn].dims.d[1] / a_.groups) % 4 == 0;
    if (pos == 0) {
      if (d.type == DataType::kINT8) return d.format == TensorFormat::kCHW4 && int8_ok;
      if (d.type == DataType::kHALF) return d.format == (v2_ ? TensorFormat::kCHW2 : TensorFormat::kLINEAR);
      return d.type == DataType::kFLOAT && d.format == TensorFormat::kLINEAR;
    }
    if (x.type == DataType::kINT8) {
      if (nbIn == 5 && pos == 4)
        return (d.type == DataType::kFLOAT || d.type == DataType::kHALF) && d.format == TensorFormat::kLINEAR;
      if (pos == 3) return d.type == DataType::kINT8 && d.format == TensorFormat::kCHW4;
      return d.type == DataType::kINT8 && d.format == TensorFormat::kLINEAR;  // offset, mask, output
    }
    (void)nbOut;
    if ((nbIn == 5 && pos == 4) || pos == nbIn || pos == 2) return d.type == x.type && d.format == TensorFormat::kLINEAR;
    return d.type == x.type && d.format == x.format;  // offset, weight follow the input's format (…Conv2dPlugin.cpp:246-249)
  }
  void configurePlugin(const DynamicPluginTensorDesc *, int32_t nbIn, const DynamicPluginTensorDesc *, int32_t) noexcept override {
    nb_inputs_ = nbIn;  // DCN: 4 inputs without bias, 5 with (…Conv2dPlugin.cpp:297-299)
  }
  size_t getWorkspaceSize(const PluginTensorDesc *in, int32_t, const PluginTensorDesc *, int32_t) const noexcept override {
    if (op_ == Op::kMSDA)  // the reference asks for 0 (…Plugin.cpp:64-69); the v2 kernels want room for the packed value stack
      return b200_msda_enqueue_workspace_size(reinterpret_cast<const b200_tensor_desc *>(in));
    if (op_ != Op::kDCN) return 0;  // grid sampler / rotate need none
    return b200_dcn_enqueue_workspace_size(reinterpret_cast<const b200_tensor_desc *>(in), a_.stride, a_.pad, a_.dil,
                                           a_.groups, a_.deform_groups);
  }
  int32_t enqueue(const PluginTensorDesc *in, const PluginTensorDesc *out, const void *const *inputs, void *const *outputs,
                  void *workspace, cudaStream_t stream) noexcept override {
    if (op_ == Op::kMSDA)
      return b200_msda_enqueue(reinterpret_cast<const b200_tensor_desc *>(in),
                               reinterpret_cast<const b200_tensor_desc *>(out), inputs, outputs, workspace, stream, v2_);
    if (op_ == Op::kRotate) {  // inputs: img, angle, center (rotatePlugin.cpp:75-113)
      const int *dims = in[0].dims.d;
      if (in[0].type == DataType::kFLOAT)
        return b200_rotate_f32(static_cast<float *>(outputs[0]), static_cast<const float *>(inputs[0]),
                               static_cast<const float *>(inputs[1]), static_cast<const float *>(inputs[2]), dims,
                               a_.interp, stream);
      if (in[0].type == DataType::kHALF)
        return (v2_ ? b200_rotate_f16_h2 : b200_rotate_f16)(outputs[0], inputs[0], inputs[1], inputs[2], dims, a_.interp,
                                                            stream);
      if (in[0].type == DataType::kINT8)
        return b200_rotate_i8(static_cast<int8_t *>(outputs[0]), out[0].scale, static_cast<const int8_t *>(inputs[0]),
                              in[0].scale, inputs[1], inputs[2], in[1].type == DataType::kHALF, dims, a_.interp, stream);
      return 1;
    }
    if (op_ == Op::kGridSampler2D || op_ == Op::kGridSampler3D)
      return b200_grid_sampler_enqueue(reinterpret_cast<const b200_tensor_desc *>(in),
                                       reinterpret_cast<const b200_tensor_desc *>(out), inputs, outputs, workspace, stream,
                                       a_.interp, a_.padding, a_.align);
    // DCN: inputs x, offset, mask, weight[, bias]; attribute element [0] -> *_w slot, [1] -> *_h, as the reference's call
    return b200_dcn_enqueue(reinterpret_cast<const b200_tensor_desc *>(in), reinterpret_cast<const b200_tensor_desc *>(out),
                            inputs, outputs, workspace, stream, nb_inputs_, a_.stride, a_.pad, a_.dil, a_.groups,
                            a_.deform_groups);
  }
  // ---- IPluginV2Ext / IPluginV2
  DataType getOutputDataType(int32_t, const DataType *types, int32_t) const noexcept override { return types[0]; }
  const char *getPluginType() const noexcept override {
    switch (op_) {
      case Op::kMSDA: return v2_ ? "MultiScaleDeformableAttnTRT2" : "MultiScaleDeformableAttnTRT";
      case Op::kGridSampler2D: return v2_ ? "GridSampler2DTRT2" : "GridSampler2DTRT";
      case Op::kGridSampler3D: return v2_ ? "GridSampler3DTRT2" : "GridSampler3DTRT";
      case Op::kRotate: return v2_ ? "RotateTRT2" : "RotateTRT";
      default: return v2_ ? "ModulatedDeformableConv2dTRT2" : "ModulatedDeformableConv2dTRT";
    }
  }
  const char *getPluginVersion() const noexcept override { return "1"; }
  int32_t getNbOutputs() const noexcept override { return 1; }
  int32_t initialize() noexcept override { return 0; }
  void terminate() noexcept override {}
  size_t getSerializationSize() const noexcept override { return op_ == Op::kMSDA ? 0 : sizeof(Attrs) + sizeof(int32_t); }
  void serialize(void *buf) const noexcept override {
    if (op_ == Op::kMSDA) return;
    std::memcpy(buf, &a_, sizeof(Attrs));
    std::memcpy(static_cast<char *>(buf) + sizeof(Attrs), &nb_inputs_, sizeof(int32_t));
  }
  void destroy() noexcept override { delete this; }
  void setPluginNamespace(const char *ns) noexcept override { ns_ = ns ? ns : ""; }
  const char *getPluginNamespace() const noexcept override { return ns_.c_str(); }
  void set_nb_inputs(int32_t n) { nb_inputs_ = n; }

 private:
  Op op_;
  bool v2_;
  Attrs a_;
  int32_t nb_inputs_ = 5;
  std::string ns_;
};

class Creator final : public IPluginCreator {
 public:
  Creator(Op op, bool v2) : op_(op), v2_(v2) {
    if (op == Op::kGridSampler2D || op == Op::kGridSampler3D) {
      fields_ = {{"interpolation_mode", nullptr, PluginFieldType::kINT32, 1},
                 {"padding_mode", nullptr, PluginFieldType::kINT32, 1},
                 {"align_corners", nullptr, PluginFieldType::kINT32, 1}};
    } else if (op == Op::kRotate) {
      fields_ = {{"interpolation", nullptr, PluginFieldType::kINT32, 1}};
    } else if (op == Op::kDCN) {
      fields_ = {{"stride", nullptr, PluginFieldType::kINT32, 2},   {"padding", nullptr, PluginFieldType::kINT32, 2},
                 {"dilation", nullptr, PluginFieldType::kINT32, 2}, {"groups", nullptr, PluginFieldType::kINT32, 1},
                 {"deform_groups", nullptr, PluginFieldType::kINT32, 1}};
    }
    fc_.nbFields = static_cast<int32_t>(fields_.size());
    fc_.fields = fields_.data();
  }
  const char *getPluginName() const noexcept override { return Plugin(op_, v2_, Attrs{}).getPluginType(); }
  const char *getPluginVersion() const noexcept override { return "1"; }
  const PluginFieldCollection *getFieldNames() noexcept override { return &fc_; }
  IPluginV2 *createPlugin(const char *, const PluginFieldCollection *fc) noexcept override {
    Attrs a;
    for (int32_t i = 0; fc && i < fc->nbFields; ++i) {
      const PluginField &f = fc->fields[i];
      const auto *v = static_cast<const int32_t *>(f.data);
      if (!v || !f.name) continue;
      const std::string n = f.name;
      if (n == "interpolation_mode" || n == "interpolation") a.interp = v[0];
      else if (n == "padding_mode") a.padding = v[0];
      else if (n == "align_corners") a.align = v[0];
      else if (n == "stride") a.stride[0] = v[0], a.stride[1] = f.length > 1 ? v[1] : v[0];
      else if (n == "padding") a.pad[0] = v[0], a.pad[1] = f.length > 1 ? v[1] : v[0];
      else if (n == "dilation") a.dil[0] = v[0], a.dil[1] = f.length > 1 ? v[1] : v[0];
      else if (n == "groups") a.groups = v[0];
      else if (n == "deform_groups") a.deform_groups = v[0];
    }
    return new (std::nothrow) Plugin(op_, v2_, a);
  }
  IPluginV2 *deserializePlugin(const char *, const void *data, size_t len) noexcept override {
    Attrs a;
    int32_t nb = 5;
    if (op_ != Op::kMSDA && len >= sizeof(Attrs) + sizeof(int32_t)) {
      std::memcpy(&a, data, sizeof(Attrs));
      std::memcpy(&nb, static_cast<const char *>(data) + sizeof(Attrs), sizeof(int32_t));
    }
    auto *p = new (std::nothrow) Plugin(op_, v2_, a);
    if (p) p->set_nb_inputs(nb);
    return p;
  }
  void setPluginNamespace(const char *ns) noexcept override { ns_ = ns ? ns : ""; }
  const char *getPluginNamespace() const noexcept override { return ns_.c_str(); }

 private:
  Op op_;
  bool v2_;
  std::vector<PluginField> fields_;
  PluginFieldCollection fc_{};
  std::string ns_;
};

#define B200_REGISTER(cls_name, op, v2)                  \
  class cls_name final : public IPluginCreator {         \
   public:                                               \
    cls_name() : c_(op, v2) {}                           \
    const char *getPluginName() const noexcept override { return c_.getPluginName(); }                          \
    const char *getPluginVersion() const noexcept override { return c_.getPluginVersion(); }                    \
    const PluginFieldCollection *getFieldNames() noexcept override { return c_.getFieldNames(); }               \
    IPluginV2 *createPlugin(const char *n, const PluginFieldCollection *f) noexcept override { return c_.createPlugin(n, f); } \
    IPluginV2 *deserializePlugin(const char *n, const void *d, size_t l) noexcept override { return c_.deserializePlugin(n, d, l); } \
    void setPluginNamespace(const char *ns) noexcept override { c_.setPluginNamespace(ns); }                    \
    const char *getPluginNamespace() const noexcept override { return c_.getPluginNamespace(); }                \
   private:                                              \
    Creator c_;                                          \
  };                                                     \
  REGISTER_TENSORRT_PLUGIN(cls_name)

B200_REGISTER(MsdaCreator, Op::kMSDA, false);
B200_REGISTER(MsdaCreator2, Op::kMSDA, true);
B200_REGISTER(GridSampler2DCreator, Op::kGridSampler2D, false);
B200_REGISTER(GridSampler2DCreator2, Op::kGridSampler2D, true);
B200_REGISTER(GridSampler3DCreator, Op::kGridSampler3D, false);
B200_REGISTER(GridSampler3DCreator2, Op::kGridSampler3D, true);
B200_REGISTER(DcnCreator, Op::kDCN, false);
B200_REGISTER(DcnCreator2, Op::kDCN, true);
B200_REGISTER(RotateCreator, Op::kRotate, false);
B200_REGISTER(RotateCreator2, Op::kRotate, true);

}  // namespace b200_trt
#endif  // __has_include(<NvInfer.h>)
