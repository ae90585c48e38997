// tests/mock_trt/NvInfer.h — TEST-ONLY, declaration-only stand-in for the subset of TensorRT 8.x's NvInfer.h that
// bevformer_tensorrt_b200/csrc/trt_plugin/b200_trt_plugins.cpp uses. It exists so the plugin shell can be compiled
// (syntax + override signatures + the PluginTensorDesc layout assertion) on a machine without TensorRT. Written from
// the public TensorRT API documentation; it is never shipped or linked.
#pragma once
#include <cstddef>
#include <cstdint>
struct CUstream_st;
typedef CUstream_st *cudaStream_t;
namespace nvinfer1 {
enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4 };
enum class TensorFormat : int32_t { kLINEAR = 0, kCHW2 = 1, kHWC8 = 2, kCHW4 = 3 };
enum class PluginFieldType : int32_t { kFLOAT16 = 0, kFLOAT32 = 1, kFLOAT64 = 2, kINT8 = 3, kINT16 = 4, kINT32 = 5 };
struct Dims { int32_t nbDims; int32_t d[8]; };
struct PluginTensorDesc { Dims dims; DataType type; TensorFormat format; float scale; };
struct DynamicPluginTensorDesc { PluginTensorDesc desc; Dims min; Dims max; };
class IDimensionExpr;
struct DimsExprs { int32_t nbDims; const IDimensionExpr *d[8]; };
class IExprBuilder;
struct PluginField { const char *name; const void *data; PluginFieldType type; int32_t length; };
struct PluginFieldCollection { int32_t nbFields; const PluginField *fields; };
class IPluginV2 {
 public:
  virtual const char *getPluginType() const noexcept = 0;
  virtual const char *getPluginVersion() const noexcept = 0;
  virtual int32_t getNbOutputs() const noexcept = 0;
  virtual int32_t initialize() noexcept = 0;
  virtual void terminate() noexcept = 0;
  virtual size_t getSerializationSize() const noexcept = 0;
  virtual void serialize(void *buffer) const noexcept = 0;
  virtual void destroy() noexcept = 0;
  virtual void setPluginNamespace(const char *ns) noexcept = 0;
  virtual const char *getPluginNamespace() const noexcept = 0;
  virtual ~IPluginV2() = default;
};
class IPluginV2Ext : public IPluginV2 {
 public:
  virtual DataType getOutputDataType(int32_t index, const DataType *inputTypes, int32_t nbInputs) const noexcept = 0;
};
class IPluginV2DynamicExt : public IPluginV2Ext {
 public:
  virtual IPluginV2DynamicExt *clone() const noexcept = 0;
  virtual DimsExprs getOutputDimensions(int32_t outputIndex, const DimsExprs *inputs, int32_t nbInputs,
                                        IExprBuilder &exprBuilder) noexcept = 0;
  virtual bool supportsFormatCombination(int32_t pos, const PluginTensorDesc *inOut, int32_t nbInputs,
                                         int32_t nbOutputs) noexcept = 0;
  virtual void configurePlugin(const DynamicPluginTensorDesc *in, int32_t nbInputs, const DynamicPluginTensorDesc *out,
                               int32_t nbOutputs) noexcept = 0;
  virtual size_t getWorkspaceSize(const PluginTensorDesc *inputs, int32_t nbInputs, const PluginTensorDesc *outputs,
                                  int32_t nbOutputs) const noexcept = 0;
  virtual int32_t enqueue(const PluginTensorDesc *inputDesc, const PluginTensorDesc *outputDesc,
                          const void *const *inputs, void *const *outputs, void *workspace,
                          cudaStream_t stream) noexcept = 0;
};
class IPluginCreator {
 public:
  virtual const char *getPluginName() const noexcept = 0;
  virtual const char *getPluginVersion() const noexcept = 0;
  virtual const PluginFieldCollection *getFieldNames() noexcept = 0;
  virtual IPluginV2 *createPlugin(const char *name, const PluginFieldCollection *fc) noexcept = 0;
  virtual IPluginV2 *deserializePlugin(const char *name, const void *serialData, size_t serialLength) noexcept = 0;
  virtual void setPluginNamespace(const char *ns) noexcept = 0;
  virtual const char *getPluginNamespace() const noexcept = 0;
  virtual ~IPluginCreator() = default;
};
}  // namespace nvinfer1
#define REGISTER_TENSORRT_PLUGIN(name) static name pluginRegistrar##name {}
