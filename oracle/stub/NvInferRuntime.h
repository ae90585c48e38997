// Test-only stand-in for TensorRT's NvInferRuntime.h (TensorRT is not installed in this image).
// It declares just enough of namespace nvinfer1 for the reference's common/helper.h
// (/root/reference/TensorRT/common/helper.h:13,27-41) to parse, so that the reference *kernel*
// translation units can be compiled as a GPU-side oracle (oracle/_ref). Nothing here is shipped
// in the product library.
#pragma once
#include <cstdint>
namespace nvinfer1 {
enum class DataType : int32_t { kFLOAT = 0, kHALF = 1, kINT8 = 2, kINT32 = 3, kBOOL = 4 };
class IPluginCreator {
public:
  virtual void setPluginNamespace(const char *ns) noexcept = 0;
  virtual const char *getPluginNamespace() const noexcept = 0;
  virtual ~IPluginCreator() = default;
};
} // namespace nvinfer1
