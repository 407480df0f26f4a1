// Engine + ops: the B200 implementation of the reference's plugin surface (core/engine.h:35-125,
// core/op.h:12-51). One Engine = one GPU + one stream; every op is a thin C++ object that owns
// launch parameters and calls the kernel-level C ABI (include/oidn_b200_kernels.h) from
// submitKernels(). Names and argument meaning follow the reference ops:
//   Conv (core/conv.h:26-61), ConcatConv (core/concat_conv.h), Pool (core/pool.h),
//   Upsample (core/upsample.h), InputProcess (core/input_process.h), OutputProcess
//   (core/output_process.h), Autoexposure (core/autoexposure.h), ImageCopy (core/image_copy.h).
#pragma once
#include "base.hpp"
#include "../../../include/oidn_b200_kernels.h"
#include <functional>
#include <vector>

namespace oidnb200 {

class Engine;

// TransferFunction (core/color.h:10-166): type + input scale (value or device pointer)
struct TransferFunction
{
  explicit TransferFunction(TransferType type = TransferType::Linear) : type(type) {}
  TransferType type;
  const float* inputScalePtr = nullptr;
  float inputScale = 1.f;
  void setInputScale(float s) { inputScalePtr = nullptr; inputScale = s; }
  void setInputScale(const float* p) { inputScalePtr = p; inputScale = 1.f; }
  oidnb200_transfer abi() const { return oidnb200_transfer{(int)type, inputScale, inputScalePtr}; }
};

// Tensor in the scratch arena: NHWC fp16, padded channels ("hwc", blockC = 16).
struct TensorDesc
{
  int C = 0;       // logical channels
  int H = 0, W = 0;
  int paddedC() const { return round_up(C, 16); }
  size_t byteSize() const { return (size_t)H * W * paddedC() * 2; }
};

class Op
{
public:
  virtual ~Op() = default;
  virtual void submitKernels() = 0;
  virtual void finalize() {}
  void setName(const std::string& n) { name = n; }
  const std::string& getName() const { return name; }
  void submit() { submitKernels(); }

protected:
  std::string name;
};

// ConvDesc: src(s), activation, post-op (core/conv.h:26-35). src2.C > 0 makes the op a ConcatConv
// (core/concat_conv.h) that reads both tensors in place. src1Upsampled = the producer's
// PostOp::Upsample was folded into this op's loader (src1 is stored at H/2 x W/2).
struct ConvDesc
{
  TensorDesc src1;
  TensorDesc src2;
  int outC = 0;               // logical output channels
  Activation activation = Activation::ReLU;
  PostOp postOp = PostOp::None;
  bool src1Upsampled = false;
  int H = 0, W = 0;           // resolution the convolution runs at
};

class Conv : public Op
{
public:
  Conv(Engine* engine, const ConvDesc& desc);
  ~Conv() override;
  const ConvDesc& getDesc() const { return desc; }
  TensorDesc getDstDesc() const; // stored dst (pooled: H/2 x W/2; upsample: stays H x W, consumer reads it 2x)
  size_t getWeightByteSize() const;
  size_t getBiasByteSize() const;
  // host-side reorder of TZA tensors into the device layout (core/tensor_reorder.cpp:8-98)
  void packWeight(const uint16_t* oihw, int O, int I1, int I2, void* dstHost) const;
  void packBias(const uint16_t* x, int O, void* dstHost) const;
  void setSrc(const void* s1, const void* s2) { src1 = s1; src2 = s2; bound = false; }
  void setWeight(const void* w) { weight = w; bound = false; }
  void setBias(const void* b) { bias = b; bound = false; }
  void setDst(void* d) { dst = d; bound = false; }
  void finalize() override;      // encodes the TMA tensor maps
  void submitKernels() override;
  oidnb200_conv_info getInfo() const;
  oidnb200_conv* getHandle() const { return handle; }
  Engine* getEngine() const { return engine; }
  bool isBound() const { return bound; }

private:
  Engine* engine;
  ConvDesc desc;
  oidnb200_conv* handle = nullptr;
  const void *src1 = nullptr, *src2 = nullptr, *weight = nullptr, *bias = nullptr;
  void* dst = nullptr;
  bool bound = false;
};

// Two chained convs as one launch (kernels/conv_pair_tc.cu): the tensor between them stays in shared memory.
// Built by the graph over two Conv ops it keeps (they own weights and bindings); submit() of the pair replaces
// the two launches.
class ConvPair
{
public:
  // nullptr when the kernel does not cover the two shapes
  static std::unique_ptr<ConvPair> tryCreate(Conv& a, Conv& b);
  ~ConvPair();
  void submit();
  oidnb200_conv_info getInfo() const;
private:
  ConvPair() = default;
  Conv* a = nullptr; Conv* b = nullptr;
  oidnb200_conv_pair* handle = nullptr;
};

class Pool : public Op
{
public:
  Pool(Engine* engine, const TensorDesc& src) : engine(engine), srcDesc(src) {}
  TensorDesc getDstDesc() const { return TensorDesc{srcDesc.C, srcDesc.H / 2, srcDesc.W / 2}; }
  void setSrc(const void* s) { src = s; }
  void setDst(void* d) { dst = d; }
  void submitKernels() override;
private:
  Engine* engine; TensorDesc srcDesc; const void* src = nullptr; void* dst = nullptr;
};

class Upsample : public Op
{
public:
  Upsample(Engine* engine, const TensorDesc& src) : engine(engine), srcDesc(src) {}
  TensorDesc getDstDesc() const { return TensorDesc{srcDesc.C, srcDesc.H * 2, srcDesc.W * 2}; }
  void setSrc(const void* s) { src = s; }
  void setDst(void* d) { dst = d; }
  void submitKernels() override;
private:
  Engine* engine; TensorDesc srcDesc; const void* src = nullptr; void* dst = nullptr;
};

class InputProcess : public Op
{
public:
  InputProcess(Engine* engine, const TensorDesc& dstDesc, std::shared_ptr<TransferFunction> tf, bool hdr, bool snorm)
    : engine(engine), dstDesc(dstDesc), transferFunc(std::move(tf)), hdr(hdr), snorm(snorm) {}
  TensorDesc getDstDesc() const { return dstDesc; }
  // core/input_process.cpp: main input = color, else albedo, else normal
  void setSrc(const Image& color, const Image& albedo, const Image& normal);
  void setTile(int hSrc, int wSrc, int hDst, int wDst, int H, int W) { tile = oidnb200_tile{hSrc, wSrc, hDst, wDst, H, W}; }
  void setDst(void* d) { dst = d; }
  void submitKernels() override;
private:
  Engine* engine; TensorDesc dstDesc; std::shared_ptr<TransferFunction> transferFunc; bool hdr, snorm;
  Image input, albedo, normal; oidnb200_tile tile{}; void* dst = nullptr;
};

class OutputProcess : public Op
{
public:
  OutputProcess(Engine* engine, const TensorDesc& srcDesc, std::shared_ptr<TransferFunction> tf, bool hdr, bool snorm)
    : engine(engine), srcDesc(srcDesc), transferFunc(std::move(tf)), hdr(hdr), snorm(snorm) {}
  void setSrc(const void* s) { src = s; }
  void setDst(const Image& image) { dst = image; }
  void setTile(int hSrc, int wSrc, int hDst, int wDst, int H, int W) { tile = oidnb200_tile{hSrc, wSrc, hDst, wDst, H, W}; }
  void submitKernels() override;
  // Folds this op (current tile, destination, transfer function) into the epilogue of the conv that
  // produces its source, or removes the fusion (on = false). Returns false when the kernel ABI does
  // not support the combination (image format): the conv is then left unfused and this op must be
  // submitted as a pass of its own.
  bool fuseInto(Conv& producer, bool on);
private:
  Engine* engine; TensorDesc srcDesc; std::shared_ptr<TransferFunction> transferFunc; bool hdr, snorm;
  const void* src = nullptr; Image dst; oidnb200_tile tile{};
};

class Autoexposure : public Op
{
public:
  Autoexposure(Engine* engine, int H, int W) : engine(engine), H(H), W(W) {}
  size_t getScratchByteSize() const { return oidnb200_autoexposure_scratch_bytes(H, W); }
  void setScratch(void* s) { scratch = s; }
  void setSrc(const Image& image);
  void setDst(float* d) { dst = d; }
  float* getDstPtr() const { return dst; }
  void submitKernels() override;
private:
  Engine* engine; int H, W; Image src; void* scratch = nullptr; float* dst = nullptr;
};

class ImageCopy : public Op
{
public:
  explicit ImageCopy(Engine* engine) : engine(engine) {}
  void setSrc(const Image& s) { src = s; }
  void setDst(const Image& d) { dst = d; }
  void submitKernels() override;
private:
  Engine* engine; Image src, dst;
};

// One GPU + one stream. Factory methods mirror Engine::new<Op> (core/engine.h:67-75).
class Engine
{
public:
  Engine(int deviceID, void* userStream);
  ~Engine();
  int getDeviceID() const { return deviceID; }
  // The stream ops launch on: the engine's main stream (the user's, or its own), or -- while a staged frame is
  // being enqueued (UNetFilter, device parameter "staging") -- the engine's internal compute stream.
  void* getStream() const { return active ? active : stream; }
  void* getMainStream() const { return stream; }
  bool ownsStream() const { return ownStream; }
  void setActiveStream(void* s) { active = s; }
  // Internal streams of the staging pipeline (created on first use): tile copy-in, compute, tile copy-out.
  enum AuxStream { CopyIn = 0, Compute = 1, CopyOut = 2 };
  void* getAuxStream(AuxStream which);
  void* newEvent();                 // cudaEvent_t without timing, owned by the engine
  void makeCurrent() const;

  std::shared_ptr<Conv> newConv(const ConvDesc& desc) { return std::make_shared<Conv>(this, desc); }
  std::shared_ptr<Pool> newPool(const TensorDesc& src) { return std::make_shared<Pool>(this, src); }
  std::shared_ptr<Upsample> newUpsample(const TensorDesc& src) { return std::make_shared<Upsample>(this, src); }
  std::shared_ptr<InputProcess> newInputProcess(const TensorDesc& dst, std::shared_ptr<TransferFunction> tf, bool hdr, bool snorm)
  { return std::make_shared<InputProcess>(this, dst, std::move(tf), hdr, snorm); }
  std::shared_ptr<OutputProcess> newOutputProcess(const TensorDesc& src, std::shared_ptr<TransferFunction> tf, bool hdr, bool snorm)
  { return std::make_shared<OutputProcess>(this, src, std::move(tf), hdr, snorm); }
  std::shared_ptr<Autoexposure> newAutoexposure(int H, int W) { return std::make_shared<Autoexposure>(this, H, W); }
  std::shared_ptr<ImageCopy> newImageCopy() { return std::make_shared<ImageCopy>(this); }
  bool isConvSupported(PostOp) const { return true; } // pool fused in the epilogue, upsample in the consumer's loader

  // memory (Engine::usmAlloc/usmFree/usmCopy, core/engine.h:78-91)
  void* malloc(size_t bytes, Storage storage = Storage::Device);
  void free(void* ptr, Storage storage = Storage::Device);
  void submitCopy(void* dst, const void* src, size_t bytes); // async on the engine's stream
  void submitCopy2D(void* dst, size_t dstPitch, const void* src, size_t srcPitch, size_t widthBytes, size_t height);
  void submitHostFunc(std::function<void()>&& f);
  void wait();                      // main stream and the internal streams

private:
  int deviceID;
  void* stream = nullptr;
  void* active = nullptr;
  void* aux[3] = {nullptr, nullptr, nullptr};
  std::vector<void*> events;
  bool ownStream = false;
};

void checkABI(int rc, const char* what);   // kernel-ABI return code -> Exception
void checkCuda(int cudaError, const char* what);

} // namespace oidnb200
