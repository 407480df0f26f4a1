#include "engine.hpp"
#include <cuda_runtime.h>

namespace oidnb200 {

void checkCuda(int e, const char* what)
{
  if (e == cudaSuccess) return;
  const std::string msg = std::string(what) + ": " + cudaGetErrorString((cudaError_t)e);
  cudaGetLastError();
  // same mapping as the reference CUDA device (devices/cuda/cuda_device.cpp:34-51)
  if (e == cudaErrorMemoryAllocation) throw Exception(Error::OutOfMemory, msg);
  if (e == cudaErrorNoDevice || e == cudaErrorInvalidConfiguration || e == cudaErrorNotSupported)
    throw Exception(Error::UnsupportedHardware, msg);
  throw Exception(Error::Unknown, msg);
}

void checkABI(int rc, const char* what)
{
  if (rc == 0) return;
  const std::string msg = std::string(what) + ": " + oidnb200_last_error();
  if (rc == OIDNB200_ERR_INVALID) throw Exception(Error::InvalidArgument, msg);
  if (rc == OIDNB200_ERR_UNSUPPORTED) throw Exception(Error::UnsupportedHardware, msg);
  if (rc == (int)cudaErrorMemoryAllocation) throw Exception(Error::OutOfMemory, msg);
  throw Exception(Error::Unknown, msg);
}

static oidnb200_image abiImage(const Image& im)
{
  return oidnb200_image{im.ptr, (int)im.format, im.W, im.H, im.pixelStride, im.rowStride};
}

// ------------------------------------------------------------------------------------------------
Engine::Engine(int deviceID, void* userStream) : deviceID(deviceID)
{
  checkCuda(cudaSetDevice(deviceID), "cudaSetDevice");
  if (userStream)
    stream = userStream;
  else
  {
    cudaStream_t s;
    checkCuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
    stream = s;
    ownStream = true;
  }
}

Engine::~Engine()
{
  cudaSetDevice(deviceID);
  for (void* e : events) cudaEventDestroy(static_cast<cudaEvent_t>(e));
  for (void* s : aux)
    if (s) cudaStreamDestroy(static_cast<cudaStream_t>(s));
  if (ownStream && stream) cudaStreamDestroy(static_cast<cudaStream_t>(stream));
}

void* Engine::getAuxStream(AuxStream which)
{
  if (!aux[which])
  {
    makeCurrent();
    cudaStream_t s;
    checkCuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
    aux[which] = s;
  }
  return aux[which];
}

void* Engine::newEvent()
{
  makeCurrent();
  cudaEvent_t e;
  checkCuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
  events.push_back(e);
  return e;
}

void Engine::makeCurrent() const { checkCuda(cudaSetDevice(deviceID), "cudaSetDevice"); }

void* Engine::malloc(size_t bytes, Storage storage)
{
  makeCurrent();
  void* p = nullptr;
  if (bytes == 0) return nullptr;
  switch (storage)
  {
  case Storage::Host:    checkCuda(cudaMallocHost(&p, bytes), "cudaMallocHost"); break;
  case Storage::Managed: checkCuda(cudaMallocManaged(&p, bytes), "cudaMallocManaged"); break;
  case Storage::Device:  checkCuda(cudaMalloc(&p, bytes), "cudaMalloc"); break;
  default: throw Exception(Error::InvalidArgument, "invalid storage mode");
  }
  return p;
}

void Engine::free(void* ptr, Storage storage)
{
  if (!ptr) return;
  cudaSetDevice(deviceID);
  if (storage == Storage::Host) cudaFreeHost(ptr); else cudaFree(ptr);
}

void Engine::submitCopy(void* dst, const void* src, size_t bytes)
{
  makeCurrent();
  checkCuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(getStream())), "cudaMemcpyAsync");
}

void Engine::submitCopy2D(void* dst, size_t dstPitch, const void* src, size_t srcPitch, size_t widthBytes, size_t height)
{
  makeCurrent();
  checkCuda(cudaMemcpy2DAsync(dst, dstPitch, src, srcPitch, widthBytes, height, cudaMemcpyDefault,
                              static_cast<cudaStream_t>(getStream())), "cudaMemcpy2DAsync");
}

static void CUDART_CB hostFuncTrampoline(void* p)
{
  std::unique_ptr<std::function<void()>> f(static_cast<std::function<void()>*>(p));
  (*f)();
}

void Engine::submitHostFunc(std::function<void()>&& f)
{
  makeCurrent();
  auto* heap = new std::function<void()>(std::move(f));
  checkCuda(cudaLaunchHostFunc(static_cast<cudaStream_t>(getStream()), hostFuncTrampoline, heap), "cudaLaunchHostFunc");
}

void Engine::wait()
{
  makeCurrent();
  checkCuda(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)), "cudaStreamSynchronize");
  for (void* s : aux)
    if (s) checkCuda(cudaStreamSynchronize(static_cast<cudaStream_t>(s)), "cudaStreamSynchronize");
}

// ------------------------------------------------------------------------------------------------
Conv::Conv(Engine* engine, const ConvDesc& d) : engine(engine), desc(d)
{
  oidnb200_conv_desc a{};
  a.H = d.H; a.W = d.W;
  a.C1 = d.src1.paddedC(); a.C2 = d.src2.C > 0 ? d.src2.paddedC() : 0;
  a.Cout = round_up(d.outC, 16);
  a.relu = d.activation == Activation::ReLU;
  a.post_op = d.postOp == PostOp::Pool ? 1 : 0; // Upsample is folded into the consumer (virtual tensor)
  a.src1_upsampled = d.src1Upsampled;
  a.shift_mode = 0;
  checkABI(oidnb200_conv_create(&a, &handle), ("conv '" + name + "'").c_str());
}

Conv::~Conv() { if (handle) oidnb200_conv_destroy(handle); }

TensorDesc Conv::getDstDesc() const
{
  if (desc.postOp == PostOp::Pool) return TensorDesc{desc.outC, desc.H / 2, desc.W / 2};
  return TensorDesc{desc.outC, desc.H, desc.W};
}

size_t Conv::getWeightByteSize() const { return oidnb200_conv_weight_bytes(handle); }
size_t Conv::getBiasByteSize() const { return oidnb200_conv_bias_bytes(handle); }

void Conv::packWeight(const uint16_t* oihw, int O, int I1, int I2, void* dstHost) const
{
  checkABI(oidnb200_conv_pack_weights(handle, oihw, O, I1, I2, dstHost), "conv weight reorder");
}

void Conv::packBias(const uint16_t* x, int O, void* dstHost) const
{
  checkABI(oidnb200_conv_pack_bias(handle, x, O, dstHost), "conv bias reorder");
}

void Conv::finalize()
{
  engine->makeCurrent();
  checkABI(oidnb200_conv_bind(handle, src1, src2, weight, bias, dst), ("conv '" + name + "' bind").c_str());
  bound = true;
}

void Conv::submitKernels()
{
  if (!bound) finalize();
  checkABI(oidnb200_conv_launch(handle, engine->getStream()), ("conv '" + name + "'").c_str());
}

oidnb200_conv_info Conv::getInfo() const
{
  oidnb200_conv_info i{};
  oidnb200_conv_get_info(handle, &i);
  return i;
}

std::unique_ptr<ConvPair> ConvPair::tryCreate(Conv& a, Conv& b)
{
  oidnb200_conv_pair* h = nullptr;
  const int rc = oidnb200_conv_pair_create(a.getHandle(), b.getHandle(), &h);
  if (rc == OIDNB200_ERR_UNSUPPORTED) return nullptr;
  checkABI(rc, "conv pair");
  std::unique_ptr<ConvPair> p(new ConvPair());
  p->a = &a; p->b = &b; p->handle = h;
  return p;
}

ConvPair::~ConvPair() { if (handle) oidnb200_conv_pair_destroy(handle); }

void ConvPair::submit()
{
  if (!a->isBound()) a->finalize();
  if (!b->isBound()) b->finalize();
  checkABI(oidnb200_conv_pair_bind(handle), "conv pair bind"); // cheap: copies the tensor maps of the two bound convs
  checkABI(oidnb200_conv_pair_launch(handle, a->getEngine()->getStream()),
           ("conv pair '" + a->getName() + "' + '" + b->getName() + "'").c_str());
}

oidnb200_conv_info ConvPair::getInfo() const
{
  oidnb200_conv_info i{};
  oidnb200_conv_pair_get_info(handle, &i);
  return i;
}

void Pool::submitKernels()
{
  checkABI(oidnb200_pool_launch(src, srcDesc.H, srcDesc.W, srcDesc.paddedC(), dst, engine->getStream()), "pool");
}

void Upsample::submitKernels()
{
  checkABI(oidnb200_upsample_launch(src, srcDesc.H, srcDesc.W, srcDesc.paddedC(), dst, engine->getStream()), "upsample");
}

void InputProcess::setSrc(const Image& color, const Image& alb, const Image& nrm)
{
  // core/input_process.cpp:29-56: the first set image is the main input
  if (color) { input = color; albedo = alb; normal = nrm; }
  else if (alb) { input = alb; albedo = Image(); normal = Image(); }
  else { input = nrm; albedo = Image(); normal = Image(); }
}

void InputProcess::submitKernels()
{
  const oidnb200_image c = abiImage(input), a = abiImage(albedo), n = abiImage(normal);
  const oidnb200_transfer tf = transferFunc->abi();
  checkABI(oidnb200_input_process_launch(&c, &a, &n, &tile, &tf, hdr, snorm, dst, dstDesc.H, dstDesc.W,
                                         dstDesc.paddedC(), engine->getStream()), "input process");
}

void OutputProcess::submitKernels()
{
  const oidnb200_image d = abiImage(dst);
  const oidnb200_transfer tf = transferFunc->abi();
  checkABI(oidnb200_output_process_launch(src, srcDesc.H, srcDesc.W, srcDesc.paddedC(), &tile, &tf, hdr, snorm, &d,
                                          engine->getStream()), "output process");
}

bool OutputProcess::fuseInto(Conv& producer, bool on)
{
  if (on)
  {
    const oidnb200_image d = abiImage(dst);
    const oidnb200_transfer tf = transferFunc->abi();
    const int rc = oidnb200_conv_set_output_process(producer.getHandle(), &tile, &tf, hdr, snorm, &d);
    if (rc == 0) return true;
    if (rc != OIDNB200_ERR_UNSUPPORTED) checkABI(rc, "fused output process");
  }
  oidnb200_conv_set_output_process(producer.getHandle(), nullptr, nullptr, 0, 0, nullptr);
  return false;
}

void Autoexposure::setSrc(const Image& image)
{
  if (!image || image.W != W || image.H != H) throw std::invalid_argument("invalid autoexposure source");
  src = image;
}

void Autoexposure::submitKernels()
{
  const oidnb200_image s = abiImage(src);
  checkABI(oidnb200_autoexposure_launch(&s, scratch, dst, engine->getStream()), "autoexposure");
}

void ImageCopy::submitKernels()
{
  const oidnb200_image s = abiImage(src), d = abiImage(dst);
  checkABI(oidnb200_image_copy_launch(&s, &d, engine->getStream()), "image copy");
}

} // namespace oidnb200
