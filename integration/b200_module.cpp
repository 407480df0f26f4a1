// Device module that plugs oidn_b200 into the reference tree (Open Image Denoise 2.4.1) as its CUDA device:
// built as lib${OIDN_LIBRARY_NAME}_device_cuda.so next to libOpenImageDenoise_core.so, it is what the reference's
// module loader picks up for DeviceType::CUDA (core/module.cpp:32-75, core/context.h:33-50), so the reference's own
// applications (oidnBenchmark, oidnDenoise, oidnTest) run on this backend through the unchanged public API.
//
// This is INTEGRATION.md's "filter-level" route: the device keeps graph construction, arena planning and tile
// scheduling inside liboidn_b200.so (oidn_b200/csrc/host) and hands the core a Filter that forwards to the
// filter-level C ABI (include/oidn_b200.h). It needs ONE word changed in the reference core: `virtual` on
// Device::newFilter (core/device.h:87); tools/build_integration_module.sh applies it to its /tmp copy. Buffers
// (oidnNewBuffer, oidnRead/WriteBuffer) stay the reference's own USMBuffer on top of this engine's usmAlloc/usmCopy
// (core/engine.h:78-91). The op factories of the Engine (newConv, ...) are never reached on this route: they throw.
//
// Built with -DOIDN_B200_OP_LEVEL the same file gives INTEGRATION.md's op-level route instead: NO core change at
// all; the reference's own filters, core/graph.cpp, arena planner and tile loop stay in charge and this module
// only supplies the Engine's ops (Conv incl. fused Pool and the paired ConcatConv, Pool, Upsample, InputProcess,
// OutputProcess, Autoexposure, ImageCopy) on the kernel-level C ABI (include/oidn_b200_kernels.h). The fusions of
// the filter-level route are kept under the core's graph by linking ops through the tensors they share:
//   * the two kernels that ConcatConvHWC asks for (core/concat_conv_hwc.cpp:27-31,62-80) are recognised and run as
//     ONE concat-conv with two K segments;
//   * PostOp::Upsample is accepted (isConvSupported), the producer stores its result at its own resolution at the
//     start of the (4x larger) tensor the core allocated and marks the tensor; a consumer that finds its source
//     marked reads it through the stride-0 "dup" TMA axis (virtual nearest upsample: no upsampled tensor in HBM);
//   * a plain conv whose tensor has exactly one reader, another plain conv, runs fused with it as ONE launch
//     (oidnb200_conv_pair_*: enc_conv0 -> enc_conv1, dec_conv1b -> dec_conv0): the first conv skips its launch, the
//     second launches the pair -- unless the core's arena planner has put the pair's destination on top of its
//     source (it may: under two launches the source is dead by then), in which case the two convs run one by one;
//   * the OutputProcess op finds the conv that produces its source tensor and, when the frame's output image is
//     packed fp32 RGB, hands it its tile / image / transfer function at submit time: the conv's epilogue writes
//     the image (oidnb200_conv_set_output_process) and the op's own kernel is skipped (core/unet_filter.cpp:228-236
//     sets the tile before Graph::submit runs the ops in order, so the conv sees the current tile).
//
// Only interfaces are taken from the reference headers; no reference code is copied.
#include "core/context.h"
#include "core/device.h"
#include "core/engine.h"
#include "core/filter.h"
#include "core/conv.h"
#include "core/pool.h"
#include "core/upsample.h"
#include "core/autoexposure.h"
#include "core/input_process.h"
#include "core/output_process.h"
#include "core/image_copy.h"
#include "oidn_b200.h"
#include "oidn_b200_kernels.h"
#include <cuda_runtime.h>
#include <map>
#include <unordered_map>
#include <vector>

OIDN_NAMESPACE_BEGIN

  namespace
  {
    void checkCuda(cudaError_t e)
    {
      if (e == cudaSuccess)
        return;
      const char* str = cudaGetErrorString(e);
      cudaGetLastError();
      // the reference CUDA device's mapping (devices/cuda/cuda_device.cpp:34-51)
      if (e == cudaErrorMemoryAllocation)
        throw Exception(Error::OutOfMemory, str);
      if (e == cudaErrorNoDevice || e == cudaErrorInvalidConfiguration || e == cudaErrorNotSupported)
        throw Exception(Error::UnsupportedHardware, str);
      throw Exception(Error::Unknown, str);
    }

  #if !defined(OIDN_B200_OP_LEVEL)
    // oidn_b200 keeps the reference's error codes (include/oidn_b200.h) and its first-error-per-device slot
    void checkB200(OIDNB200Device h)
    {
      const char* message = nullptr;
      const int code = oidnb200GetDeviceError(h, &message);
      if (code != 0)
        throw Exception(static_cast<Error>(code), message ? message : "oidn_b200 error");
    }
  #endif
  }

  class B200PhysicalDevice final : public PhysicalDevice
  {
  public:
    int deviceID;

    B200PhysicalDevice(int deviceID, const cudaDeviceProp& prop, int score)
      : PhysicalDevice(DeviceType::CUDA, score),
        deviceID(deviceID)
    {
      name = prop.name;
      memcpy(uuid.bytes, prop.uuid.bytes, sizeof(uuid.bytes));
      uuidSupported = true;
      pciDomain = prop.pciDomainID;
      pciBus = prop.pciBusID;
      pciDevice = prop.pciDeviceID;
      pciFunction = 0;
      pciAddressSupported = true;
    }
  };

  class B200Device;

  // One GPU + one stream: memory and ordering services for the core's buffers
  class B200Engine final : public Engine
  {
  public:
    B200Engine(Device* device, cudaStream_t stream) : device(device), stream(stream) {}

    Device* getDevice() const override { return device; }
    cudaStream_t getStream() const { return stream; }

  #if defined(OIDN_B200_OP_LEVEL)
    // Pool is fused into the conv's epilogue, Upsample into the consumer's loader (see the file comment)
    bool isConvSupported(PostOp postOp) override { return postOp == PostOp::Pool || postOp == PostOp::Upsample; }
    Ref<Conv> newConv(const ConvDesc& desc) override;
    Ref<Pool> newPool(const PoolDesc& desc) override;
    Ref<Upsample> newUpsample(const UpsampleDesc& desc) override;
    Ref<Autoexposure> newAutoexposure(const ImageDesc& srcDesc) override;
    Ref<InputProcess> newInputProcess(const InputProcessDesc& desc) override;
    Ref<OutputProcess> newOutputProcess(const OutputProcessDesc& desc) override;
    Ref<ImageCopy> newImageCopy() override;

    // ConcatConvHWC = conv1{src1, W1, bias, no activation} then conv2{src2, W2, bias = dst tensor, activation}
    // (core/concat_conv_hwc.cpp:27-31,62-67): conv1 registers under its dst tensor, conv2 finds it by its bias
    // tensor and the pair runs as one kernel with two K segments.
    std::unordered_map<const Tensor*, class B200Conv*> convByDst;
    std::unordered_map<const Tensor*, class B200Conv*> producerOf;   // every conv, by its destination tensor
    std::unordered_map<const Tensor*, int> halfRes;                  // tensors stored at half resolution (virtual upsample)
    std::unordered_map<const Tensor*, int> readers;                  // how many ops read a tensor (conv pairs need exactly one)
    std::unordered_map<const Tensor*, class B200Conv*> convReader;   // the last conv that registered as reader of a tensor
  #else
    Ref<Conv> newConv(const ConvDesc&) override { unsupported(); return nullptr; }
    Ref<Pool> newPool(const PoolDesc&) override { unsupported(); return nullptr; }
    Ref<Upsample> newUpsample(const UpsampleDesc&) override { unsupported(); return nullptr; }
    Ref<Autoexposure> newAutoexposure(const ImageDesc&) override { unsupported(); return nullptr; }
    Ref<InputProcess> newInputProcess(const InputProcessDesc&) override { unsupported(); return nullptr; }
    Ref<OutputProcess> newOutputProcess(const OutputProcessDesc&) override { unsupported(); return nullptr; }
    Ref<ImageCopy> newImageCopy() override { unsupported(); return nullptr; }
  #endif

    void* usmAlloc(size_t byteSize, Storage storage) override
    {
      if (byteSize == 0)
        return nullptr;
      void* ptr = nullptr;
      switch (storage)
      {
      case Storage::Host:    checkCuda(cudaMallocHost(&ptr, byteSize)); break;
      case Storage::Device:  checkCuda(cudaMalloc(&ptr, byteSize)); break;
      case Storage::Managed: checkCuda(cudaMallocManaged(&ptr, byteSize)); break;
      default:               throw Exception(Error::InvalidArgument, "invalid storage mode");
      }
      return ptr;
    }

    void usmFree(void* ptr, Storage storage) override
    {
      if (ptr == nullptr)
        return;
      checkCuda(storage == Storage::Host ? cudaFreeHost(ptr) : cudaFree(ptr));
    }

    void usmCopy(void* dstPtr, const void* srcPtr, size_t byteSize) override
    {
      checkCuda(cudaMemcpy(dstPtr, srcPtr, byteSize, cudaMemcpyDefault));
    }

    void submitUSMCopy(void* dstPtr, const void* srcPtr, size_t byteSize) override
    {
      checkCuda(cudaMemcpyAsync(dstPtr, srcPtr, byteSize, cudaMemcpyDefault, stream));
    }

    void submitHostFunc(std::function<void()>&& f, const Ref<CancellationToken>&) override
    {
      auto* heap = new std::function<void()>(std::move(f));
      checkCuda(cudaLaunchHostFunc(stream, [](void* p) {
        std::unique_ptr<std::function<void()>> g(static_cast<std::function<void()>*>(p));
        (*g)();
      }, heap));
    }

    void wait() override { checkCuda(cudaStreamSynchronize(stream)); }

  private:
  #if !defined(OIDN_B200_OP_LEVEL)
    static void unsupported()
    {
      throw Exception(Error::InvalidOperation, "the oidn_b200 device builds its own graph: core ops are not created through the engine");
    }
  #endif

    Device* device;
    cudaStream_t stream;
  };

  class B200Device final : public Device
  {
  public:
    static bool isSupported(const cudaDeviceProp& prop) { return prop.major == 10 && prop.unifiedAddressing; }

    static bool isSupported(int deviceID)
    {
      cudaDeviceProp prop{};
      return cudaGetDeviceProperties(&prop, deviceID) == cudaSuccess && isSupported(prop);
    }

    static std::vector<Ref<PhysicalDevice>> getPhysicalDevices()
    {
      int numDevices = 0;
      if (cudaGetDeviceCount(&numDevices) != cudaSuccess)
      {
        cudaGetLastError();
        return {};
      }
      std::vector<Ref<PhysicalDevice>> devices;
      for (int id = 0; id < numDevices; ++id)
      {
        cudaDeviceProp prop{};
        if (cudaGetDeviceProperties(&prop, id) == cudaSuccess && isSupported(prop))
          devices.push_back(makeRef<B200PhysicalDevice>(id, prop, (19 << 16) - 1 - id)); // the CUDA device's score
      }
      return devices;
    }

    B200Device(int deviceID, cudaStream_t stream) : deviceID(deviceID), stream(stream) {}
    explicit B200Device(const Ref<B200PhysicalDevice>& pd) : deviceID(pd->deviceID) {}

    ~B200Device()
    {
      try
      {
        enter();
        if (handle)
          oidnb200ReleaseDevice(handle);
        subdevices.clear();
        if (ownStream)
          cudaStreamDestroy(stream);
        leave();
      }
      catch (...) {}
    }

    void enter() override
    {
      if (prevDeviceID >= 0)
      {
        checkCuda(cudaGetDevice(&prevDeviceID));
        if (prevDeviceID != deviceID)
          checkCuda(cudaSetDevice(deviceID));
      }
      cudaGetLastError();
    }

    void leave() override
    {
      if (prevDeviceID >= 0 && prevDeviceID != deviceID)
        checkCuda(cudaSetDevice(prevDeviceID));
    }

    DeviceType getType() const override { return DeviceType::CUDA; }

    Storage getPtrStorage(const void* ptr) override
    {
      cudaPointerAttributes attrib;
      if (cudaPointerGetAttributes(&attrib, ptr) != cudaSuccess)
      {
        cudaGetLastError();
        return Storage::Undefined;
      }
      switch (attrib.type)
      {
      case cudaMemoryTypeHost:    return Storage::Host;
      case cudaMemoryTypeDevice:  return Storage::Device;
      case cudaMemoryTypeManaged: return Storage::Managed;
      default:                    return systemMemorySupported ? Storage::Managed : Storage::Undefined;
      }
    }

    void wait() override
    {
      for (auto& subdevice : subdevices)
        subdevice->getEngine()->wait();
    }

  #if defined(OIDN_B200_OP_LEVEL)
    // final weights / biases stay in host memory (core/graph.cpp:113-140): the conv repacks them into its own
    // layout and uploads that
    bool needWeightAndBiasOnDevice() const override { return false; }
  #else
    // Needs `virtual` on Device::newFilter (the one-word core change)
    Ref<Filter> newFilter(const std::string& type) override;
  #endif

    OIDNB200Device getHandle() const { return handle; }

  private:
    void init() override
    {
      cudaDeviceProp prop{};
      checkCuda(cudaGetDeviceProperties(&prop, deviceID));
      if (!isSupported(prop))
        throw Exception(Error::UnsupportedHardware, "the oidn_b200 device needs a compute capability 10.x GPU");
      if (isVerbose())
      {
        std::cout << "  Device    : " << prop.name << std::endl;
        std::cout << "    Type    : CUDA (oidn_b200: tcgen05/TMEM/TMA kernels for sm_100a)" << std::endl;
        std::cout << "    SMs     : " << prop.multiProcessorCount << std::endl;
      }

      checkCuda(cudaGetDevice(&prevDeviceID));
      if (prevDeviceID != deviceID)
        checkCuda(cudaSetDevice(deviceID));

      // what the core asks a device for (devices/cuda/cuda_device.cpp:215-219 is the precedent); the filters of
      // this device do not go through core/graph.cpp, so only the memory capabilities matter
      tensorDataType = DataType::Float16;
      weightDataType = DataType::Float16;
      tensorLayout   = TensorLayout::hwc;
      weightLayout   = TensorLayout::ohwi;
      tensorBlockC   = 16;
      systemMemorySupported  = prop.pageableMemoryAccess;
      managedMemorySupported = prop.managedMemory;

      if (!stream)
      {
        checkCuda(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        ownStream = true;
      }
      subdevices.emplace_back(new Subdevice(std::unique_ptr<Engine>(new B200Engine(this, stream))));

    #if !defined(OIDN_B200_OP_LEVEL)
      // the backend runs on the same stream as the core's buffer copies: everything stays stream-ordered
      void* streams[1] = {stream};
      handle = oidnb200NewCUDADevice(&deviceID, streams, 1);
      if (!handle)
        throw Exception(Error::Unknown, "oidnb200NewCUDADevice failed");
      oidnb200SetDeviceInt(handle, "verbose", verbose);
      oidnb200CommitDevice(handle);
      checkB200(handle);
    #endif
    }

    int deviceID = 0;
    int prevDeviceID = -1;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    OIDNB200Device handle = nullptr;
  };

#if !defined(OIDN_B200_OP_LEVEL)
  // core/filter.h on top of the filter-level C ABI. Parameter names, defaults, dirty tracking and error behaviour are
  // those of the reference filters because oidn_b200/csrc/host/filter.cpp mirrors core/unet_filter.cpp, rt_filter.cpp
  // and rtlightmap_filter.cpp; this class only forwards.
  class B200Filter final : public Filter
  {
  public:
    B200Filter(const Ref<B200Device>& device, const std::string& type)
      : Filter(device),
        dev(device->getHandle())
    {
      handle = oidnb200NewFilter(dev, type.c_str());
      checkB200(dev);
      if (!handle)
        throw Exception(Error::InvalidArgument, "unknown filter type: '" + type + "'");
    }

    ~B200Filter()
    {
      if (handle)
        oidnb200ReleaseFilter(handle);
    }

    void setImage(const std::string& name, const Ref<Image>& image) override
    {
      if (image && *image)
      {
        const ImageDesc& d = image->getDesc();
        oidnb200SetSharedFilterImage(handle, name.c_str(), image->getPtr(), static_cast<int>(d.format),
                                     d.width, d.height, 0, d.wByteStride, d.hByteStride);
        checkB200(dev);
        images[name] = image; // the filter keeps its images (and their buffers) alive, as the reference filters do
      }
      else
        unsetImage(name);
    }

    void unsetImage(const std::string& name) override
    {
      oidnb200UnsetFilterImage(handle, name.c_str());
      checkB200(dev);
      images.erase(name);
    }

    void setData(const std::string& name, const Data& data) override
    {
      oidnb200SetSharedFilterData(handle, name.c_str(), const_cast<void*>(data.ptr), data.size);
      checkB200(dev);
    }

    void updateData(const std::string& name) override
    {
      oidnb200UpdateFilterData(handle, name.c_str());
      checkB200(dev);
    }

    void unsetData(const std::string& name) override
    {
      oidnb200UnsetFilterData(handle, name.c_str());
      checkB200(dev);
    }

    void setInt(const std::string& name, int value) override
    {
      if (isBool(name)) // the API layer passes bools through setInt (api/api.cpp:850-859)
        oidnb200SetFilterBool(handle, name.c_str(), value != 0);
      else
        oidnb200SetFilterInt(handle, name.c_str(), value);
      checkB200(dev);
    }

    int getInt(const std::string& name) override
    {
      const int value = isBool(name) ? int(oidnb200GetFilterBool(handle, name.c_str()))
                                     : oidnb200GetFilterInt(handle, name.c_str());
      checkB200(dev);
      return value;
    }

    void setFloat(const std::string& name, float value) override
    {
      oidnb200SetFilterFloat(handle, name.c_str(), value);
      checkB200(dev);
    }

    float getFloat(const std::string& name) override
    {
      const float value = oidnb200GetFilterFloat(handle, name.c_str());
      checkB200(dev);
      return value;
    }

    void commit() override
    {
      oidnb200CommitFilter(handle);
      checkB200(dev);
    }

    void execute(SyncMode sync) override
    {
      oidnb200SetFilterProgressMonitorFunction(handle, progressFunc, progressUserPtr);
      if (sync == SyncMode::Blocking)
        oidnb200ExecuteFilter(handle);
      else
        oidnb200ExecuteFilterAsync(handle);
      checkB200(dev);
    }

  private:
    static bool isBool(const std::string& name)
    {
      return name == "hdr" || name == "srgb" || name == "cleanAux" || name == "directional";
    }

    OIDNB200Device dev;
    OIDNB200Filter handle = nullptr;
    std::map<std::string, Ref<Image>> images;
  };

  Ref<Filter> B200Device::newFilter(const std::string& type)
  {
    if (isVerbose(2))
      std::cout << "Filter: " << type << " (oidn_b200)" << std::endl;
    return makeRef<B200Filter>(Ref<B200Device>(this), type);
  }
#else
  // ---------------------------------------------------------------------------------------------------------------
  // Op-level route: the Engine's ops on the kernel-level C ABI
  // ---------------------------------------------------------------------------------------------------------------
  namespace
  {
    void checkKernel(int rc, const char* what)
    {
      if (rc == 0)
        return;
      const std::string msg = std::string(what) + ": " + oidnb200_last_error();
      if (rc == OIDNB200_ERR_INVALID)     throw Exception(Error::InvalidArgument, msg);
      if (rc == OIDNB200_ERR_UNSUPPORTED) throw Exception(Error::UnsupportedHardware, msg);
      if (rc == (int)cudaErrorMemoryAllocation) throw Exception(Error::OutOfMemory, msg);
      throw Exception(Error::Unknown, msg);
    }

    oidnb200_image toABI(const Ref<Image>& image)
    {
      if (!image || !*image)
        return oidnb200_image{nullptr, 0, 0, 0, 0, 0};
      const ImageDesc& d = image->getDesc();
      return oidnb200_image{image->getPtr(), static_cast<int>(d.format), d.getW(), d.getH(), d.wByteStride, d.hByteStride};
    }

    oidnb200_transfer toABI(const TransferFunction& tf)
    {
      return oidnb200_transfer{static_cast<int>(tf.type), tf.inputScale, tf.inputScalePtr};
    }

    oidnb200_tile toABI(const Tile& t)
    {
      return oidnb200_tile{t.hSrcBegin, t.wSrcBegin, t.hDstBegin, t.wDstBegin, t.H, t.W};
    }
  }

  class B200OutputProcess;

  class B200Conv final : public Conv
  {
  public:
    B200Conv(B200Engine* engine, const ConvDesc& desc)
      : Conv(desc),
        engine(engine)
    {
      if (srcDesc.layout != TensorLayout::hwc || srcDesc.dataType != DataType::Float16 ||
          weightDesc.layout != TensorLayout::ohwi || weightDesc.dataType != DataType::Float16 ||
          weightDesc.getH() != 3 || weightDesc.getW() != 3)
        throw std::invalid_argument("unsupported convolution");
      accumulate = biasDesc.getRank() == 3; // second half of a ConcatConvHWC: bias = dst, dst += conv
    }

    ~B200Conv()
    {
      unregister();
      release();
      detachOutputProcess();
    }
    friend class B200OutputProcess;

    // set by the OutputProcess that consumes this conv's tensor (B200OutputProcess::finalize)
    B200OutputProcess* outputProcess = nullptr;
    void detachOutputProcess();
    bool canFuseOutput() const
    {
      return postOp == PostOp::None && !accumulate && weightDesc.getPaddedO() == 16 && activation == Activation::ReLU;
    }

    Engine* getEngine() const override { return engine; }

    void finalize() override
    {
      if (!src || !weight || !bias || !dst)
        throw std::logic_error("convolution source/weight/bias/destination not set");
      absorbed = false;
      partner = nullptr;
      unregister();
      registeredDst = dst.get();
      engine->producerOf[registeredDst] = this;
      if (postOp == PostOp::Upsample)
        engine->halfRes[registeredDst] = 1;
      srcHalfRes = engine->halfRes.count(src.get()) != 0;
      registeredSrc = src.get();
      engine->readers[registeredSrc] += 1;
      engine->convReader[registeredSrc] = this;
      pairChecked = false;
      if (accumulate)
      {
        auto it = engine->convByDst.find(bias.get());
        if (it == engine->convByDst.end() || it->second->activation != Activation::None || it->second->accumulate)
          throw std::logic_error("accumulating convolution without the first half of its ConcatConv");
        partner = it->second;
        partner->absorbed = true;   // conv1 becomes a no-op: this op runs both K segments
      }
      else if (activation == Activation::None)
        engine->convByDst[dst.get()] = this; // may be the first half of a ConcatConv
      prepared = false;
    }

    void submitKernels(const Ref<CancellationToken>&) override;

  private:
    void updateWeight() override { prepared = false; }
    void updateBias() override { prepared = false; }

    void unregister()
    {
      if (!registeredDst)
        return;
      auto erase = [&](std::unordered_map<const Tensor*, B200Conv*>& m) {
        auto it = m.find(registeredDst);
        if (it != m.end() && it->second == this)
          m.erase(it);
      };
      erase(engine->convByDst);
      erase(engine->producerOf);
      if (postOp == PostOp::Upsample)
        engine->halfRes.erase(registeredDst);
      registeredDst = nullptr;
      if (registeredSrc)
      {
        auto it = engine->readers.find(registeredSrc);
        if (it != engine->readers.end() && --it->second <= 0)
          engine->readers.erase(it);
        auto cr = engine->convReader.find(registeredSrc);
        if (cr != engine->convReader.end() && cr->second == this)
          engine->convReader.erase(cr);
        registeredSrc = nullptr;
      }
      dropPair();
    }

    // conv pair (this = first conv A): `pair` runs A and `pairB` in one launch
    void dropPair()
    {
      if (pair) { oidnb200_conv_pair_destroy(pair); pair = nullptr; }
      if (pairB) { pairB->pairA = nullptr; pairB = nullptr; }
      if (pairA) { pairA->dropPair(); }
    }

    bool isPlain() const { return !accumulate && !absorbed && !partner; }

    void tryPair()
    {
      pairChecked = true;
      if (!isPlain() || postOp != PostOp::None || !dst)
        return;
      auto rd = engine->readers.find(dst.get());
      auto cr = engine->convReader.find(dst.get());
      if (rd == engine->readers.end() || rd->second != 1 || cr == engine->convReader.end())
        return;
      B200Conv* b = cr->second;
      if (b == this || !b->isPlain() || b->src.get() != dst.get() || b->pairA)
        return;
      prepare(); bind();
      b->prepare(); b->bind();
      oidnb200_conv_pair* h = nullptr;
      if (oidnb200_conv_pair_create(handle, b->handle, &h) != 0)
        return; // shapes not covered: two launches
      pair = h; pairB = b; b->pairA = this;
    }

    void release()
    {
      dropPair();       // the pair refers to the two kernel-level ops
      pairChecked = false;
      if (handle) { oidnb200_conv_destroy(handle); handle = nullptr; }
      if (packedWeight) { cudaFree(packedWeight); packedWeight = nullptr; }
      if (packedBias) { cudaFree(packedBias); packedBias = nullptr; }
    }

    // ohwi fp16 host tensor [paddedO][3][3][paddedI] -> logical oihw rows appended to `out` ([O][Itotal][3][3])
    static void gatherOIHW(const Tensor& w, int O, int Itotal, int iOffset, std::vector<uint16_t>& out)
    {
      const uint16_t* p = static_cast<const uint16_t*>(w.getPtr());
      const int I = w.getI(), Ipad = w.getPaddedI();
      for (int o = 0; o < O; ++o)
        for (int i = 0; i < I; ++i)
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw)
              out[(((size_t)o * Itotal + iOffset + i) * 3 + kh) * 3 + kw] = p[(((size_t)o * 3 + kh) * 3 + kw) * Ipad + i];
    }

    // creates the kernel-level op and uploads weights / bias in its packed layout (once per weight update)
    void prepare()
    {
      if (prepared)
        return;
      release();
      const B200Conv* first = partner ? partner : this;     // conv1 of a pair holds src1, W1 and the rank-1 bias
      oidnb200_conv_desc d{};
      d.H = srcDesc.getH(); d.W = srcDesc.getW();
      d.C1 = first->srcDesc.getPaddedC();
      d.C2 = partner ? srcDesc.getPaddedC() : 0;
      d.Cout = weightDesc.getPaddedO();
      d.relu = activation == Activation::ReLU;
      d.post_op = postOp == PostOp::Pool ? 1 : 0;  // Upsample: stored at the conv's own resolution, consumers dup it
      d.src1_upsampled = first->srcHalfRes ? 1 : 0;
      if (partner && srcHalfRes)
        throw std::logic_error("the second source of a concat-conv cannot be an upsampled tensor");
      d.shift_mode = 0;
      checkKernel(oidnb200_conv_create(&d, &handle), "conv create");

      const int O = weightDesc.getO();
      const int I1 = first->weight->getI(), I2 = partner ? weight->getI() : 0;
      std::vector<uint16_t> oihw((size_t)O * (I1 + I2) * 9);
      gatherOIHW(*first->weight, O, I1 + I2, 0, oihw);
      if (partner)
        gatherOIHW(*weight, O, I1 + I2, I1, oihw);
      std::vector<uint8_t> hostW(oidnb200_conv_weight_bytes(handle)), hostB(oidnb200_conv_bias_bytes(handle));
      checkKernel(oidnb200_conv_pack_weights(handle, oihw.data(), O, I1, I2, hostW.data()), "conv weight reorder");
      checkKernel(oidnb200_conv_pack_bias(handle, static_cast<const uint16_t*>(first->bias->getPtr()), O, hostB.data()),
                  "conv bias reorder");
      checkCuda(cudaMalloc(&packedWeight, hostW.size()));
      checkCuda(cudaMalloc(&packedBias, hostB.size()));
      checkCuda(cudaMemcpy(packedWeight, hostW.data(), hostW.size(), cudaMemcpyHostToDevice));
      checkCuda(cudaMemcpy(packedBias, hostB.data(), hostB.size(), cudaMemcpyHostToDevice));
      boundSrc1 = boundSrc2 = boundDst = nullptr;
      prepared = true;
    }

    // tensor pointers move when the shared scratch heap is reallocated (Tensor::postRealloc): re-encode the TMA
    // tensor maps whenever they differ from what is bound
    void bind()
    {
      const void* s1 = partner ? partner->src->getPtr() : src->getPtr();
      const void* s2 = partner ? src->getPtr() : nullptr;
      void* d = dst->getPtr();
      if (s1 == boundSrc1 && s2 == boundSrc2 && d == boundDst)
        return;
      checkKernel(oidnb200_conv_bind(handle, s1, s2, packedWeight, packedBias, d), "conv bind");
      boundSrc1 = s1; boundSrc2 = s2; boundDst = d;
    }

    B200Engine* engine;
    const Tensor* registeredDst = nullptr;
    const Tensor* registeredSrc = nullptr;
    bool pairChecked = false;
    oidnb200_conv_pair* pair = nullptr;   // this conv is the first of a fused pair
    B200Conv* pairB = nullptr;            // ... whose second conv launches it
    B200Conv* pairA = nullptr;            // this conv is the second of a fused pair
    bool srcHalfRes = false;   // the source tensor holds a half-resolution image (its producer had PostOp::Upsample)
    bool accumulate = false;
    bool absorbed = false;
    B200Conv* partner = nullptr;
    bool prepared = false;
    oidnb200_conv* handle = nullptr;
    void* packedWeight = nullptr;
    void* packedBias = nullptr;
    const void* boundSrc1 = nullptr;
    const void* boundSrc2 = nullptr;
    void* boundDst = nullptr;
  };

  class B200Pool final : public Pool
  {
  public:
    B200Pool(B200Engine* engine, const PoolDesc& desc) : Pool(desc), engine(engine) {}
    Engine* getEngine() const override { return engine; }
    void submitKernels(const Ref<CancellationToken>&) override
    {
      if (!src || !dst)
        throw std::logic_error("pooling source/destination not set");
      checkKernel(oidnb200_pool_launch(src->getPtr(), srcDesc.getH(), srcDesc.getW(), srcDesc.getPaddedC(), dst->getPtr(),
                                       engine->getStream()), "pool");
    }
  private:
    B200Engine* engine;
  };

  class B200Upsample final : public Upsample
  {
  public:
    B200Upsample(B200Engine* engine, const UpsampleDesc& desc) : Upsample(desc), engine(engine) {}
    Engine* getEngine() const override { return engine; }
    void submitKernels(const Ref<CancellationToken>&) override
    {
      if (!src || !dst)
        throw std::logic_error("upsampling source/destination not set");
      checkKernel(oidnb200_upsample_launch(src->getPtr(), srcDesc.getH(), srcDesc.getW(), srcDesc.getPaddedC(), dst->getPtr(),
                                           engine->getStream()), "upsample");
    }
  private:
    B200Engine* engine;
  };

  class B200InputProcess final : public InputProcess
  {
  public:
    B200InputProcess(B200Engine* engine, const InputProcessDesc& desc) : InputProcess(engine, desc), engine(engine) {}
    Engine* getEngine() const override { return engine; }
    void submitKernels(const Ref<CancellationToken>&) override
    {
      check();
      // the kernel ABI's first image is the main input: color, else albedo, else normal (InputProcess::getMainSrc)
      const oidnb200_image none = toABI(Ref<Image>());
      oidnb200_image c = toABI(color), a = toABI(albedo), n = toABI(normal);
      if (!color)
      {
        c = albedo ? a : n;
        a = n = none;
      }
      const oidnb200_tile t = toABI(tile);
      const oidnb200_transfer tf = toABI(*transferFunc);
      checkKernel(oidnb200_input_process_launch(&c, &a, &n, &t, &tf, hdr, snorm, dst->getPtr(), dstDesc.getH(),
                                                dstDesc.getW(), dstDesc.getPaddedC(), engine->getStream()), "input process");
    }
  private:
    B200Engine* engine;
  };

  class B200OutputProcess final : public OutputProcess
  {
  public:
    B200OutputProcess(B200Engine* engine, const OutputProcessDesc& desc) : OutputProcess(desc), engine(engine) {}
    ~B200OutputProcess()
    {
      if (producer)
        producer->outputProcess = nullptr;
      if (countedSrc)
      {
        auto itc = engine->readers.find(countedSrc);
        if (itc != engine->readers.end() && --itc->second <= 0)
          engine->readers.erase(itc);
      }
    }
    Engine* getEngine() const override { return engine; }

    void finalize() override
    {
      if (countedSrc)
      {
        auto itc = engine->readers.find(countedSrc);
        if (itc != engine->readers.end() && --itc->second <= 0)
          engine->readers.erase(itc);
      }
      countedSrc = src ? src.get() : nullptr;
      if (countedSrc)
        engine->readers[countedSrc] += 1;
      if (producer)
        producer->outputProcess = nullptr;
      producer = nullptr;
      auto it = src ? engine->producerOf.find(src.get()) : engine->producerOf.end();
      if (it != engine->producerOf.end() && it->second->canFuseOutput())
      {
        producer = it->second;
        producer->outputProcess = this;
      }
    }

    // Called by the producing conv right before its launch: folds this op (current tile, image, transfer function)
    // into the conv's epilogue when the kernel supports the image. Returns false -> the conv stores its tensor and
    // this op runs as a pass of its own.
    bool fuseInto(oidnb200_conv* conv)
    {
      fusedThisSubmit = false;
      if (src && dst)
      {
        check();
        const oidnb200_image d = toABI(dst);
        const oidnb200_tile t = toABI(tile);
        const oidnb200_transfer tf = toABI(*transferFunc);
        const int rc = oidnb200_conv_set_output_process(conv, &t, &tf, hdr, snorm, &d);
        if (rc == 0)
          fusedThisSubmit = true;
        else if (rc != OIDNB200_ERR_UNSUPPORTED)
          checkKernel(rc, "fused output process");
      }
      if (!fusedThisSubmit)
        oidnb200_conv_set_output_process(conv, nullptr, nullptr, 0, 0, nullptr);
      return fusedThisSubmit;
    }

    void producerGone() { producer = nullptr; }

    void submitKernels(const Ref<CancellationToken>&) override
    {
      if (fusedThisSubmit)
      {
        fusedThisSubmit = false; // the conv's epilogue has written this tile
        return;
      }
      check();
      const oidnb200_image d = toABI(dst);
      const oidnb200_tile t = toABI(tile);
      const oidnb200_transfer tf = toABI(*transferFunc);
      checkKernel(oidnb200_output_process_launch(src->getPtr(), srcDesc.getH(), srcDesc.getW(), srcDesc.getPaddedC(), &t, &tf,
                                                 hdr, snorm, &d, engine->getStream()), "output process");
    }
  private:
    B200Engine* engine;
    B200Conv* producer = nullptr;
    const Tensor* countedSrc = nullptr;
    bool fusedThisSubmit = false;
  };

  void B200Conv::detachOutputProcess()
  {
    if (outputProcess)
      outputProcess->producerGone();
    outputProcess = nullptr;
  }

  void B200Conv::submitKernels(const Ref<CancellationToken>&)
  {
    if (absorbed)
      return;
    if (!pairChecked)
      tryPair();       // every op of the graph has been finalized by the first submit
    if (pair)
      return;          // the second conv of the pair launches both
    prepare();
    bind();
    const bool fusedOut = outputProcess && outputProcess->fuseInto(handle);
    if (pairA)
    {
      B200Conv* a = pairA;
      a->prepare(); a->bind();
      // The core's arena may alias this conv's destination with the first conv's source (dead by now under two
      // launches, still being read under one): then run them one by one. With the output process in the epilogue
      // the destination tensor is not written at all.
      const char* s0 = static_cast<const char*>(a->src->getPtr()); const char* s1 = s0 + a->src->getByteSize();
      const char* d0 = static_cast<const char*>(dst->getPtr());    const char* d1 = d0 + dst->getByteSize();
      if (fusedOut || d1 <= s0 || s1 <= d0)
      {
        checkKernel(oidnb200_conv_pair_bind(a->pair), "conv pair bind");
        checkKernel(oidnb200_conv_pair_launch(a->pair, engine->getStream()), "conv pair");
        return;
      }
      checkKernel(oidnb200_conv_launch(a->handle, engine->getStream()), "conv");
    }
    checkKernel(oidnb200_conv_launch(handle, engine->getStream()), "conv");
  }

  class B200Autoexposure final : public Autoexposure
  {
  public:
    B200Autoexposure(B200Engine* engine, const ImageDesc& srcDesc) : Autoexposure(srcDesc), engine(engine) {}
    Engine* getEngine() const override { return engine; }
    size_t getScratchByteSize() override { return oidnb200_autoexposure_scratch_bytes(srcDesc.getH(), srcDesc.getW()); }
    void setScratch(const Ref<Buffer>& buffer) override { scratch = buffer; }
    void submitKernels(const Ref<CancellationToken>&) override
    {
      if (!src || !dst || !scratch)
        throw std::logic_error("autoexposure source/destination/scratch not set");
      const oidnb200_image s = toABI(src);
      checkKernel(oidnb200_autoexposure_launch(&s, scratch->getPtr(), getDstPtr(), engine->getStream()), "autoexposure");
    }
  private:
    B200Engine* engine;
    Ref<Buffer> scratch;
  };

  class B200ImageCopy final : public ImageCopy
  {
  public:
    explicit B200ImageCopy(B200Engine* engine) : engine(engine) {}
    Engine* getEngine() const override { return engine; }
    void submitKernels(const Ref<CancellationToken>&) override
    {
      check();
      const oidnb200_image s = toABI(src), d = toABI(dst);
      checkKernel(oidnb200_image_copy_launch(&s, &d, engine->getStream()), "image copy");
    }
  private:
    B200Engine* engine;
  };

  Ref<Conv> B200Engine::newConv(const ConvDesc& desc) { return makeRef<B200Conv>(this, desc); }
  Ref<Pool> B200Engine::newPool(const PoolDesc& desc) { return makeRef<B200Pool>(this, desc); }
  Ref<Upsample> B200Engine::newUpsample(const UpsampleDesc& desc) { return makeRef<B200Upsample>(this, desc); }
  Ref<Autoexposure> B200Engine::newAutoexposure(const ImageDesc& srcDesc) { return makeRef<B200Autoexposure>(this, srcDesc); }
  Ref<InputProcess> B200Engine::newInputProcess(const InputProcessDesc& desc) { return makeRef<B200InputProcess>(this, desc); }
  Ref<OutputProcess> B200Engine::newOutputProcess(const OutputProcessDesc& desc) { return makeRef<B200OutputProcess>(this, desc); }
  Ref<ImageCopy> B200Engine::newImageCopy() { return makeRef<B200ImageCopy>(this); }
#endif // OIDN_B200_OP_LEVEL

  class B200DeviceFactory final : public CUDADeviceFactoryBase
  {
  public:
    bool isDeviceSupported(int deviceID) override { return B200Device::isSupported(deviceID); }

    Ref<Device> newDevice(const int* deviceIDs, const cudaStream_t* streams, int numPairs) override
    {
      // one (device, stream) pair through this entry point; several GPUs: oidnb200NewCUDADevice directly, or one
      // process per GPU with the filter parameters numShards / shardIndex (oidn_b200/sharded.py)
      if (numPairs != 1)
        throw Exception(Error::InvalidArgument, "invalid number of CUDA devices/streams");
      if (deviceIDs == nullptr)
        throw Exception(Error::InvalidArgument, "array of CUDA devices is null");
      if (streams == nullptr)
        throw Exception(Error::InvalidArgument, "array of CUDA streams is null");
      return makeRef<B200Device>(deviceIDs[0], streams[0]);
    }

    Ref<Device> newDevice(const Ref<PhysicalDevice>& physicalDevice) override
    {
      return makeRef<B200Device>(staticRefCast<B200PhysicalDevice>(physicalDevice));
    }
  };

  OIDN_DECLARE_INIT_MODULE(device_cuda)
  {
    if (oidnb200GetNumPhysicalDevices() > 0)
      Context::registerDeviceType<B200DeviceFactory>(DeviceType::CUDA, B200Device::getPhysicalDevices());
  }

OIDN_NAMESPACE_END
