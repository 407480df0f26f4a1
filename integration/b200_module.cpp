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
#include <cuda_runtime.h>
#include <map>

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

    // oidn_b200 keeps the reference's error codes (include/oidn_b200.h) and its first-error-per-device slot
    void checkB200(OIDNB200Device h)
    {
      const char* message = nullptr;
      const int code = oidnb200GetDeviceError(h, &message);
      if (code != 0)
        throw Exception(static_cast<Error>(code), message ? message : "oidn_b200 error");
    }
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

    Ref<Conv> newConv(const ConvDesc&) override { unsupported(); return nullptr; }
    Ref<Pool> newPool(const PoolDesc&) override { unsupported(); return nullptr; }
    Ref<Upsample> newUpsample(const UpsampleDesc&) override { unsupported(); return nullptr; }
    Ref<Autoexposure> newAutoexposure(const ImageDesc&) override { unsupported(); return nullptr; }
    Ref<InputProcess> newInputProcess(const InputProcessDesc&) override { unsupported(); return nullptr; }
    Ref<OutputProcess> newOutputProcess(const OutputProcessDesc&) override { unsupported(); return nullptr; }
    Ref<ImageCopy> newImageCopy() override { unsupported(); return nullptr; }

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
    static void unsupported()
    {
      throw Exception(Error::InvalidOperation, "the oidn_b200 device builds its own graph: core ops are not created through the engine");
    }

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

    // Needs `virtual` on Device::newFilter (the one-word core change)
    Ref<Filter> newFilter(const std::string& type) override;

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

      // the backend runs on the same stream as the core's buffer copies: everything stays stream-ordered
      void* streams[1] = {stream};
      handle = oidnb200NewCUDADevice(&deviceID, streams, 1);
      if (!handle)
        throw Exception(Error::Unknown, "oidnb200NewCUDADevice failed");
      oidnb200SetDeviceInt(handle, "verbose", verbose);
      oidnb200CommitDevice(handle);
      checkB200(handle);
    }

    int deviceID = 0;
    int prevDeviceID = -1;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    OIDNB200Device handle = nullptr;
  };

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
