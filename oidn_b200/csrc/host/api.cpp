// C entry points of include/oidn_b200.h. Same structure as the reference's api/api.cpp:
// handle cast, per-device mutex around every call (api/api.cpp:42-69), exceptions converted to
// the device's first-error slot (api/api.cpp:17-31).
#include "../../../include/oidn_b200.h"
#include "filter.hpp"
#include <atomic>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstring>
#include <unistd.h>
#include <map>

using namespace oidnb200;

struct oidnb200_device_t
{
  std::atomic<int> refs{1};
  std::unique_ptr<Device> impl;
};

struct oidnb200_buffer_t
{
  std::atomic<int> refs{1};
  oidnb200_device_t* device;
  void* ptr = nullptr;
  size_t size = 0;
  Storage storage = Storage::Device;
  bool imported = false; // opened from a CUDA IPC handle
  // external memory (oidnb200NewSharedBufferFromFD) and exportable memory (oidnb200NewExportableBuffer)
  enum Kind { Plain, External, Vmm } kind = Plain;
  cudaExternalMemory_t ext = nullptr;         // External: imported with cudaImportExternalMemory
  unsigned long long vmmHandle = 0;           // Vmm: CUmemGenericAllocationHandle (created or imported)
  size_t vmmSize = 0;                         // Vmm: mapped (granularity-rounded) size
};

struct oidnb200_filter_t
{
  std::atomic<int> refs{1};
  oidnb200_device_t* device;
  std::shared_ptr<Filter> impl;
  // buffers behind the images set with oidnb200SetFilterImage: the filter keeps them alive as the reference's
  // Image holds a Ref<Buffer> (core/image.h), so releasing a buffer after setting it is legal
  std::map<std::string, oidnb200_buffer_t*> buffers;
};

namespace {

void reportError(oidnb200_device_t* d, Error code, const std::string& msg)
{
  if (d && d->impl) d->impl->setError(code, msg); else Device::setGlobalError(code, msg);
}

template <typename F>
void guarded(oidnb200_device_t* d, F&& f)
{
  try
  {
    if (!d || !d->impl) throw Exception(Error::InvalidArgument, "invalid handle");
    std::lock_guard<std::mutex> lock(d->impl->getMutex());
    f();
  }
  catch (const Exception& e) { reportError(d, e.code(), e.what()); }
  catch (const std::bad_alloc&) { reportError(d, Error::OutOfMemory, "out of memory"); }
  catch (const std::exception& e) { reportError(d, Error::Unknown, e.what()); }
  catch (...) { reportError(d, Error::Unknown, "unknown exception caught"); }
}

void retainDevice(oidnb200_device_t* d) { d->refs.fetch_add(1); }
void releaseDevice(oidnb200_device_t* d)
{
  if (d->refs.fetch_sub(1) == 1)
  {
    try { if (d->impl && d->impl->isCommitted()) d->impl->wait(); } catch (...) {}
    delete d;
  }
}

// CUDA virtual-memory-management entry points (driver API, resolved at run time: the library links only the
// static runtime). Used for device memory that can be exported as / imported from an opaque POSIX fd.
struct VmmApi
{
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addrFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*exportHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*importHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};

const VmmApi& vmmApi()
{
  static VmmApi api;
  static bool tried = false;
  if (!tried)
  {
    tried = true;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    api.ok = get("cuMemCreate", (void**)&api.create) && get("cuMemRelease", (void**)&api.release) &&
             get("cuMemAddressReserve", (void**)&api.reserve) && get("cuMemAddressFree", (void**)&api.addrFree) &&
             get("cuMemMap", (void**)&api.map) && get("cuMemUnmap", (void**)&api.unmap) &&
             get("cuMemSetAccess", (void**)&api.setAccess) && get("cuMemExportToShareableHandle", (void**)&api.exportHandle) &&
             get("cuMemImportFromShareableHandle", (void**)&api.importHandle) &&
             get("cuMemGetAllocationGranularity", (void**)&api.granularity);
    cudaGetLastError();
  }
  return api;
}

void checkCu(CUresult r, const char* what)
{
  if (r == CUDA_SUCCESS) return;
  if (r == CUDA_ERROR_OUT_OF_MEMORY) throw Exception(Error::OutOfMemory, std::string(what) + ": out of memory");
  throw Exception(Error::Unknown, std::string(what) + ": CUDA driver error " + std::to_string((int)r));
}

CUmemAllocationProp vmmProp(int deviceID)
{
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = deviceID;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

// Maps a VMM allocation handle into this process's address space with read/write access for the device.
void* vmmMapHandle(CUmemGenericAllocationHandle h, size_t size, int deviceID)
{
  const VmmApi& v = vmmApi();
  CUdeviceptr va = 0;
  checkCu(v.reserve(&va, size, 0, 0, 0), "cuMemAddressReserve");
  CUresult r = v.map(va, size, 0, h, 0);
  if (r == CUDA_SUCCESS)
  {
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = deviceID;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = v.setAccess(va, size, &acc, 1);
    if (r != CUDA_SUCCESS) v.unmap(va, size);
  }
  if (r != CUDA_SUCCESS) { v.addrFree(va, size); checkCu(r, "cuMemMap"); }
  return reinterpret_cast<void*>(va);
}

// Frees the buffer's memory; the device mutex is held by the caller.
void destroyBufferLocked(oidnb200_buffer_t* b)
{
  Device* dev = b->device->impl.get();
  dev->wait();
  if (b->kind == oidnb200_buffer_t::External)
  {
    dev->getEngine(0)->makeCurrent();
    cudaFree(b->ptr);                       // devices/cuda/cuda_external_buffer.cpp:86-90
    cudaDestroyExternalMemory(b->ext);
  }
  else if (b->kind == oidnb200_buffer_t::Vmm)
  {
    dev->getEngine(0)->makeCurrent();
    const VmmApi& v = vmmApi();
    v.unmap((CUdeviceptr)b->ptr, b->vmmSize);
    v.addrFree((CUdeviceptr)b->ptr, b->vmmSize);
    v.release(b->vmmHandle);
  }
  else if (b->imported) { dev->getEngine(0)->makeCurrent(); cudaIpcCloseMemHandle(b->ptr); }
  else if (b->storage == Storage::Host) dev->freeHost(b->ptr);
  else dev->getEngine(0)->free(b->ptr, b->storage);
}

// Drops a reference held by a filter (device mutex held; the filter's own device reference keeps the device alive).
void dropBufferLocked(oidnb200_buffer_t* b)
{
  if (b && b->refs.fetch_sub(1) == 1)
  {
    oidnb200_device_t* d = b->device;
    try { destroyBufferLocked(b); } catch (...) {}
    delete b;
    d->refs.fetch_sub(1);
  }
}

void holdBuffer(oidnb200_filter_t* f, const std::string& name, oidnb200_buffer_t* b)
{
  if (b) b->refs.fetch_add(1);
  auto it = f->buffers.find(name);
  oidnb200_buffer_t* old = it == f->buffers.end() ? nullptr : it->second;
  if (b) f->buffers[name] = b; else if (it != f->buffers.end()) f->buffers.erase(it);
  dropBufferLocked(old);
}

void checkString(const char* s)
{
  if (!s) throw Exception(Error::InvalidArgument, "string pointer is null");
}

Image makeImage(void* base, int format, size_t width, size_t height, size_t byteOffset, size_t pixelStride,
                size_t rowStride)
{
  // core/image.cpp:41-48 + ImageDesc checks (core/image.cpp:11-39)
  Image im;
  const Format f = static_cast<Format>(format);
  if (formatChannels(f) == 0 && base) throw Exception(Error::InvalidArgument, "invalid image format");
  if (width > 65536 || height > 65536) throw Exception(Error::InvalidArgument, "image size too large");
  const size_t px = formatBytes(f);
  if (pixelStride == 0) pixelStride = px;
  else if (pixelStride < px) throw Exception(Error::InvalidArgument, "pixel stride smaller than pixel size");
  if (rowStride == 0) rowStride = width * pixelStride;
  else if (rowStride < width * pixelStride) throw Exception(Error::InvalidArgument, "row stride smaller than width * pixel stride");
  if (base && (pixelStride % (formatIsHalf(f) ? 2 : 4) || rowStride % (formatIsHalf(f) ? 2 : 4) ||
               (reinterpret_cast<uintptr_t>(base) + byteOffset) % (formatIsHalf(f) ? 2 : 4)))
    throw Exception(Error::InvalidArgument, "image pointer and strides must be aligned to the channel type");
  im.ptr = base ? static_cast<uint8_t*>(base) + byteOffset : nullptr;
  im.format = f;
  im.W = (int)width; im.H = (int)height;
  im.pixelStride = pixelStride; im.rowStride = rowStride;
  return im;
}

} // namespace

extern "C" {

int oidnb200GetNumPhysicalDevices(void) { return oidnb200_device_count(); }

OIDNB200Device oidnb200NewCUDADevice(const int* deviceIDs, void* const* streams, int numPairs)
{
  try
  {
    if (!deviceIDs || numPairs < 1) throw Exception(Error::InvalidArgument, "invalid number of CUDA device/stream pairs");
    std::vector<int> ids(deviceIDs, deviceIDs + numPairs);
    std::vector<void*> st(numPairs, nullptr);
    if (streams) for (int i = 0; i < numPairs; ++i) st[i] = streams[i];
    auto* d = new oidnb200_device_t();
    d->impl.reset(new Device(ids, st));
    return d;
  }
  catch (const Exception& e) { Device::setGlobalError(e.code(), e.what()); }
  catch (const std::exception& e) { Device::setGlobalError(Error::Unknown, e.what()); }
  return nullptr;
}

OIDNB200Device oidnb200NewDevice(void)
{
  const int id = 0;
  return oidnb200NewCUDADevice(&id, nullptr, 1);
}

void oidnb200RetainDevice(OIDNB200Device d) { if (d) retainDevice(d); }
void oidnb200ReleaseDevice(OIDNB200Device d) { if (d) releaseDevice(d); }

void oidnb200SetDeviceInt(OIDNB200Device d, const char* name, int value)
{
  guarded(d, [&] { checkString(name); d->impl->setInt(name, value); });
}

int oidnb200GetDeviceInt(OIDNB200Device d, const char* name)
{
  int v = 0;
  guarded(d, [&] { checkString(name); v = d->impl->getInt(name); });
  return v;
}

void oidnb200SetDeviceString(OIDNB200Device d, const char* name, const char* value)
{
  guarded(d, [&] { checkString(name); checkString(value); d->impl->setString(name, value); });
}

void oidnb200CommitDevice(OIDNB200Device d) { guarded(d, [&] { d->impl->commit(); }); }
void oidnb200SyncDevice(OIDNB200Device d) { guarded(d, [&] { d->impl->checkCommitted(); d->impl->wait(); }); }

int oidnb200GetDeviceError(OIDNB200Device d, const char** outMessage)
{
  if (!d || !d->impl) return (int)Device::getGlobalError(outMessage);
  std::lock_guard<std::mutex> lock(d->impl->getMutex());
  return (int)d->impl->getError(outMessage);
}

// ---- buffers ----------------------------------------------------------------------------------
OIDNB200Buffer oidnb200NewBufferWithStorage(OIDNB200Device d, size_t byteSize, int storage)
{
  oidnb200_buffer_t* b = nullptr;
  guarded(d, [&] {
    d->impl->checkCommitted();
    const Storage s = static_cast<Storage>(storage);
    if (s != Storage::Host && s != Storage::Device && s != Storage::Managed)
      throw Exception(Error::InvalidArgument, "invalid storage mode");
    void* p = s == Storage::Host ? d->impl->allocHost(byteSize) : d->impl->getEngine(0)->malloc(byteSize, s);
    b = new oidnb200_buffer_t();
    b->device = d; b->ptr = p; b->size = byteSize; b->storage = s;
    retainDevice(d);
  });
  return b;
}

OIDNB200Buffer oidnb200NewBuffer(OIDNB200Device d, size_t byteSize)
{
  return oidnb200NewBufferWithStorage(d, byteSize, (int)Storage::Device);
}

void* oidnb200GetBufferData(OIDNB200Buffer b) { return b ? b->ptr : nullptr; }
size_t oidnb200GetBufferSize(OIDNB200Buffer b) { return b ? b->size : 0; }

static void bufferCopy(oidnb200_buffer_t* b, size_t off, size_t n, void* dst, const void* src, bool read, bool sync)
{
  if (!b) return;
  guarded(b->device, [&] {
    if (off + n > b->size || off + n < off) throw Exception(Error::InvalidArgument, "buffer region is out of bounds");
    if (n == 0) return;
    if ((read && !dst) || (!read && !src)) throw Exception(Error::InvalidArgument, "host pointer is null");
    b->device->impl->joinStaged();
    Engine* e = b->device->impl->getEngine(0);
    uint8_t* p = static_cast<uint8_t*>(b->ptr) + off;
    if (read) e->submitCopy(dst, p, n); else e->submitCopy(p, src, n);
    if (sync) e->wait();
  });
}

void oidnb200ReadBuffer(OIDNB200Buffer b, size_t off, size_t n, void* dst) { bufferCopy(b, off, n, dst, nullptr, true, true); }
void oidnb200WriteBuffer(OIDNB200Buffer b, size_t off, size_t n, const void* src) { bufferCopy(b, off, n, nullptr, src, false, true); }
void oidnb200ReadBufferAsync(OIDNB200Buffer b, size_t off, size_t n, void* dst) { bufferCopy(b, off, n, dst, nullptr, true, false); }
void oidnb200WriteBufferAsync(OIDNB200Buffer b, size_t off, size_t n, const void* src) { bufferCopy(b, off, n, nullptr, src, false, false); }

void oidnb200CopyRectAsync(OIDNB200Device d, void* dst, size_t dstPitch, const void* src, size_t srcPitch,
                           size_t widthBytes, size_t height)
{
  guarded(d, [&] {
    d->impl->checkCommitted();
    if (widthBytes == 0 || height == 0) return;
    if (!dst || !src) throw Exception(Error::InvalidArgument, "pointer is null");
    if (dstPitch < widthBytes || srcPitch < widthBytes) throw Exception(Error::InvalidArgument, "pitch is smaller than the row");
    d->impl->joinStaged();
    d->impl->getEngine(0)->submitCopy2D(dst, dstPitch, src, srcPitch, widthBytes, height);
  });
}

void oidnb200ReleaseBuffer(OIDNB200Buffer b)
{
  if (!b) return;
  if (b->refs.fetch_sub(1) == 1)
  {
    oidnb200_device_t* d = b->device;
    guarded(d, [&] { destroyBufferLocked(b); });
    delete b;
    releaseDevice(d);
  }
}

void oidnb200GetBufferIpcHandle(OIDNB200Buffer b, void* outHandle64)
{
  if (!b) return;
  guarded(b->device, [&] {
    if (!outHandle64) throw Exception(Error::InvalidArgument, "handle pointer is null");
    if (b->storage != Storage::Device || b->imported || b->kind != oidnb200_buffer_t::Plain)
      throw Exception(Error::InvalidOperation, "only device-storage buffers owned by this process can be exported");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    b->device->impl->getEngine(0)->makeCurrent();
    checkCuda(cudaIpcGetMemHandle(&h, b->ptr), "cudaIpcGetMemHandle");
    memcpy(outHandle64, &h, 64);
  });
}

OIDNB200Buffer oidnb200NewSharedBufferFromIpcHandle(OIDNB200Device d, const void* handle64, size_t byteSize)
{
  oidnb200_buffer_t* b = nullptr;
  guarded(d, [&] {
    d->impl->checkCommitted();
    if (!handle64) throw Exception(Error::InvalidArgument, "handle pointer is null");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    d->impl->getEngine(0)->makeCurrent();
    checkCuda(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    b = new oidnb200_buffer_t();
    b->device = d; b->ptr = p; b->size = byteSize; b->storage = Storage::Device; b->imported = true;
    retainDevice(d);
  });
  return b;
}

OIDNB200Buffer oidnb200NewSharedBufferFromFD(OIDNB200Device d, int fdType, int fd, size_t byteSize)
{
  oidnb200_buffer_t* b = nullptr;
  guarded(d, [&] {
    d->impl->checkCommitted();
    // the CUDA device supports opaque fds only (devices/cuda/cuda_external_buffer.cpp:13-14, api/api.cpp:603-604)
    if (fdType != OIDNB200_EXTERNAL_MEMORY_TYPE_FLAG_OPAQUE_FD)
      throw Exception(Error::InvalidArgument, "external memory type not supported by the device");
    if (fd < 0 || byteSize == 0) throw Exception(Error::InvalidArgument, "invalid file descriptor or size");
    Engine* e = d->impl->getEngine(0);
    e->makeCurrent();
    // 1. an external-memory object of another API (Vulkan, ...): cudaImportExternalMemory owns the fd on success
    cudaExternalMemoryHandleDesc hd{};
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    hd.size = byteSize;
    cudaExternalMemory_t ext = nullptr;
    if (cudaImportExternalMemory(&ext, &hd) == cudaSuccess)
    {
      void* p = nullptr;
      cudaExternalMemoryBufferDesc bd{};
      bd.offset = 0; bd.size = byteSize; bd.flags = 0;
      const cudaError_t me = cudaExternalMemoryGetMappedBuffer(&p, ext, &bd);
      if (me != cudaSuccess) { cudaDestroyExternalMemory(ext); checkCuda(me, "cudaExternalMemoryGetMappedBuffer"); }
      b = new oidnb200_buffer_t();
      b->device = d; b->ptr = p; b->size = byteSize; b->storage = Storage::Device;
      b->kind = oidnb200_buffer_t::External; b->ext = ext;
      retainDevice(d);
      return;
    }
    cudaGetLastError();
    // 2. a CUDA allocation exported with cuMemExportToShareableHandle (oidnb200GetBufferFD, another CUDA process)
    const VmmApi& v = vmmApi();
    if (!v.ok) throw Exception(Error::InvalidArgument, "the file descriptor is not an importable memory object");
    CUmemGenericAllocationHandle h = 0;
    if (v.importHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS)
      throw Exception(Error::InvalidArgument, "the file descriptor is not an importable memory object");
    size_t gran = 0;
    const CUmemAllocationProp prop = vmmProp(e->getDeviceID());
    checkCu(v.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM), "cuMemGetAllocationGranularity");
    const size_t mapped = round_up(byteSize, gran);
    void* p = nullptr;
    try { p = vmmMapHandle(h, mapped, e->getDeviceID()); }
    catch (...) { v.release(h); throw; }
    close(fd); // ownership passes to the buffer, as with cudaImportExternalMemory
    b = new oidnb200_buffer_t();
    b->device = d; b->ptr = p; b->size = byteSize; b->storage = Storage::Device;
    b->kind = oidnb200_buffer_t::Vmm; b->vmmHandle = h; b->vmmSize = mapped;
    retainDevice(d);
  });
  return b;
}

OIDNB200Buffer oidnb200NewExportableBuffer(OIDNB200Device d, size_t byteSize)
{
  oidnb200_buffer_t* b = nullptr;
  guarded(d, [&] {
    d->impl->checkCommitted();
    if (byteSize == 0) throw Exception(Error::InvalidArgument, "buffer size is zero");
    const VmmApi& v = vmmApi();
    if (!v.ok) throw Exception(Error::UnsupportedHardware, "CUDA virtual memory management is not available");
    Engine* e = d->impl->getEngine(0);
    e->makeCurrent();
    cudaFree(nullptr); // the primary context must exist before driver-API calls
    const CUmemAllocationProp prop = vmmProp(e->getDeviceID());
    size_t gran = 0;
    checkCu(v.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM), "cuMemGetAllocationGranularity");
    const size_t mapped = round_up(byteSize, gran);
    CUmemGenericAllocationHandle h = 0;
    checkCu(v.create(&h, mapped, &prop, 0), "cuMemCreate");
    void* p = nullptr;
    try { p = vmmMapHandle(h, mapped, e->getDeviceID()); }
    catch (...) { v.release(h); throw; }
    b = new oidnb200_buffer_t();
    b->device = d; b->ptr = p; b->size = byteSize; b->storage = Storage::Device;
    b->kind = oidnb200_buffer_t::Vmm; b->vmmHandle = h; b->vmmSize = mapped;
    retainDevice(d);
  });
  return b;
}

int oidnb200GetBufferFD(OIDNB200Buffer b)
{
  int fd = -1;
  if (!b) return fd;
  guarded(b->device, [&] {
    if (b->kind != oidnb200_buffer_t::Vmm)
      throw Exception(Error::InvalidOperation, "only buffers created with oidnb200NewExportableBuffer can be exported");
    b->device->impl->getEngine(0)->makeCurrent();
    checkCu(vmmApi().exportHandle(&fd, b->vmmHandle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
  });
  return fd;
}

// ---- filters ----------------------------------------------------------------------------------
OIDNB200Filter oidnb200NewFilter(OIDNB200Device d, const char* type)
{
  oidnb200_filter_t* f = nullptr;
  guarded(d, [&] {
    checkString(type);
    auto impl = d->impl->newFilter(type);
    f = new oidnb200_filter_t();
    f->device = d; f->impl = impl;
    retainDevice(d);
  });
  return f;
}

void oidnb200RetainFilter(OIDNB200Filter f) { if (f) f->refs.fetch_add(1); }

void oidnb200ReleaseFilter(OIDNB200Filter f)
{
  if (!f) return;
  if (f->refs.fetch_sub(1) == 1)
  {
    oidnb200_device_t* d = f->device;
    guarded(d, [&] {
      d->impl->wait(); f->impl.reset(); // api/api.cpp:111-141: wait for idle, then destroy
      for (auto& kv : f->buffers) dropBufferLocked(kv.second);
      f->buffers.clear();
    });
    delete f;
    releaseDevice(d);
  }
}

void oidnb200SetFilterImage(OIDNB200Filter f, const char* name, OIDNB200Buffer buffer, int format, size_t width,
                            size_t height, size_t byteOffset, size_t pixelByteStride, size_t rowByteStride)
{
  if (!f) return;
  guarded(f->device, [&] {
    checkString(name);
    if (!buffer) throw Exception(Error::InvalidArgument, "buffer is null");
    Image im = makeImage(buffer->ptr, format, width, height, byteOffset, pixelByteStride, rowByteStride);
    if (im && im.end() > static_cast<uint8_t*>(buffer->ptr) + buffer->size)
      throw Exception(Error::InvalidArgument, "buffer region is out of bounds");
    f->impl->setImage(name, im);
    holdBuffer(f, name, im ? buffer : nullptr);
  });
}

void oidnb200SetSharedFilterImage(OIDNB200Filter f, const char* name, void* devPtr, int format, size_t width,
                                  size_t height, size_t byteOffset, size_t pixelByteStride, size_t rowByteStride)
{
  if (!f) return;
  guarded(f->device, [&] {
    checkString(name);
    f->impl->setImage(name, makeImage(devPtr, format, width, height, byteOffset, pixelByteStride, rowByteStride));
    holdBuffer(f, name, nullptr);
  });
}

void oidnb200UnsetFilterImage(OIDNB200Filter f, const char* name)
{
  if (f) guarded(f->device, [&] { checkString(name); f->impl->unsetImage(name); holdBuffer(f, name, nullptr); });
}

void oidnb200SetSharedFilterData(OIDNB200Filter f, const char* name, void* hostPtr, size_t byteSize)
{
  if (!f) return;
  guarded(f->device, [&] {
    checkString(name);
    if (!hostPtr && byteSize) throw Exception(Error::InvalidArgument, "data pointer is null but the size is not zero");
    Data d;
    d.ptr = byteSize ? hostPtr : nullptr; d.size = byteSize;
    f->impl->setData(name, d);
  });
}

void oidnb200UpdateFilterData(OIDNB200Filter f, const char* name)
{
  if (f) guarded(f->device, [&] { checkString(name); f->impl->updateData(name); });
}

void oidnb200UnsetFilterData(OIDNB200Filter f, const char* name)
{
  if (f) guarded(f->device, [&] { checkString(name); f->impl->unsetData(name); });
}

void oidnb200SetFilterBool(OIDNB200Filter f, const char* name, bool value)
{
  if (f) guarded(f->device, [&] { checkString(name); f->impl->setInt(name, value ? 1 : 0); });
}

bool oidnb200GetFilterBool(OIDNB200Filter f, const char* name)
{
  int v = 0;
  if (f) guarded(f->device, [&] { checkString(name); v = f->impl->getInt(name); });
  return v != 0;
}

void oidnb200SetFilterInt(OIDNB200Filter f, const char* name, int value)
{
  if (f) guarded(f->device, [&] { checkString(name); f->impl->setInt(name, value); });
}

int oidnb200GetFilterInt(OIDNB200Filter f, const char* name)
{
  int v = 0;
  if (f) guarded(f->device, [&] { checkString(name); v = f->impl->getInt(name); });
  return v;
}

void oidnb200SetFilterFloat(OIDNB200Filter f, const char* name, float value)
{
  if (f) guarded(f->device, [&] { checkString(name); f->impl->setFloat(name, value); });
}

float oidnb200GetFilterFloat(OIDNB200Filter f, const char* name)
{
  float v = 0.f;
  if (f) guarded(f->device, [&] { checkString(name); v = f->impl->getFloat(name); });
  return v;
}

void oidnb200SetFilterProgressMonitorFunction(OIDNB200Filter f, OIDNB200ProgressMonitorFunction func, void* userPtr)
{
  if (f) guarded(f->device, [&] { f->impl->setProgressMonitorFunction(func, userPtr); });
}

void oidnb200CommitFilter(OIDNB200Filter f) { if (f) guarded(f->device, [&] { f->impl->commit(); }); }
void oidnb200ExecuteFilter(OIDNB200Filter f) { if (f) guarded(f->device, [&] { f->impl->execute(SyncMode::Blocking); }); }
void oidnb200ExecuteFilterAsync(OIDNB200Filter f) { if (f) guarded(f->device, [&] { f->impl->execute(SyncMode::Async); }); }

void oidnb200GetFilterInfo(OIDNB200Filter f, oidnb200_filter_info* info)
{
  if (!f || !info) return;
  memset(info, 0, sizeof(*info));
  guarded(f->device, [&] {
    auto* u = dynamic_cast<UNetFilter*>(f->impl.get());
    if (!u) return;
    const TilePlan& p = u->getTilePlan();
    info->tileH = p.tileH; info->tileW = p.tileW; info->tileCountH = p.tileCountH; info->tileCountW = p.tileCountW;
    info->tileOverlap = p.tileOverlap; info->tileAlignment = p.tileAlignment;
    info->largeModel = u->isLargeModel();
    info->numOps = u->getNumOps();
    info->staged = u->wasStaged();
    info->memoryBytes = u->getScratchByteSize();
  });
}

int oidnb200GetFilterProfile(OIDNB200Filter f, oidnb200_op_time* out, int maxOps)
{
  int n = 0;
  if (!f) return 0;
  guarded(f->device, [&] {
    auto* u = dynamic_cast<UNetFilter*>(f->impl.get());
    if (!u) return;
    const auto prof = u->getProfile();
    n = (int)prof.size();
    for (int i = 0; i < n && i < maxOps; ++i)
    {
      memset(&out[i], 0, sizeof(out[i]));
      strncpy(out[i].name, prof[i].name.c_str(), sizeof(out[i].name) - 1);
      out[i].kind = prof[i].kind; out[i].launches = prof[i].launches; out[i].ms = prof[i].ms;
    }
  });
  return n;
}

void oidnb200ResetFilterProfile(OIDNB200Filter f)
{
  if (!f) return;
  guarded(f->device, [&] { if (auto* u = dynamic_cast<UNetFilter*>(f->impl.get())) u->resetProfile(); });
}

void oidnb200PlanTiles(int H, int W, int largeModel, int deviceMinAlignment, int numEngines, long maxTilePixels,
                       oidnb200_tile_plan* out)
{
  const TilePlan p = planTiles(H, W, largeModel != 0, deviceMinAlignment, numEngines > 0 ? numEngines : 1,
                               maxTilePixels > 0 ? maxTilePixels : 2160L * 2160L, [](const TilePlan&) { return true; });
  *out = oidnb200_tile_plan{p.H, p.W, p.tileH, p.tileW, p.tilePadH, p.tilePadW, p.tileCountH, p.tileCountW,
                            p.tileAlignment, p.tileOverlap};
}

void oidnb200PlanTilesMinOverlap(int H, int W, int largeModel, int deviceMinAlignment, int numUnits, long maxTilePixels,
                                 oidnb200_tile_plan* out)
{
  const TilePlan p = planTilesMinOverlap(H, W, largeModel != 0, deviceMinAlignment, numUnits > 0 ? numUnits : 1,
                                         maxTilePixels > 0 ? maxTilePixels : 7680L * 4352L, [](const TilePlan&) { return true; });
  *out = oidnb200_tile_plan{p.H, p.W, p.tileH, p.tileW, p.tilePadH, p.tilePadW, p.tileCountH, p.tileCountW,
                            p.tileAlignment, p.tileOverlap};
}

void oidnb200PlanTilesStripAware(int H, int W, int largeModel, int deviceMinAlignment, int numUnits, long maxTilePixels,
                                  oidnb200_tile_plan* out)
{
  const TilePlan p = planTilesMinOverlap(H, W, largeModel != 0, deviceMinAlignment, numUnits > 0 ? numUnits : 1,
                                         maxTilePixels > 0 ? maxTilePixels : 7680L * 4352L, [](const TilePlan&) { return true; }, true);
  *out = oidnb200_tile_plan{p.H, p.W, p.tileH, p.tileW, p.tilePadH, p.tilePadW, p.tileCountH, p.tileCountW,
                            p.tileAlignment, p.tileOverlap};
}

int oidnb200EnumerateTiles(const oidnb200_tile_plan* pl, int* out, int maxTiles)
{
  TilePlan p;
  p.H = pl->H; p.W = pl->W; p.tileH = pl->tileH; p.tileW = pl->tileW; p.tilePadH = pl->tilePadH; p.tilePadW = pl->tilePadW;
  p.tileCountH = pl->tileCountH; p.tileCountW = pl->tileCountW; p.tileAlignment = pl->tileAlignment; p.tileOverlap = pl->tileOverlap;
  const std::vector<TileRect> t = enumerateTiles(p);
  for (int i = 0; i < (int)t.size() && i < maxTiles; ++i)
  {
    const int v[12] = {t[i].hSrc, t[i].wSrc, t[i].hBuf, t[i].wBuf, t[i].H1, t[i].W1,
                       t[i].hOutBuf, t[i].wOutBuf, t[i].hDst, t[i].wDst, t[i].H2, t[i].W2};
    memcpy(out + 12 * i, v, sizeof(v));
  }
  return (int)t.size();
}

int oidnb200ParseTZA(const void* blob, size_t size, const char** outMessage)
{
  static thread_local std::string msg;
  try
  {
    auto m = parseTZA(blob, size);
    if (outMessage) *outMessage = nullptr;
    return (int)m->size();
  }
  catch (const Exception& e)
  {
    msg = e.what();
    if (outMessage) *outMessage = msg.c_str();
    return -(int)e.code();
  }
}

size_t oidnb200PlanArena(int n, const size_t* sizes, const int* firstOp, const int* lastOp, size_t* offsets)
{
  ArenaPlanner pl;
  for (int i = 0; i < n; ++i)
  {
    const int id = pl.newAlloc(firstOp[i], sizes[i]);
    pl.addDep(lastOp[i], id);
  }
  pl.commit();
  for (int i = 0; i < n; ++i) offsets[i] = pl.getAllocByteOffset(i);
  return pl.validate() ? pl.getByteSize() : (size_t)-1;
}

} // extern "C"
