#include "device.hpp"
#include "filter.hpp"
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace oidnb200 {

namespace {
std::mutex g_globalMutex;
Error g_globalCode = Error::None;
std::string g_globalMessage;
thread_local std::string g_globalMessageOut;
} // namespace

Device::Device(const std::vector<int>& ids, const std::vector<void*>& streams)
  : deviceIDs(ids), userStreams(streams)
{
  if (ids.empty() || ids.size() > 16) throw Exception(Error::InvalidArgument, "invalid number of CUDA device/stream pairs");
  userStreams.resize(ids.size(), nullptr);
  // The same GPU may appear in several pairs: each pair is an engine with its own stream (tiles of a frame
  // then run concurrently on one GPU; also how the multi-engine path is exercised on a single-GPU machine).
  // Up to 8K (7680x4320: a 7.6 GB arena for the base UNet, every tensor < 4 GiB) is one tile on a
  // B200's 180 GB; the reference's default of 2160x2160 (core/unet_filter.h:39) exists for GPUs
  // with little memory. Tiling is then driven by the number of engines / shards only.
  maxTilePixels = 7680L * 4352L;
  if (const char* e = getenv("OIDN_B200_TILE_POLICY")) tilePolicy = atoi(e);
  if (const char* e = getenv("OIDN_B200_FUSE_OUTPUT")) fuseOutput = atoi(e);
  if (const char* e = getenv("OIDN_B200_MAX_TILE_PIXELS")) maxTilePixels = atol(e);
  if (const char* e = getenv("OIDN_B200_GRAPH")) graph = atoi(e);
  if (const char* e = getenv("OIDN_B200_STAGING")) staging = atoi(e);
  if (const char* e = getenv("OIDN_B200_FUSE_PAIRS")) fusePairs = atoi(e);
  if (const char* e = getenv("OIDN_B200_WEIGHTS_DIR")) weightsDir = e;
  if (const char* e = getenv("OIDN_VERBOSE")) verbose = atoi(e);
}

Device::~Device()
{
  for (auto& m : hostMaps)
  {
    cudaHostUnregister(m.first);
    munmap(m.first, m.second);
  }
  for (size_t i = 0; i < events.size(); ++i)
  {
    cudaSetDevice(deviceIDs[i]);
    if (events[i]) cudaEventDestroy(static_cast<cudaEvent_t>(events[i]));
  }
}

void Device::checkCommitted() const
{
  if (!committed) throw Exception(Error::InvalidOperation, "changes to the device are not committed");
}

void Device::commit()
{
  if (committed) throw Exception(Error::InvalidOperation, "device can be committed only once");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
  {
    cudaGetLastError();
    throw Exception(Error::UnsupportedHardware, "no CUDA device found");
  }
  for (int id : deviceIDs)
  {
    if (id < 0 || id >= count) throw Exception(Error::InvalidArgument, "invalid CUDA device ID");
    int major = 0;
    checkCuda(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, id), "cudaDeviceGetAttribute");
    if (major != 10)
      throw Exception(Error::UnsupportedHardware, "the B200 backend needs a compute capability 10.x GPU (sm_100a kernels)");
  }
  for (size_t i = 0; i < deviceIDs.size(); ++i)
  {
    engines.emplace_back(new Engine(deviceIDs[i], userStreams[i]));
    cudaEvent_t ev;
    checkCuda(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate");
    events.push_back(ev);
  }
  // all-pairs peer access: any engine may read the input images / write the output image in place
  for (size_t i = 0; i < deviceIDs.size(); ++i)
    for (size_t j = 0; j < deviceIDs.size(); ++j)
    {
      if (deviceIDs[i] == deviceIDs[j]) continue;
      int can = 0;
      checkCuda(cudaDeviceCanAccessPeer(&can, deviceIDs[i], deviceIDs[j]), "cudaDeviceCanAccessPeer");
      if (!can) throw Exception(Error::UnsupportedHardware, "CUDA devices of a multi-GPU device must be peer accessible");
      checkCuda(cudaSetDevice(deviceIDs[i]), "cudaSetDevice");
      const cudaError_t e = cudaDeviceEnablePeerAccess(deviceIDs[j], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) checkCuda(e, "cudaDeviceEnablePeerAccess");
      cudaGetLastError();
    }
  committed = true;
}

void Device::submitBarrier()
{
  if (engines.size() == 1) return; // a single stream is already ordered
  cudaStream_t s0 = static_cast<cudaStream_t>(engines[0]->getMainStream());
  for (size_t i = 1; i < engines.size(); ++i)
  {
    engines[i]->makeCurrent();
    checkCuda(cudaEventRecord(static_cast<cudaEvent_t>(events[i]), static_cast<cudaStream_t>(engines[i]->getMainStream())), "cudaEventRecord");
    checkCuda(cudaStreamWaitEvent(s0, static_cast<cudaEvent_t>(events[i]), 0), "cudaStreamWaitEvent");
  }
  engines[0]->makeCurrent();
  checkCuda(cudaEventRecord(static_cast<cudaEvent_t>(events[0]), s0), "cudaEventRecord");
  for (size_t i = 1; i < engines.size(); ++i)
    checkCuda(cudaStreamWaitEvent(static_cast<cudaStream_t>(engines[i]->getMainStream()), static_cast<cudaEvent_t>(events[0]), 0), "cudaStreamWaitEvent");
}

void Device::joinStaged()
{
  if (pendingJoin.empty()) return;
  engines[0]->makeCurrent();
  cudaStream_t s0 = static_cast<cudaStream_t>(engines[0]->getMainStream());
  for (void* ev : pendingJoin)
    checkCuda(cudaStreamWaitEvent(s0, static_cast<cudaEvent_t>(ev), 0), "cudaStreamWaitEvent");
  pendingJoin.clear();
  if (engines.size() > 1)
  {
    checkCuda(cudaEventRecord(static_cast<cudaEvent_t>(events[0]), s0), "cudaEventRecord");
    for (size_t i = 1; i < engines.size(); ++i)
    {
      engines[i]->makeCurrent();
      checkCuda(cudaStreamWaitEvent(static_cast<cudaStream_t>(engines[i]->getMainStream()), static_cast<cudaEvent_t>(events[0]), 0), "cudaStreamWaitEvent");
    }
  }
}

void Device::wait()
{
  for (auto& e : engines) e->wait();
  pendingJoin.clear();
}

std::shared_ptr<Filter> Device::newFilter(const std::string& type)
{
  checkCommitted();
  if (type == "RT") return std::make_shared<RTFilter>(this);
  if (type == "RTLightmap") return std::make_shared<RTLightmapFilter>(this);
  throw Exception(Error::InvalidArgument, "unknown filter type: '" + type + "'"); // core/device.cpp:283-298
}

void Device::setInt(const std::string& name, int value)
{
  if (name == "verbose") verbose = value;
  else if (name == "profile") profile = value; // backend specific: per-op CUDA-event timing
  else if (name == "maxTilePixels") maxTilePixels = value; // backend specific
  else if (name == "graph") graph = value;                 // backend specific: 1 = replay frames as a CUDA graph (frame streams)
  else if (name == "fuseOutput") fuseOutput = value;       // backend specific: 1 = output process inside the last conv's epilogue
  else if (name == "tilePolicy") tilePolicy = value;       // backend specific: 0 = reference search, 1 = fewest recomputed pixels
  else if (name == "fusePairs") fusePairs = value;         // backend specific: 1 = conv -> conv pairs as one launch (kernels/conv_pair_tc.cu)
  else if (name == "staging") staging = value;             // backend specific: tile staging, see UNetFilter::wantStaging
  else if (committed) throw Exception(Error::InvalidOperation, "device can be committed only once");
  else throw Exception(Error::InvalidArgument, "unknown device parameter or type mismatch: '" + name + "'");
}

int Device::getInt(const std::string& name) const
{
  if (name == "type") return 3; // OIDN_DEVICE_TYPE_CUDA
  if (name == "version") return 20401;
  if (name == "versionMajor") return 2;
  if (name == "versionMinor") return 4;
  if (name == "versionPatch") return 1;
  if (name == "verbose") return verbose;
  if (name == "profile") return profile;
  if (name == "numSubdevices") return (int)deviceIDs.size();
  if (name == "maxTilePixels") return (int)maxTilePixels;
  if (name == "tilePolicy") return tilePolicy;
  if (name == "graph") return graph;
  if (name == "fuseOutput") return fuseOutput;
  if (name == "staging") return staging;
  if (name == "fusePairs") return fusePairs;
  if (name == "systemMemorySupported" || name == "managedMemorySupported")
  {
    int v = 0;
    cudaDeviceGetAttribute(&v, name[0] == 's' ? cudaDevAttrPageableMemoryAccess : cudaDevAttrManagedMemory, deviceIDs[0]);
    cudaGetLastError();
    return v;
  }
  throw Exception(Error::InvalidArgument, "unknown device parameter or type mismatch: '" + name + "'");
}

void Device::setString(const std::string& name, const std::string& value)
{
  if (name == "weightsDir") weightsDir = value;
  else throw Exception(Error::InvalidArgument, "unknown device parameter or type mismatch: '" + name + "'");
}

// ---- host memory -----------------------------------------------------------------------------------
static long mbindInterleave(void* addr, size_t len)
{
#if defined(__linux__) && defined(SYS_mbind)
  // nodes that have memory: /sys/devices/system/node/has_memory, e.g. "0-1"
  unsigned long mask[16] = {0};
  int maxNode = -1;
  if (FILE* f = fopen("/sys/devices/system/node/has_memory", "r"))
  {
    char buf[256] = {0};
    if (fgets(buf, sizeof(buf), f))
      for (char* p = buf; *p;)
      {
        char* end;
        const long a = strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        if (*end == '-') { p = end + 1; b = strtol(p, &end, 10); }
        for (long n = a; n <= b && n < 1024; ++n) { mask[n / 64] |= 1ul << (n % 64); maxNode = std::max(maxNode, (int)n); }
        p = (*end == ',') ? end + 1 : end;
        if (*end != ',' ) break;
      }
    fclose(f);
  }
  if (maxNode < 1) return 0; // one node: nothing to interleave
  constexpr int MPOL_INTERLEAVE_ = 3;
  return syscall(SYS_mbind, addr, len, MPOL_INTERLEAVE_, mask, (unsigned long)(maxNode + 2), 0u);
#else
  (void)addr; (void)len;
  return 0;
#endif
}

void* Device::allocHost(size_t bytes)
{
  if (bytes == 0) return nullptr;
  engines.at(0)->makeCurrent();
  if (engines.size() > 1)
  {
    const size_t len = round_up(bytes, (size_t)(2u << 20));
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p != MAP_FAILED)
    {
      mbindInterleave(p, len);  // best effort: without it the pages land on the touching thread's node
      memset(p, 0, len);        // fault the pages in under the policy
      if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess)
      {
        hostMaps[p] = len;
        return p;
      }
      cudaGetLastError();
      munmap(p, len);
    }
  }
  void* p = nullptr;
  checkCuda(cudaHostAlloc(&p, bytes, cudaHostAllocPortable), "cudaHostAlloc");
  return p;
}

void Device::freeHost(void* ptr)
{
  if (!ptr) return;
  auto it = hostMaps.find(ptr);
  if (it != hostMaps.end())
  {
    cudaHostUnregister(it->first);
    munmap(it->first, it->second);
    hostMaps.erase(it);
    return;
  }
  cudaFreeHost(ptr);
}

Storage Device::getPtrStorage(const void* ptr) const
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess)
  {
    cudaGetLastError();
    return Storage::Undefined;
  }
  switch (a.type)
  {
  case cudaMemoryTypeHost:    return Storage::Host;
  case cudaMemoryTypeDevice:  return Storage::Device;
  case cudaMemoryTypeManaged: return Storage::Managed;
  default:                    return Storage::Undefined;
  }
}

void Device::setError(Error code, const std::string& message)
{
  if (errorCode == Error::None)
  {
    errorCode = code;
    errorMessage = message;
  }
  if (isVerbose(1)) fprintf(stderr, "Error: %s\n", message.c_str());
}

Error Device::getError(const char** outMessage)
{
  const Error c = errorCode;
  errorMessageOut = errorMessage;
  if (outMessage) *outMessage = c == Error::None ? nullptr : errorMessageOut.c_str();
  errorCode = Error::None;
  return c;
}

void Device::setGlobalError(Error code, const std::string& message)
{
  std::lock_guard<std::mutex> lock(g_globalMutex);
  if (g_globalCode == Error::None) { g_globalCode = code; g_globalMessage = message; }
}

Error Device::getGlobalError(const char** outMessage)
{
  std::lock_guard<std::mutex> lock(g_globalMutex);
  const Error c = g_globalCode;
  g_globalMessageOut = g_globalMessage;
  if (outMessage) *outMessage = c == Error::None ? nullptr : g_globalMessageOut.c_str();
  g_globalCode = Error::None;
  return c;
}

} // namespace oidnb200
