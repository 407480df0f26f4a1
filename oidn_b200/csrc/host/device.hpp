// Device: a set of engines (one per GPU/stream pair), the error slot, and device parameters.
// Follows core/device.h:50-173 / devices/cuda/cuda_device.cpp for the single-GPU contract and the
// multi-subdevice model of the reference's SYCL device (devices/sycl/sycl_device.cpp:286-320,
// :488-508) for several GPUs: tiles are dealt round-robin to the engines (core/unet_filter.cpp:219),
// every GPU dereferences the user's images directly through NVLink peer mappings, and
// submitBarrier() is an event join across the engines' streams.
#pragma once
#include "engine.hpp"
#include <map>
#include <mutex>
#include <vector>

namespace oidnb200 {

class Filter;

class Device
{
public:
  // ids/streams as in oidnNewCUDADevice(deviceIDs, streams, numPairs) (api/api.cpp:425-434);
  // streams[i] == nullptr -> the engine creates its own stream.
  Device(const std::vector<int>& deviceIDs, const std::vector<void*>& streams);
  ~Device();

  void commit();
  bool isCommitted() const { return committed; }
  void checkCommitted() const;

  int getNumEngines() const { return (int)engines.size(); }
  Engine* getEngine(int i = 0) const { return engines.at(i).get(); }

  // engine 0 waits for every engine, then every engine waits for engine 0
  void submitBarrier();
  void wait(); // block until every stream is idle
  // Staged frames (UNetFilter::submitFrameStaged) finish on the engines' copy-out streams: deferJoin() notes the
  // events that mark their end, joinStaged() makes every main stream wait for them. Called before anything else
  // is enqueued on a main stream (buffer copies, in-place frames) and by a caller-supplied stream's frames.
  void deferJoin(const std::vector<void*>& endEvents) { pendingJoin = endEvents; }
  void joinStaged();

  std::shared_ptr<Filter> newFilter(const std::string& type);

  // parameters (core/device.cpp:180-220 + backend specific ones)
  void setInt(const std::string& name, int value);
  int getInt(const std::string& name) const;
  void setString(const std::string& name, const std::string& value);

  bool isVerbose(int level) const { return verbose >= level; }
  long getMaxTilePixels() const { return maxTilePixels; }
  const std::string& getWeightsDir() const { return weightsDir; }
  int getMinTileAlignment() const { return 1; }

  // Classifies a user pointer (Device::getPtrStorage, devices/cuda/cuda_device.cpp:240-262)
  Storage getPtrStorage(const void* ptr) const;

  // error slot: first error sticks until read (core/device.cpp:97-161)
  void setError(Error code, const std::string& message);
  Error getError(const char** outMessage);
  static void setGlobalError(Error code, const std::string& message);
  static Error getGlobalError(const char** outMessage);

  std::mutex& getMutex() { return mutex; }

  // Pinned host memory for buffers with host storage. On a multi-GPU device the pages are interleaved over the
  // host's NUMA nodes (mbind) before they are page-locked, so the GPUs of both sockets pull their tiles of a
  // frame from local and remote memory alike instead of all crossing to the socket that first touched it.
  void* allocHost(size_t bytes);
  void freeHost(void* ptr);

private:
  std::vector<int> deviceIDs;
  std::vector<void*> userStreams;
  std::vector<std::unique_ptr<Engine>> engines;
  std::vector<void*> events; // one cudaEvent_t per engine
  std::vector<void*> pendingJoin;
  bool committed = false;
  int verbose = 0;
  int profile = 0;
  long maxTilePixels;
  int tilePolicy = 1;
  int fuseOutput = 1;
  int graph = 0;
  int staging = -1;
  int fusePairs = 1;
  std::map<void*, size_t> hostMaps; // interleaved host allocations (mmap + cudaHostRegister): ptr -> bytes
  std::string weightsDir;
  std::mutex mutex;
  Error errorCode = Error::None;
  std::string errorMessage, errorMessageOut;
};

} // namespace oidnb200
