// Filters: the parameter surface, model selection, tile/overlap scheduler and per-frame execute
// loop of the reference's UNet filters, rebuilt for the B200 engine.
//   Filter            core/filter.h:12-40, core/filter.cpp:23-86 (dirty tracking)
//   UNetFilter        core/unet_filter.{h,cpp} (params :43-113, commit :115-143, execute :145-252,
//                     tile planner :254-335, checkParams :346-392, model choice :394-466,
//                     topology :468-531, buildModel :534-653)
//   RTFilter          core/rt_filter.cpp:31-129
//   RTLightmapFilter  core/rtlightmap_filter.cpp:11-75
#pragma once
#include "device.hpp"
#include "graph.hpp"

namespace oidnb200 {

typedef bool (*ProgressMonitorFunction)(void* userPtr, double n);

struct Data
{
  const void* ptr = nullptr;
  size_t size = 0;
  explicit operator bool() const { return ptr != nullptr; }
};

class Filter
{
public:
  explicit Filter(Device* device) : device(device) {}
  virtual ~Filter() = default;

  virtual void setImage(const std::string& name, const Image& image) = 0;
  virtual void unsetImage(const std::string& name) = 0;
  virtual void setData(const std::string& name, const Data& data) = 0;
  virtual void updateData(const std::string& name) = 0;
  virtual void unsetData(const std::string& name) = 0;
  virtual void setInt(const std::string& name, int value) = 0;
  virtual int getInt(const std::string& name) = 0;
  virtual void setFloat(const std::string& name, float value) = 0;
  virtual float getFloat(const std::string& name) = 0;
  virtual void commit() = 0;
  virtual void execute(SyncMode sync) = 0;

  void setProgressMonitorFunction(ProgressMonitorFunction func, void* userPtr) { progressFunc = func; progressUserPtr = userPtr; }
  Device* getDevice() const { return device; }

protected:
  // set a parameter and track whether the model must be rebuilt
  void setParam(int& dst, int src) { dirtyParam |= dst != src; dst = src; }
  void setParam(bool& dst, int src) { dirtyParam |= dst != (src != 0); dst = src != 0; }
  void setParam(Quality& dst, Quality src) { dirtyParam |= dst != src; dst = src; }
  void setParam(Image& dst, const Image& src);
  void removeParam(Image& dst) { dirtyParam |= bool(dst); dst = Image(); }
  void setParam(Data& dst, const Data& src);
  void removeParam(Data& dst) { dirtyParam |= bool(dst); dst = Data(); }

  Device* device;
  ProgressMonitorFunction progressFunc = nullptr;
  void* progressUserPtr = nullptr;
  bool dirty = true;
  bool dirtyParam = true;
};

// Tile grid (also exposed to tests through the C API)
struct TilePlan
{
  int H = 0, W = 0;
  int tileH = 0, tileW = 0, tilePadH = 0, tilePadW = 0;
  int tileCountH = 1, tileCountW = 1;
  int tileAlignment = 16, tileOverlap = 0;
};

struct TileRect
{
  // input tile: source origin in the image, origin in the tile buffer, size (incl. overlaps)
  int hSrc, wSrc, hBuf, wBuf, H1, W1;
  // output tile: origin in the tile buffer, origin in the image, size
  int hOutBuf, wOutBuf, hDst, wDst, H2, W2;
};

// Shrinks tiles until numTiles % numEngines == 0, tile pixels <= maxTilePixels and fits(plan).
// Same search as core/unet_filter.cpp:283-326. `fits` is the memory test (buildModel).
TilePlan planTiles(int H, int W, bool largeModel, int deviceMinAlignment, int numEngines, long maxTilePixels,
                   const std::function<bool(const TilePlan&)>& fits);
// Own search (device parameter tilePolicy=1, the default): the grid with the fewest recomputed pixels.
// stripAware (tilePolicy=2): tile widths count in whole 128-pixel conv strips per UNet level.
TilePlan planTilesMinOverlap(int H, int W, bool largeModel, int deviceMinAlignment, int numUnits, long maxTilePixels,
                             const std::function<bool(const TilePlan&)>& fits, bool stripAware = false);
std::vector<TileRect> enumerateTiles(const TilePlan& plan);

class UNetFilter : public Filter
{
public:
  explicit UNetFilter(Device* device);
  ~UNetFilter() override;

  void setData(const std::string& name, const Data& data) override;
  void updateData(const std::string& name) override;
  void unsetData(const std::string& name) override;
  void setInt(const std::string& name, int value) override;
  int getInt(const std::string& name) override;
  void setFloat(const std::string& name, float value) override;
  float getFloat(const std::string& name) override;
  void commit() override;
  void execute(SyncMode sync) override;

  const TilePlan& getTilePlan() const { return plan; }
  bool isLargeModel() const { return largeModel; }
  size_t getScratchByteSize() const { return totalMemoryByteSize; }
  int getNumOps() const { return instances.empty() ? 0 : instances[0].graph->getNumOps(); }
  bool wasStaged() const { return lastStaged; }
  // per-op device times accumulated since the last reset (device param "profile" = 1)
  std::vector<Graph::OpTime> getProfile();
  void resetProfile() { profile.clear(); }

protected:
  virtual std::shared_ptr<TransferFunction> newTransferFunc() = 0;

  static constexpr int minTileAlignment = 16;       // core/unet_filter.h:35-39
  static constexpr int receptiveFieldBase = 174;
  static constexpr int receptiveFieldLarge = 202;
  static constexpr Quality defaultQuality = Quality::High;

  // Built-in model slots (core/unet_filter.h). A slot is a weights file name stem; the blob is
  // loaded from the device's weights directory on demand. The reference compiles these blobs in;
  // its weights submodule is Git-LFS pointers in this checkout, so they are looked up at run time.
  struct Model
  {
    const char* base = nullptr;
    const char* small = nullptr;
    const char* large = nullptr;
  };
  struct
  {
    Model hdr, hdr_alb, hdr_alb_nrm, hdr_calb_cnrm;
    Model ldr, ldr_alb, ldr_alb_nrm, ldr_calb_cnrm;
    Model dir, alb, nrm;
  } models;

  Image color, albedo, normal, output;
  Image outputTemp;
  bool hdr = false, srgb = false, directional = false, cleanAux = false;
  float inputScale;
  Quality quality = defaultQuality;
  int maxMemoryMB = -1;
  // Multi-process tile sharding (one process per GPU): this filter instance executes the tiles
  // whose index % numShards == shardIndex of the common plan, and reads the input scale from a
  // device pointer the caller keeps up to date (e.g. broadcast from the rank that ran the
  // autoexposure). Backend-specific parameters "numShards", "shardIndex", data "inputScalePtr".
  int numShards = 1, shardIndex = 0;
  const float* inputScaleDevPtr = nullptr;

private:
  void init();
  void cleanup();
  void checkParams();
  Data getWeights();
  Graph::Value addUNet(Graph& graph, Graph::Value input);
  Graph::Value addUNetLarge(Graph& graph, Graph::Value input);
  bool buildModel(const TilePlan& candidate, size_t maxMemoryByteSize, bool commitModel);
  void freeScratch();

  struct Instance
  {
    std::unique_ptr<Graph> graph;
    std::shared_ptr<InputProcess> inputProcess;
    std::shared_ptr<OutputProcess> outputProcess;
    std::shared_ptr<TransferFunction> transferFunc; // per engine: a staged frame gives every GPU its own copy of the input scale
    void* scratch = nullptr;
    size_t scratchByteSize = 0;
  };

  // Tile staging (device parameter "staging": -1 auto, 0 off, 1 on). When the frame does not live
  // in the memory of the GPU that runs a tile -- pinned host memory, or another GPU of a multi-GPU device --
  // the tile's source rectangles are brought into engine-local staging images by copy engines, the network
  // runs on local data, and the output rectangle goes back by a copy engine: three internal streams per engine
  // (copy-in, compute, copy-out) linked by events, two slot sets alternating between frames, so the copy-in of
  // frame f+1 and the copy-out of frame f-1 overlap the convolutions of frame f (what `oidnBenchmark --buffer
  // hostcopy` leaves to the application: apps/oidnBenchmark.cpp:165-180,343-359) and every GPU of a multi-GPU
  // device pulls its own tiles over its own PCIe link / NVLink port. Autoexposure on a staged frame: each tile
  // fills the bins it owns straight into engine 0's bin array (peer stores), engine 0 folds it in the fixed
  // order and hands every engine the scale -- bit-identical to the in-place pass. No NCCL, no host thread.
  struct Slot
  {
    void* in[3] = {nullptr, nullptr, nullptr};   // staged color / albedo / normal source rectangle (tile sized)
    void* out = nullptr;                          // staged output rectangle
    void* evIn = nullptr;                         // inputs have landed
    void* evDone = nullptr;                       // the network has consumed the inputs and written the output
    void* evOut = nullptr;                        // the output has left the slot
  };
  struct Staging
  {
    bool allocated = false;
    size_t inPitch[3] = {0, 0, 0}, outPitch = 0;
    std::vector<std::vector<int>> tilesOf;        // [engine] -> indices into `tiles`
    std::vector<std::vector<Slot>> slots;         // [engine][parity * tilesOf[engine].size() + j]
    std::vector<float*> scale;                    // [engine] -> 2 floats (one per parity)
    float* bins[2] = {nullptr, nullptr};          // engine 0: autoexposure bin array per parity
    int numBins = 0;
    std::vector<void*> evBins[2];                 // [parity][engine]: the engine's tiles have filled their bins
    void* evScale[2] = {nullptr, nullptr};        // engine 0 has distributed the scale
    std::vector<void*> evStart, evEnd;            // [engine]: frame start on the main stream / all copy-outs issued
    void* evJoin = nullptr;
    std::vector<void*> buffers;                   // every device allocation, with its engine
    std::vector<int> bufferEngine;
    uint64_t frame = 0;
  } staging;
  bool wantStaging() const;
  void ensureStaging();
  void freeStaging();
  struct ProgressState;
  void submitFrameStaged(const std::shared_ptr<ProgressState>& progress);
  void submitFrame(const std::shared_ptr<ProgressState>& progress);

  Data userWeightsBlob;
  std::shared_ptr<std::vector<uint8_t>> builtinBlob; // keeps a file-loaded model alive
  std::shared_ptr<TensorMap> constTensors;
  std::vector<Instance> instances;
  std::shared_ptr<TransferFunction> transferFunc;
  std::shared_ptr<Autoexposure> autoexposure;
  std::shared_ptr<ImageCopy> imageCopy;
  std::vector<TileRect> tiles;
  TilePlan plan;
  bool largeModel = false;
  bool lastStaged = false;
  bool inplace = false;
  int inplaceParam = 0;
  size_t totalMemoryByteSize = 0;
  std::vector<Graph::OpTime> profile;

  // Frame streams (device parameter "graph" = 1): the launch sequence of a whole frame
  // is captured into a CUDA graph the second time the filter runs with the same image pointers and
  // input scale, and replayed with one cudaGraphLaunch afterwards. Pointer / stride / scale changes
  // re-capture (they are kernel arguments); commit() and scratch reallocation drop the graph.
  // A few keys are remembered (least recently used goes first), so a caller that alternates between two or
  // three output buffers still replays graphs instead of re-capturing (or never capturing) every frame.
  struct FrameGraph
  {
    std::vector<uint64_t> key;
    int seen = 0;         // frames run with this key so far
    void* exec = nullptr; // cudaGraphExec_t
    uint64_t lastUse = 0;
  };
  static constexpr size_t maxFrameGraphs = 4;
  std::vector<FrameGraph> frameGraphs;
  uint64_t frameGraphClock = 0;
  std::vector<uint64_t> frameKey() const;
  void dropFrameGraph();
};

class RTFilter final : public UNetFilter
{
public:
  explicit RTFilter(Device* device);
  void setImage(const std::string& name, const Image& image) override;
  void unsetImage(const std::string& name) override;
  void setInt(const std::string& name, int value) override;
  int getInt(const std::string& name) override;
protected:
  std::shared_ptr<TransferFunction> newTransferFunc() override;
};

class RTLightmapFilter final : public UNetFilter
{
public:
  explicit RTLightmapFilter(Device* device);
  void setImage(const std::string& name, const Image& image) override;
  void unsetImage(const std::string& name) override;
  void setInt(const std::string& name, int value) override;
  int getInt(const std::string& name) override;
protected:
  std::shared_ptr<TransferFunction> newTransferFunc() override;
};

} // namespace oidnb200
