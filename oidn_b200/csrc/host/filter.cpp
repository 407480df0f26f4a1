#include "filter.hpp"
#include <cuda_runtime.h>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>

namespace oidnb200 {

// ------------------------------------------------------------------------------------------------
// Filter: dirty tracking (core/filter.cpp:23-86)
// ------------------------------------------------------------------------------------------------
void Filter::setParam(Image& dst, const Image& src)
{
  // The image must be dereferenceable by the GPU(s): device, managed or pinned host memory
  if (src && !device->getInt("systemMemorySupported") && device->getPtrStorage(src.ptr) == Storage::Undefined)
    throw Exception(Error::InvalidArgument,
                    "image data not accessible by the device, please use OIDNBuffer or device allocator for storage");
  // not dirty if only the pointer and/or strides change (except to/from nullptr)
  dirtyParam |= (!dst && src) || (dst && !src) ||
                (dst && src && (dst.W != src.W || dst.H != src.H || dst.format != src.format));
  dst = src ? src : Image();
}

void Filter::setParam(Data& dst, const Data& src)
{
  if (src && device->getPtrStorage(src.ptr) == Storage::Device)
    throw Exception(Error::InvalidArgument, "the specified data is not accessible to the host, please use host malloc");
  dirtyParam = bool(dst) || bool(src);
  dst = src;
}

// ------------------------------------------------------------------------------------------------
// Tile planner (core/unet_filter.cpp:254-335) and tile rectangles (:198-241)
// ------------------------------------------------------------------------------------------------
TilePlan planTiles(int H, int W, bool largeModel, int deviceMinAlignment, int numEngines, long maxTilePixels,
                   const std::function<bool(const TilePlan&)>& fits)
{
  constexpr int minAlign = 16;
  TilePlan p;
  p.H = H; p.W = W;
  const int receptiveField = largeModel ? 202 : 174;
  p.tileAlignment = lcm_(minAlign, std::max(deviceMinAlignment, 1));
  p.tileOverlap = round_up(receptiveField / 2, p.tileAlignment);
  p.tileH = round_up(H, minAlign);
  p.tileW = round_up(W, minAlign);
  p.tilePadH = p.tileH % p.tileAlignment;
  p.tilePadW = p.tileW % p.tileAlignment;
  p.tileCountH = p.tileCountW = 1;

  const int minTileDim = std::max(4 * p.tileOverlap, 768);
  const int minTileH = round_up(minTileDim, p.tileAlignment, p.tilePadH);
  const int minTileW = round_up(minTileDim, p.tileAlignment, p.tilePadW);
  const int ovH = 2 * p.tileOverlap + p.tilePadH, ovW = 2 * p.tileOverlap + p.tilePadW;

  while ((p.tileCountH * p.tileCountW) % numEngines != 0 || (long)p.tileH * p.tileW > maxTilePixels || !fits(p))
  {
    if (p.tileH > minTileH && p.tileH > p.tileW)
    {
      const int newTileH = ceil_div(H + ovH * p.tileCountH, p.tileCountH + 1);
      p.tileH = clamp_(round_up(newTileH, p.tileAlignment, p.tilePadH), minTileH, p.tileH - p.tileAlignment);
      p.tileCountH = std::max(ceil_div(H - ovH, p.tileH - ovH), 1);
    }
    else if (p.tileW > minTileW)
    {
      const int newTileW = ceil_div(W + ovW * p.tileCountW, p.tileCountW + 1);
      p.tileW = clamp_(round_up(newTileW, p.tileAlignment, p.tilePadW), minTileW, p.tileW - p.tileAlignment);
      p.tileCountW = std::max(ceil_div(W - ovW, p.tileW - ovW), 1);
    }
    else
      break; // cannot divide further; the caller builds the model without a memory limit
  }
  return p;
}

// Same geometry rules (overlap, alignment, padding, bottom/right-aligned last tiles), but instead of
// halving the longer side until the plan fits, every grid countH x countW with a tile count
// divisible by numUnits is costed and the one that recomputes the fewest pixels wins. On a B200
// memory rarely forces tiling, the number of GPUs/engines does: 8K on 2 GPUs becomes 1x2 tiles
// (+4 % pixels) where the halving search gives 3x2 (+10 %).
// stripAware (tilePolicy=2, opt-in until measured at 8 GPUs): the conv kernel works in strips of 128
// pixels at every UNet level, so a 2064-wide tile costs 17 strips at full resolution (and 9, 5, 3, 2
// at the pooled levels), not 16.1: the width enters the cost rounded up to whole strips per level,
// weighted by the level's share of the FLOPs.
static double effectiveWidth(int tileW)
{
  static const double share[5] = {0.583, 0.248, 0.124, 0.040, 0.005}; // base UNet, SURVEY.md section 8(d) table by level
  double w = 0;
  for (int l = 0; l < 5; ++l)
  {
    const int wl = tileW >> l;
    w += share[l] * (double)(ceil_div(wl, 128) * 128) * (double)(1 << l);
  }
  return w;
}

TilePlan planTilesMinOverlap(int H, int W, bool largeModel, int deviceMinAlignment, int numUnits, long maxTilePixels,
                             const std::function<bool(const TilePlan&)>& fits, bool stripAware)
{
  constexpr int minAlign = 16;
  TilePlan base;
  base.H = H; base.W = W;
  const int receptiveField = largeModel ? 202 : 174;
  base.tileAlignment = lcm_(minAlign, std::max(deviceMinAlignment, 1));
  base.tileOverlap = round_up(receptiveField / 2, base.tileAlignment);
  base.tileH = round_up(H, minAlign);
  base.tileW = round_up(W, minAlign);
  base.tilePadH = base.tileH % base.tileAlignment;
  base.tilePadW = base.tileW % base.tileAlignment;
  base.tileCountH = base.tileCountW = 1;
  const int minTileDim = std::max(4 * base.tileOverlap, 768);
  const int minTileH = round_up(minTileDim, base.tileAlignment, base.tilePadH);
  const int minTileW = round_up(minTileDim, base.tileAlignment, base.tilePadW);
  const int ovH = 2 * base.tileOverlap + base.tilePadH, ovW = 2 * base.tileOverlap + base.tilePadW;

  // tile size along one axis for a target count, and the count that size really needs
  auto axis = [&](int L, int full, int ov, int pad, int minTile, int count, int& tile, int& realCount) {
    if (count == 1) { tile = full; realCount = 1; return true; }
    tile = round_up(ceil_div(L + ov * (count - 1), count), base.tileAlignment, pad);
    tile = std::max(tile, minTile);
    if (tile >= full) return false;
    realCount = std::max(ceil_div(L - ov, tile - ov), 1);
    return realCount == count;
  };

  TilePlan best; bool found = false; double bestCost = 0;
  const int maxCount = 64;
  for (int ch = 1; ch <= maxCount; ++ch)
    for (int cw = 1; cw <= maxCount; ++cw)
    {
      if ((ch * cw) % numUnits != 0) continue;
      TilePlan p = base;
      int rh, rw;
      if (!axis(H, base.tileH, ovH, base.tilePadH, minTileH, ch, p.tileH, rh)) continue;
      if (!axis(W, base.tileW, ovW, base.tilePadW, minTileW, cw, p.tileW, rw)) continue;
      p.tileCountH = ch; p.tileCountW = cw;
      if ((long)p.tileH * p.tileW > maxTilePixels) continue;
      // cost: pixels pushed through the network per unit (tiles are dealt round-robin), then fewer tiles
      const double cost = (double)(ch * cw / numUnits) * p.tileH * (stripAware ? effectiveWidth(p.tileW) : (double)p.tileW) +
                          1e-3 * ch * cw;
      if (found && cost >= bestCost) continue;
      if (!fits(p)) continue;
      best = p; bestCost = cost; found = true;
    }
  if (found) return best;
  return planTiles(H, W, largeModel, deviceMinAlignment, numUnits, maxTilePixels, fits); // nothing fits: reference search
}

std::vector<TileRect> enumerateTiles(const TilePlan& p)
{
  std::vector<TileRect> out;
  const int ovH = 2 * p.tileOverlap + p.tilePadH, ovW = 2 * p.tileOverlap + p.tilePadW;
  for (int i = 0; i < p.tileCountH; ++i)
  {
    const int h = i * (p.tileH - ovH);
    const int obH = i > 0 ? p.tileOverlap : 0;
    const int oeH = i < p.tileCountH - 1 ? p.tileOverlap + p.tilePadH : 0;
    const int H1 = std::min(p.H - h, p.tileH), H2 = H1 - obH - oeH;
    const int alignH = p.tileH - round_up(H1, 16); // bottom-aligned: keeps the 16-px pooling grid
    for (int j = 0; j < p.tileCountW; ++j)
    {
      const int w = j * (p.tileW - ovW);
      const int obW = j > 0 ? p.tileOverlap : 0;
      const int oeW = j < p.tileCountW - 1 ? p.tileOverlap + p.tilePadW : 0;
      const int W1 = std::min(p.W - w, p.tileW), W2 = W1 - obW - oeW;
      const int alignW = p.tileW - round_up(W1, 16);
      out.push_back(TileRect{h, w, alignH, alignW, H1, W1, alignH + obH, alignW + obW, h + obH, w + obW, H2, W2});
    }
  }
  return out;
}

// ------------------------------------------------------------------------------------------------
// UNetFilter
// ------------------------------------------------------------------------------------------------
UNetFilter::UNetFilter(Device* device) : Filter(device), inputScale(std::numeric_limits<float>::quiet_NaN()) {}

UNetFilter::~UNetFilter()
{
  try { device->wait(); } catch (...) {}
  cleanup();
}

static void warnUnknown(Device* device, const std::string& name)
{
  if (device->isVerbose(1))
    std::cerr << "Warning: unknown filter parameter or type mismatch: '" << name << "'" << std::endl;
}

void UNetFilter::setData(const std::string& name, const Data& data)
{
  if (name == "weights") setParam(userWeightsBlob, data);
  else if (name == "inputScalePtr")
  {
    if (data && data.size < sizeof(float)) throw Exception(Error::InvalidArgument, "inputScalePtr must point to one float");
    inputScaleDevPtr = static_cast<const float*>(data.ptr); // device memory, dereferenced by the kernels only
  }
  else warnUnknown(device, name);
  dirty = true;
}

void UNetFilter::updateData(const std::string& name)
{
  if (name == "weights") dirtyParam |= bool(userWeightsBlob); else warnUnknown(device, name);
  dirty = true;
}

void UNetFilter::unsetData(const std::string& name)
{
  if (name == "weights") removeParam(userWeightsBlob);
  else if (name == "inputScalePtr") inputScaleDevPtr = nullptr;
  else warnUnknown(device, name);
  dirty = true;
}

void UNetFilter::setInt(const std::string& name, int value)
{
  if (name == "quality")
  {
    Quality q = static_cast<Quality>(value);
    if (q == Quality::Default) q = defaultQuality;
    else if (q != Quality::High && q != Quality::Balanced && q != Quality::Fast)
      throw Exception(Error::InvalidArgument, "unknown filter quality mode");
    setParam(quality, q);
  }
  else if (name == "maxMemoryMB") setParam(maxMemoryMB, value);
  else if (name == "numShards")
  {
    if (value < 1) throw Exception(Error::InvalidArgument, "numShards must be >= 1");
    setParam(numShards, value);
  }
  else if (name == "shardIndex")
  {
    if (value < 0) throw Exception(Error::InvalidArgument, "shardIndex must be >= 0");
    setParam(shardIndex, value);
  }
  else warnUnknown(device, name);
  dirty = true;
}

int UNetFilter::getInt(const std::string& name)
{
  if (name == "quality") return static_cast<int>(quality);
  if (name == "maxMemoryMB") return maxMemoryMB;
  if (name == "numShards") return numShards;
  if (name == "shardIndex") return shardIndex;
  if (name == "tileAlignment" || name == "alignment") return plan.tileAlignment;
  if (name == "tileOverlap" || name == "overlap") return plan.tileOverlap;
  throw Exception(Error::InvalidArgument, "unknown filter parameter or type mismatch: '" + name + "'");
}

void UNetFilter::setFloat(const std::string& name, float value)
{
  if (name == "inputScale" || name == "hdrScale") inputScale = value; else warnUnknown(device, name);
  dirty = true;
}

float UNetFilter::getFloat(const std::string& name)
{
  if (name == "inputScale" || name == "hdrScale") return inputScale;
  throw Exception(Error::InvalidArgument, "unknown filter parameter or type mismatch: '" + name + "'");
}

void UNetFilter::commit()
{
  if (!dirty) return;
  const bool inplaceNew = output && ((color && output.overlaps(color)) || (albedo && output.overlaps(albedo)) ||
                                     (normal && output.overlaps(normal)));
  setParam(inplaceParam, (int)inplaceNew);
  inplace = inplaceNew;
  if (dirtyParam)
  {
    device->wait(); // all asynchronous work must have completed before the model is rebuilt
    init();
  }
  dirty = false;
  dirtyParam = false;
}

void UNetFilter::cleanup()
{
  freeScratch();
  instances.clear();
  transferFunc.reset();
  autoexposure.reset();
  imageCopy.reset();
  outputTemp = Image();
  tiles.clear();
}

void UNetFilter::dropFrameGraph()
{
  for (FrameGraph& g : frameGraphs)
    if (g.exec)
    {
      device->getEngine(0)->makeCurrent();
      cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(g.exec));
    }
  frameGraphs.clear();
}

// Everything a frame's kernel arguments depend on besides the committed model.
std::vector<uint64_t> UNetFilter::frameKey() const
{
  std::vector<uint64_t> k;
  for (const Image* im : {&color, &albedo, &normal, &output})
  {
    k.push_back((uint64_t)(uintptr_t)im->ptr);
    k.push_back(((uint64_t)(uint32_t)im->W << 32) | (uint32_t)im->H);
    k.push_back((uint64_t)im->format);
    k.push_back((uint64_t)im->pixelStride);
    k.push_back((uint64_t)im->rowStride);
  }
  uint32_t bits;
  memcpy(&bits, &inputScale, sizeof(bits));
  k.push_back(bits);
  k.push_back((uint64_t)(uintptr_t)inputScaleDevPtr);
  k.push_back(((uint64_t)(uint32_t)numShards << 32) | (uint32_t)shardIndex);
  k.push_back((uint64_t)device->getInt("fuseOutput") | ((uint64_t)device->getInt("fusePairs") << 1));
  return k;
}

void UNetFilter::freeScratch()
{
  dropFrameGraph();
  freeStaging();
  for (size_t i = 0; i < instances.size(); ++i)
    if (instances[i].scratch)
    {
      device->getEngine((int)i)->free(instances[i].scratch);
      instances[i].scratch = nullptr;
    }
}

void UNetFilter::checkParams()
{
  if (!color && !albedo && !normal) throw Exception(Error::InvalidOperation, "input image not specified");
  if (!output) throw Exception(Error::InvalidOperation, "output image not specified");
  auto supported = [](Format f) {
    return f == Format::Float3 || f == Format::Half3 || f == Format::Float2 || f == Format::Half2 ||
           f == Format::Float || f == Format::Half;
  };
  if ((color && !supported(color.format)) || (albedo && !supported(albedo.format)) || (normal && !supported(normal.format)))
    throw Exception(Error::InvalidOperation, "unsupported input image format");
  if (!supported(output.format)) throw Exception(Error::InvalidOperation, "unsupported output image format");
  const Image& input = color ? color : (albedo ? albedo : normal);
  if (input.C() != output.C()) throw Exception(Error::InvalidOperation, "input/output image channel count mismatch");
  if ((color && (color.W != output.W || color.H != output.H)) || (albedo && (albedo.W != output.W || albedo.H != output.H)) ||
      (normal && (normal.W != output.W || normal.H != output.H)))
    throw Exception(Error::InvalidOperation, "image size mismatch");
  if (directional && (hdr || srgb))
    throw Exception(Error::InvalidOperation, "directional and hdr/srgb modes cannot be enabled at the same time");
  if (hdr && srgb) throw Exception(Error::InvalidOperation, "hdr and srgb modes cannot be enabled at the same time");
  if (shardIndex >= numShards) throw Exception(Error::InvalidOperation, "shardIndex must be smaller than numShards");
}

// Reads <weightsDir>/<stem>.tza; a missing file or a Git-LFS pointer counts as "model not available".
static std::shared_ptr<std::vector<uint8_t>> loadModelFile(const std::string& dir, const char* stem)
{
  if (!stem || dir.empty()) return nullptr;
  std::ifstream f(dir + "/" + stem + ".tza", std::ios::binary);
  if (!f) return nullptr;
  auto blob = std::make_shared<std::vector<uint8_t>>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (blob->size() < 12 || memcmp(blob->data(), "version ", 8) == 0) return nullptr;
  return blob;
}

Data UNetFilter::getWeights()
{
  // model slot: core/unet_filter.cpp:394-441
  Model* model = nullptr;
  if (color)
  {
    if (!albedo && !normal) model = directional ? &models.dir : (hdr ? &models.hdr : &models.ldr);
    else if (albedo && !normal) model = hdr ? &models.hdr_alb : &models.ldr_alb;
    else if (albedo && normal)
    {
      if (cleanAux) model = hdr ? &models.hdr_calb_cnrm : &models.ldr_calb_cnrm;
      else model = hdr ? &models.hdr_alb_nrm : &models.ldr_alb_nrm;
    }
  }
  else
  {
    if (albedo && !normal)
    {
      if (hdr) throw Exception(Error::InvalidOperation, "hdr mode is not supported for albedo filtering");
      model = &models.alb;
    }
    else if (!albedo && normal)
    {
      if (hdr || srgb) throw Exception(Error::InvalidOperation, "hdr and srgb modes are not supported for normal filtering");
      model = &models.nrm;
    }
    else
      throw Exception(Error::InvalidOperation, "invalid combination of input features");
  }

  if (userWeightsBlob) return userWeightsBlob;

  // quality -> variant with the reference's fallbacks (core/unet_filter.cpp:451-463)
  builtinBlob.reset();
  if (model)
  {
    const std::string& dir = device->getWeightsDir();
    auto pick = [&](const char* first, const char* second) {
      auto b = loadModelFile(dir, first);
      return b ? b : loadModelFile(dir, second);
    };
    switch (quality)
    {
    case Quality::Default:
    case Quality::High:     builtinBlob = pick(model->large, model->base); break;
    case Quality::Balanced: builtinBlob = loadModelFile(dir, model->base); break;
    case Quality::Fast:     builtinBlob = pick(model->small, model->base); break;
    }
  }
  if (!builtinBlob) throw Exception(Error::InvalidOperation, "unsupported combination of input features");
  Data d;
  d.ptr = builtinBlob->data();
  d.size = builtinBlob->size();
  return d;
}

Graph::Value UNetFilter::addUNet(Graph& g, Graph::Value input)
{
  // core/unet_filter.cpp:468-498 == training/model.py:55-155
  auto x = g.addConv("enc_conv0", input, Activation::ReLU);
  auto pool1 = x = g.addConv("enc_conv1", x, Activation::ReLU, PostOp::Pool);
  auto pool2 = x = g.addConv("enc_conv2", x, Activation::ReLU, PostOp::Pool);
  auto pool3 = x = g.addConv("enc_conv3", x, Activation::ReLU, PostOp::Pool);
  auto pool4 = x = g.addConv("enc_conv4", x, Activation::ReLU, PostOp::Pool);
  x = g.addConv("enc_conv5a", pool4, Activation::ReLU);
  x = g.addConv("enc_conv5b", x, Activation::ReLU, PostOp::Upsample);
  x = g.addConcatConv("dec_conv4a", x, pool3, Activation::ReLU);
  x = g.addConv("dec_conv4b", x, Activation::ReLU, PostOp::Upsample);
  x = g.addConcatConv("dec_conv3a", x, pool2, Activation::ReLU);
  x = g.addConv("dec_conv3b", x, Activation::ReLU, PostOp::Upsample);
  x = g.addConcatConv("dec_conv2a", x, pool1, Activation::ReLU);
  x = g.addConv("dec_conv2b", x, Activation::ReLU, PostOp::Upsample);
  x = g.addConcatConv("dec_conv1a", x, input, Activation::ReLU);
  x = g.addConv("dec_conv1b", x, Activation::ReLU);
  x = g.addConv("dec_conv0", x, Activation::ReLU);
  return x;
}

Graph::Value UNetFilter::addUNetLarge(Graph& g, Graph::Value input)
{
  // core/unet_filter.cpp:500-531 == training/model.py:161-260
  auto x = g.addConv("enc_conv1a", input, Activation::ReLU);
  auto pool1 = x = g.addConv("enc_conv1b", x, Activation::ReLU, PostOp::Pool);
  x = g.addConv("enc_conv2a", x, Activation::ReLU);
  auto pool2 = x = g.addConv("enc_conv2b", x, Activation::ReLU, PostOp::Pool);
  x = g.addConv("enc_conv3a", x, Activation::ReLU);
  auto pool3 = x = g.addConv("enc_conv3b", x, Activation::ReLU, PostOp::Pool);
  x = g.addConv("enc_conv4a", x, Activation::ReLU);
  auto pool4 = x = g.addConv("enc_conv4b", x, Activation::ReLU, PostOp::Pool);
  x = g.addConv("enc_conv5a", pool4, Activation::ReLU);
  x = g.addConv("enc_conv5b", x, Activation::ReLU, PostOp::Upsample);
  x = g.addConcatConv("dec_conv4a", x, pool3, Activation::ReLU);
  x = g.addConv("dec_conv4b", x, Activation::ReLU, PostOp::Upsample);
  x = g.addConcatConv("dec_conv3a", x, pool2, Activation::ReLU);
  x = g.addConv("dec_conv3b", x, Activation::ReLU, PostOp::Upsample);
  x = g.addConcatConv("dec_conv2a", x, pool1, Activation::ReLU);
  x = g.addConv("dec_conv2b", x, Activation::ReLU, PostOp::Upsample);
  x = g.addConcatConv("dec_conv1a", x, input, Activation::ReLU);
  x = g.addConv("dec_conv1b", x, Activation::ReLU);
  x = g.addConv("dec_conv1c", x, Activation::ReLU);
  return x;
}

// Builds one graph per engine for the candidate tile size. Returns false if the memory estimate
// exceeds maxMemoryByteSize (core/unet_filter.cpp:534-653). With commitModel the scratch buffers
// are allocated and the graphs finalized.
bool UNetFilter::buildModel(const TilePlan& cand, size_t maxMemoryByteSize, bool commitModel)
{
  freeScratch();
  instances.clear();
  autoexposure.reset();
  imageCopy.reset();
  outputTemp = Image();
  if (cand.H <= 0 || cand.W <= 0) return true;

  int inputC = 0;
  if (color) inputC += 3; // always broadcast to 3 channels
  if (albedo) inputC += 3;
  if (normal) inputC += 3;
  const bool snorm = directional || (!color && normal);
  const int numEngines = device->getNumEngines();

  if (hdr) autoexposure = device->getEngine(0)->newAutoexposure(color.H, color.W);

  size_t outputTempOffset = SIZE_MAX, autoexposureDstOffset = SIZE_MAX;
  for (int id = 0; id < numEngines; ++id)
  {
    Engine* engine = device->getEngine(id);
    instances.emplace_back();
    Instance& inst = instances.back();
    inst.graph.reset(new Graph(engine, constTensors));
    Graph& g = *inst.graph;
    inst.transferFunc = newTransferFunc();
    auto in = g.addInputProcess("input", TensorDesc{inputC, cand.tileH, cand.tileW}, inst.transferFunc, hdr, snorm);
    auto x = largeModel ? addUNetLarge(g, in) : addUNet(g, in);
    g.addOutputProcess("output", x, inst.transferFunc, hdr, snorm);

    const size_t graphScratch = round_up(g.getScratchByteSize(), memoryAlignment);
    size_t scratch = graphScratch;
    if (id == 0 && hdr) scratch = round_up(std::max(scratch, autoexposure->getScratchByteSize()), memoryAlignment);
    if (id == 0 && inplace && cand.tileCountH * cand.tileCountW > 1)
    {
      outputTempOffset = scratch;
      scratch += round_up((size_t)output.W * output.H * formatBytes(output.format), memoryAlignment);
    }
    if (id == 0 && hdr)
    {
      autoexposureDstOffset = scratch;
      scratch += round_up(sizeof(float), memoryAlignment);
    }
    inst.scratchByteSize = scratch;
    if (id == 0)
    {
      totalMemoryByteSize = (scratch + g.getPrivateByteSize()) + (graphScratch + g.getPrivateByteSize()) * (size_t)(numEngines - 1);
      if (totalMemoryByteSize > maxMemoryByteSize)
      {
        instances.clear();
        return false;
      }
    }
  }
  if (!commitModel) return true;

  for (int id = 0; id < numEngines; ++id)
  {
    Instance& inst = instances[id];
    inst.scratch = device->getEngine(id)->malloc(inst.scratchByteSize);
    inst.graph->setScratch(inst.scratch, inst.scratchByteSize);
    inst.graph->finalize();
    inst.inputProcess = inst.graph->getInputProcess();
    inst.outputProcess = inst.graph->getOutputProcess();
  }
  uint8_t* s0 = static_cast<uint8_t*>(instances[0].scratch);
  if (hdr)
  {
    autoexposure->setScratch(s0);
    autoexposure->setDst(reinterpret_cast<float*>(s0 + autoexposureDstOffset));
  }
  if (outputTempOffset != SIZE_MAX)
  {
    outputTemp = output;
    outputTemp.ptr = s0 + outputTempOffset;
    outputTemp.pixelStride = formatBytes(output.format);
    outputTemp.rowStride = outputTemp.pixelStride * output.W;
    imageCopy = device->getEngine(0)->newImageCopy();
    imageCopy->setSrc(outputTemp);
  }
  return true;
}

void UNetFilter::init()
{
  cleanup();
  checkParams();

  const Data weights = getWeights();
  constTensors = parseTZA(weights.ptr, weights.size);
  largeModel = constTensors->find("enc_conv1b.weight") != constTensors->end();
  transferFunc = newTransferFunc();

  const int H = output.H, W = output.W;
  const long maxTilePixels = maxMemoryMB < 0 ? device->getMaxTilePixels() : LONG_MAX;
  const size_t maxMemoryByteSize = maxMemoryMB >= 0 ? (size_t)maxMemoryMB * 1024 * 1024 : SIZE_MAX;

  const auto fits = [&](const TilePlan& c) { return buildModel(c, maxMemoryByteSize, false); };
  const int units = device->getNumEngines() * numShards;
  const int policy = device->getInt("tilePolicy");
  plan = policy == 0
           ? planTiles(H, W, largeModel, device->getMinTileAlignment(), units, maxTilePixels, fits)
           : planTilesMinOverlap(H, W, largeModel, device->getMinTileAlignment(), units, maxTilePixels, fits, policy == 2);
  if (!buildModel(plan, SIZE_MAX, true)) throw std::runtime_error("could not build filter model");
  tiles = enumerateTiles(plan);

  if (device->isVerbose(2))
  {
    std::cout << "Image size: " << W << "x" << H << std::endl;
    std::cout << "Tile size : " << plan.tileW << "x" << plan.tileH << std::endl;
    std::cout << "Tile count: " << plan.tileCountW << "x" << plan.tileCountH << std::endl;
    std::cout << "In-place  : " << (inplace ? "true" : "false") << std::endl;
    std::cout << "Memory usage: " << totalMemoryByteSize << std::endl;
  }
}

// Progress of one execute(): every op advances it by one unit (core/op.cpp:8-22, core/progress.cpp:17-39),
// from host callbacks in stream order; engines of a multi-GPU device report concurrently.
struct UNetFilter::ProgressState
{
  ProgressMonitorFunction func;
  void* userPtr;
  double total;
  double done = 0;
  std::mutex mutex;
  std::atomic<bool> cancelled{false};
  void update(double amount)
  {
    std::lock_guard<std::mutex> lock(mutex);
    if (cancelled) return;
    done += amount;
    if (!func(userPtr, std::min(done / total, 1.0))) cancelled = true;
  }
};

// ------------------------------------------------------------------------------------------------
// Tile staging
// ------------------------------------------------------------------------------------------------
static bool packedPixels(const Image& im) { return !im || im.pixelStride == formatBytes(im.format); }

bool UNetFilter::wantStaging() const
{
  const int mode = device->getInt("staging");
  if (mode == 0 || numShards > 1) return false;  // sharded processes stage their own tiles (oidn_b200/sharded.py)
  // copy engines move whole row segments: every image must have packed pixels
  if (!packedPixels(color) || !packedPixels(albedo) || !packedPixels(normal) || !packedPixels(output)) return false;
  if (mode > 0) return true;
  // auto: the frame is not in the memory of (all of) the GPUs that run its tiles
  const bool multi = device->getNumEngines() > 1;
  for (const Image* im : {&color, &albedo, &normal, &output})
  {
    if (!*im) continue;
    const Storage st = device->getPtrStorage(im->ptr);
    if (st == Storage::Host || (multi && st == Storage::Device)) return true;
  }
  return false;
}

void UNetFilter::freeStaging()
{
  for (size_t i = 0; i < staging.buffers.size(); ++i)
    device->getEngine(staging.bufferEngine[i])->free(staging.buffers[i]);
  staging = Staging(); // events belong to the engines
}

void UNetFilter::ensureStaging()
{
  if (staging.allocated) return;
  const int E = device->getNumEngines();
  const Image* ins[3] = {&color, &albedo, &normal};
  auto pitchOf = [&](const Image& im) { return im ? round_up((size_t)plan.tileW * im.pixelStride, memoryAlignment) : (size_t)0; };
  for (int k = 0; k < 3; ++k) staging.inPitch[k] = pitchOf(*ins[k]);
  staging.outPitch = pitchOf(output);
  staging.tilesOf.assign(E, {});
  for (size_t i = 0; i < tiles.size(); ++i) staging.tilesOf[i % E].push_back((int)i);
  auto alloc = [&](int e, size_t bytes) {
    void* p = device->getEngine(e)->malloc(bytes);
    staging.buffers.push_back(p); staging.bufferEngine.push_back(e);
    return p;
  };
  staging.slots.assign(E, {});
  staging.scale.assign(E, nullptr);
  staging.evStart.assign(E, nullptr); staging.evEnd.assign(E, nullptr);
  for (int par = 0; par < 2; ++par) staging.evBins[par].assign(E, nullptr);
  for (int e = 0; e < E; ++e)
  {
    Engine* eng = device->getEngine(e);
    staging.slots[e].resize(2 * staging.tilesOf[e].size());
    for (Slot& sl : staging.slots[e])
    {
      for (int k = 0; k < 3; ++k)
        if (*ins[k]) sl.in[k] = alloc(e, staging.inPitch[k] * plan.tileH);
      sl.out = alloc(e, staging.outPitch * plan.tileH);
      sl.evIn = eng->newEvent(); sl.evDone = eng->newEvent(); sl.evOut = eng->newEvent();
    }
    staging.scale[e] = static_cast<float*>(alloc(e, memoryAlignment));
    staging.evStart[e] = eng->newEvent(); staging.evEnd[e] = eng->newEvent();
    for (int par = 0; par < 2; ++par) staging.evBins[par][e] = eng->newEvent();
  }
  if (hdr && color)
  {
    int nbh = 0, nbw = 0;
    oidnb200_autoexposure_bin_grid(color.H, color.W, &nbh, &nbw);
    staging.numBins = nbh * nbw;
    for (int par = 0; par < 2; ++par)
    {
      staging.bins[par] = static_cast<float*>(alloc(0, (size_t)staging.numBins * sizeof(float)));
      staging.evScale[par] = device->getEngine(0)->newEvent();
    }
  }
  staging.evJoin = device->getEngine(0)->newEvent();
  staging.allocated = true;
}

// Smallest bin index i in [0, n] whose first pixel i*size/n is >= x (bins of core/autoexposure.h:20-24).
static int firstBinAtOrAfter(int x, int n, int size)
{
  int i = (int)std::min<long>(n, ((long)x * n + size - 1) / size);
  while (i > 0 && (long)(i - 1) * size / n >= x) --i;
  while (i < n && (long)i * size / n < x) ++i;
  return i;
}

static inline cudaStream_t cs(void* s) { return static_cast<cudaStream_t>(s); }
static inline cudaEvent_t ce(void* e) { return static_cast<cudaEvent_t>(e); }

void UNetFilter::submitFrameStaged(const std::shared_ptr<ProgressState>& progress)
{
  ensureStaging();
  Staging& S = staging;
  const int E = device->getNumEngines();
  const int par = (int)(S.frame++ & 1);
  const bool autoScale = hdr && color && std::isnan(inputScale) && !inputScaleDevPtr;
  const int profiling = device->getInt("profile");
  const Image* ins[3] = {&color, &albedo, &normal};
  // The copy-in is ordered after everything enqueued on the main stream(s) before this call (buffer writes, the
  // caller's own work). The frame's join back into the main stream is deferred on streams the engines own (see
  // the end of this function), so that order does not drag the previous frame's copy-out along.

  auto report = [&](Engine* e) {
    if (progress) { auto p = progress; e->submitHostFunc([p]() { p->update(1.); }); }
  };
  auto virtualImage = [](const Image& user, void* buf, size_t pitch, int h0, int w0) {
    Image v = user; // full-frame coordinates; only the staged rectangle is ever touched
    v.ptr = static_cast<uint8_t*>(buf) - ((size_t)h0 * pitch + (size_t)w0 * user.pixelStride);
    v.rowStride = pitch;
    return v;
  };

  // ---- copy-in (+ the autoexposure bins of the tiles as they land)
  int nbh = 0, nbw = 0;
  if (autoScale) oidnb200_autoexposure_bin_grid(color.H, color.W, &nbh, &nbw);
  for (int e = 0; e < E; ++e)
  {
    Engine* eng = device->getEngine(e);
    eng->makeCurrent();
    cudaStream_t in = cs(eng->getAuxStream(Engine::CopyIn));
    checkCuda(cudaEventRecord(ce(S.evStart[e]), cs(eng->getMainStream())), "cudaEventRecord");
    checkCuda(cudaStreamWaitEvent(in, ce(S.evStart[e]), 0), "cudaStreamWaitEvent");
    const size_t T = S.tilesOf[e].size();
    for (size_t j = 0; j < T; ++j)
    {
      const TileRect& t = tiles[S.tilesOf[e][j]];
      Slot& sl = S.slots[e][par * T + j];
      checkCuda(cudaStreamWaitEvent(in, ce(sl.evDone), 0), "cudaStreamWaitEvent"); // frame f-2 has consumed the slot
      for (int k = 0; k < 3; ++k)
      {
        const Image& im = *ins[k];
        if (!im) continue;
        const uint8_t* src = static_cast<const uint8_t*>(im.ptr) + (size_t)t.hSrc * im.rowStride + (size_t)t.wSrc * im.pixelStride;
        checkCuda(cudaMemcpy2DAsync(sl.in[k], S.inPitch[k], src, im.rowStride, (size_t)t.W1 * im.pixelStride, t.H1,
                                    cudaMemcpyDefault, in), "cudaMemcpy2DAsync (tile copy-in)");
      }
      checkCuda(cudaEventRecord(ce(sl.evIn), in), "cudaEventRecord");
    }
  }

  // ---- autoexposure bins of the tiles as they land (compute streams: the copy-in stream carries copies only)
  if (autoScale)
    for (int e = 0; e < E; ++e)
    {
      Engine* eng = device->getEngine(e);
      eng->makeCurrent();
      cudaStream_t comp = cs(eng->getAuxStream(Engine::Compute));
      const size_t T = S.tilesOf[e].size();
      for (size_t j = 0; j < T; ++j)
      {
        const TileRect& t = tiles[S.tilesOf[e][j]];
        Slot& sl = S.slots[e][par * T + j];
        checkCuda(cudaStreamWaitEvent(comp, ce(sl.evIn), 0), "cudaStreamWaitEvent");
        // bins whose first pixel lies in the tile's destination rectangle: these rectangles partition the bin
        // grid, and a bin (<= 16 px) never leaves the source rectangle (overlap >= 96 px at interior edges)
        const int bh0 = firstBinAtOrAfter(t.hDst, nbh, color.H), bh1 = firstBinAtOrAfter(t.hDst + t.H2, nbh, color.H);
        const int bw0 = firstBinAtOrAfter(t.wDst, nbw, color.W), bw1 = firstBinAtOrAfter(t.wDst + t.W2, nbw, color.W);
        const Image v = virtualImage(color, sl.in[0], S.inPitch[0], t.hSrc, t.wSrc);
        const oidnb200_image vi{v.ptr, (int)v.format, v.W, v.H, v.pixelStride, v.rowStride};
        checkABI(oidnb200_autoexposure_bins_launch(&vi, bh0, bh1, bw0, bw1, S.bins[par], comp), "autoexposure bins");
      }
      checkCuda(cudaEventRecord(ce(S.evBins[par][e]), comp), "cudaEventRecord");
    }

  // ---- input scale (core/unet_filter.cpp:172-189)
  if (autoScale)
  {
    Engine* e0 = device->getEngine(0);
    e0->makeCurrent();
    cudaStream_t c0 = cs(e0->getAuxStream(Engine::Compute));
    for (int e = 0; e < E; ++e) checkCuda(cudaStreamWaitEvent(c0, ce(S.evBins[par][e]), 0), "cudaStreamWaitEvent");
    checkABI(oidnb200_autoexposure_reduce_launch(S.bins[par], S.numBins, S.scale[0] + par, c0), "autoexposure reduce");
    for (int e = 1; e < E; ++e)
      checkCuda(cudaMemcpyAsync(S.scale[e] + par, S.scale[0] + par, sizeof(float), cudaMemcpyDefault, c0), "cudaMemcpyAsync (input scale)");
    checkCuda(cudaEventRecord(ce(S.evScale[par]), c0), "cudaEventRecord");
    if (progress) { e0->setActiveStream(c0); report(e0); e0->setActiveStream(nullptr); }
  }

  // ---- the network on the staged tiles
  for (int e = 0; e < E; ++e)
  {
    Engine* eng = device->getEngine(e);
    Instance& inst = instances[e];
    eng->makeCurrent();
    cudaStream_t comp = cs(eng->getAuxStream(Engine::Compute));
    if (autoScale)
    {
      checkCuda(cudaStreamWaitEvent(comp, ce(S.evScale[par]), 0), "cudaStreamWaitEvent");
      inst.transferFunc->setInputScale(S.scale[e] + par);
    }
    else if (inputScaleDevPtr) inst.transferFunc->setInputScale(inputScaleDevPtr);
    else inst.transferFunc->setInputScale(std::isnan(inputScale) ? 1.f : inputScale);
    const size_t T = S.tilesOf[e].size();
    eng->setActiveStream(comp);
    try
    {
      for (size_t j = 0; j < T; ++j)
      {
        if (progress && progress->cancelled) break;
        const TileRect& t = tiles[S.tilesOf[e][j]];
        Slot& sl = S.slots[e][par * T + j];
        checkCuda(cudaStreamWaitEvent(comp, ce(sl.evIn), 0), "cudaStreamWaitEvent");
        checkCuda(cudaStreamWaitEvent(comp, ce(sl.evOut), 0), "cudaStreamWaitEvent"); // frame f-2's output has left the slot
        Image vin[3];
        for (int k = 0; k < 3; ++k)
          if (*ins[k]) vin[k] = virtualImage(*ins[k], sl.in[k], S.inPitch[k], t.hSrc, t.wSrc);
        inst.inputProcess->setSrc(vin[0], vin[1], vin[2]);
        inst.outputProcess->setDst(virtualImage(output, sl.out, S.outPitch, t.hDst, t.wDst));
        inst.graph->setProfiling(profiling);
        inst.graph->setFuseOutput(device->getInt("fuseOutput") != 0);
        inst.graph->setFusePairs(device->getInt("fusePairs") != 0);
        inst.graph->setOpCallback(progress ? std::function<void()>([&, eng]() { report(eng); }) : std::function<void()>());
        inst.inputProcess->setTile(t.hSrc, t.wSrc, t.hBuf, t.wBuf, t.H1, t.W1);
        inst.outputProcess->setTile(t.hOutBuf, t.wOutBuf, t.hDst, t.wDst, t.H2, t.W2);
        inst.graph->submit();
        inst.graph->setOpCallback(std::function<void()>());
        checkCuda(cudaEventRecord(ce(sl.evDone), comp), "cudaEventRecord");
      }
    }
    catch (...) { eng->setActiveStream(nullptr); throw; }
    eng->setActiveStream(nullptr);
  }

  // ---- copy-out; in-place filtering: no rectangle may be written before every tile of the frame has been read
  for (int e = 0; e < E; ++e)
  {
    Engine* eng = device->getEngine(e);
    eng->makeCurrent();
    cudaStream_t out = cs(eng->getAuxStream(Engine::CopyOut));
    const size_t T = S.tilesOf[e].size();
    if (inplace)
      for (int e2 = 0; e2 < E; ++e2)
        for (size_t j2 = 0; j2 < S.tilesOf[e2].size(); ++j2)
          checkCuda(cudaStreamWaitEvent(out, ce(S.slots[e2][par * S.tilesOf[e2].size() + j2].evIn), 0), "cudaStreamWaitEvent");
    for (size_t j = 0; j < T; ++j)
    {
      const TileRect& t = tiles[S.tilesOf[e][j]];
      Slot& sl = S.slots[e][par * T + j];
      checkCuda(cudaStreamWaitEvent(out, ce(sl.evDone), 0), "cudaStreamWaitEvent");
      uint8_t* dst = static_cast<uint8_t*>(output.ptr) + (size_t)t.hDst * output.rowStride + (size_t)t.wDst * output.pixelStride;
      checkCuda(cudaMemcpy2DAsync(dst, output.rowStride, sl.out, S.outPitch, (size_t)t.W2 * output.pixelStride, t.H2,
                                  cudaMemcpyDefault, out), "cudaMemcpy2DAsync (tile copy-out)");
      checkCuda(cudaEventRecord(ce(sl.evOut), out), "cudaEventRecord");
    }
    checkCuda(cudaEventRecord(ce(S.evEnd[e]), out), "cudaEventRecord");
  }

  // ---- join: the main stream(s) continue when the whole frame is in the output image (Device::submitBarrier).
  // A stream the caller supplied is joined now (the caller may enqueue anything behind this call). Streams the
  // engines own are joined lazily -- before the next buffer operation, in-place frame or synchronisation
  // (Device::joinStaged) -- so back-to-back staged frames overlap: copy-in of frame f+1 and copy-out of frame
  // f-1 run under the convolutions of frame f.
  device->deferJoin(S.evEnd);
  bool userStream = false;
  for (int e = 0; e < E; ++e) userStream |= !device->getEngine(e)->ownsStream();
  if (userStream) device->joinStaged();
}

// The frame with every image dereferenced in place by the kernels (the reference's contract).
void UNetFilter::submitFrame(const std::shared_ptr<ProgressState>& progress)
{
  const int numEngines = device->getNumEngines();
  const int profiling = device->getInt("profile");
  auto report = [&](Engine* e) {
    if (progress) { auto p = progress; e->submitHostFunc([p]() { p->update(1.); }); }
  };
  auto checkCancel = [&]() {
    if (progress && progress->cancelled)
    {
      device->wait();
      throw Exception(Error::Cancelled, "execution was cancelled");
    }
  };
  auto setScale = [&](float v) { for (auto& inst : instances) inst.transferFunc->setInputScale(v); };
  auto setScalePtr = [&](const float* p) { for (auto& inst : instances) inst.transferFunc->setInputScale(p); };

  // input scale (core/unet_filter.cpp:172-189)
  if (inputScaleDevPtr)
    setScalePtr(inputScaleDevPtr);
  else if (std::isnan(inputScale))
  {
    if (hdr)
    {
      autoexposure->setSrc(color);
      device->getEngine(0)->makeCurrent();
      autoexposure->submit();
      report(device->getEngine(0));
      device->submitBarrier();
      setScalePtr(autoexposure->getDstPtr());
    }
    else
      setScale(1.f);
  }
  else
    setScale(inputScale);

  for (auto& inst : instances)
  {
    inst.inputProcess->setSrc(color, albedo, normal);
    inst.outputProcess->setDst(outputTemp ? outputTemp : output);
  }

  int tileIndex = 0, globalIndex = 0;
  std::vector<const TileRect*> mine;
  for (const TileRect& t : tiles)
  {
    if (globalIndex++ % numShards != shardIndex) continue; // another process's tile
    checkCancel();
    Engine* eng = device->getEngine(tileIndex % numEngines);
    Instance& inst = instances[tileIndex % numEngines];
    inst.graph->setProfiling(profiling);
    inst.graph->setFuseOutput(device->getInt("fuseOutput") != 0);
    inst.graph->setFusePairs(device->getInt("fusePairs") != 0);
    inst.graph->setOpCallback(progress ? std::function<void()>([&, eng]() { report(eng); }) : std::function<void()>());
    inst.inputProcess->setTile(t.hSrc, t.wSrc, t.hBuf, t.wBuf, t.H1, t.W1);
    inst.outputProcess->setTile(t.hOutBuf, t.wOutBuf, t.hDst, t.wDst, t.H2, t.W2);
    inst.graph->submit();
    inst.graph->setOpCallback(std::function<void()>());
    mine.push_back(&t);
    ++tileIndex;
  }
  device->submitBarrier();

  if (outputTemp)
  {
    // in-place filtering of a multi-tile frame went through a temporary (core/unet_filter.cpp:245-249)
    device->getEngine(0)->makeCurrent();
    if (numShards == 1)
    {
      imageCopy->setSrc(outputTemp);
      imageCopy->setDst(output);
      imageCopy->submit();
    }
    else
      // a shard has written only its own tiles: copy back exactly those rectangles (the rest of the
      // temporary is uninitialised, and the other shards' rectangles of `output` are theirs)
      for (const TileRect* t : mine)
      {
        auto view = [&](const Image& im) {
          Image v = im;
          v.ptr = static_cast<uint8_t*>(im.ptr) + (size_t)t->hDst * im.rowStride + (size_t)t->wDst * im.pixelStride;
          v.H = t->H2; v.W = t->W2;
          return v;
        };
        imageCopy->setSrc(view(outputTemp));
        imageCopy->setDst(view(output));
        imageCopy->submit();
      }
    report(device->getEngine(0));
  }
}

void UNetFilter::execute(SyncMode sync)
{
  if (dirty) throw Exception(Error::InvalidOperation, "changes to the filter are not committed");
  if (plan.H <= 0 || plan.W <= 0) return;
  const int numEngines = device->getNumEngines();
  const bool staged = wantStaging();
  lastStaged = staged;
  const bool autoScale = hdr && std::isnan(inputScale) && !inputScaleDevPtr;
  const int profiling = device->getInt("profile");

  // Progress: one unit per op of every tile (+1 for the autoexposure, +1 for the final copy of an in-place
  // multi-tile frame), as core/unet_filter.cpp:155-168 counts it.
  std::shared_ptr<ProgressState> progress;
  if (progressFunc)
  {
    progress = std::make_shared<ProgressState>();
    progress->func = progressFunc; progress->userPtr = progressUserPtr;
    const double myTiles = (double)((tiles.size() + numShards - 1 - shardIndex) / numShards);
    progress->total = myTiles * getNumOps() + (autoScale ? 1 : 0) + ((outputTemp && !staged) ? 1 : 0);
    if (!progressFunc(progressUserPtr, 0.)) throw Exception(Error::Cancelled, "execution was cancelled");
  }

  // Frame-stream path: replay (or capture) the frame as one CUDA graph.
  const bool graphable = device->getInt("graph") != 0 && numEngines == 1 && !progress && !profiling && !staged;
  if (staged)
    submitFrameStaged(progress);
  else if (graphable)
  {
    device->joinStaged();
    Engine* e0 = device->getEngine(0);
    cudaStream_t st = static_cast<cudaStream_t>(e0->getStream());
    const std::vector<uint64_t> key = frameKey();
    e0->makeCurrent();
    FrameGraph* fg = nullptr;
    for (FrameGraph& g : frameGraphs)
      if (g.key == key) fg = &g;
    if (!fg)
    {
      if (frameGraphs.size() >= maxFrameGraphs)
      {
        size_t lru = 0;
        for (size_t i = 1; i < frameGraphs.size(); ++i)
          if (frameGraphs[i].lastUse < frameGraphs[lru].lastUse) lru = i;
        if (frameGraphs[lru].exec) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(frameGraphs[lru].exec));
        frameGraphs.erase(frameGraphs.begin() + lru);
      }
      frameGraphs.emplace_back();
      fg = &frameGraphs.back();
      fg->key = key;
    }
    fg->lastUse = ++frameGraphClock;
    if (fg->exec)
      checkCuda(cudaGraphLaunch(static_cast<cudaGraphExec_t>(fg->exec), st), "cudaGraphLaunch");
    else if (fg->seen++ == 0)
      submitFrame(progress); // first frame with these arguments: run it eagerly (also warms one-time kernel attributes)
    else
    {
      cudaGraph_t g = nullptr;
      checkCuda(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
      try
      {
        submitFrame(progress);
      }
      catch (...)
      {
        cudaStreamEndCapture(st, &g);
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        throw;
      }
      checkCuda(cudaStreamEndCapture(st, &g), "cudaStreamEndCapture");
      cudaGraphExec_t exec = nullptr;
      const cudaError_t ie = cudaGraphInstantiate(&exec, g, 0);
      cudaGraphDestroy(g);
      checkCuda(ie, "cudaGraphInstantiate");
      fg->exec = exec;
      checkCuda(cudaGraphLaunch(exec, st), "cudaGraphLaunch");
    }
  }
  else
  {
    device->joinStaged();
    submitFrame(progress);
  }

  // events (mode 1) are read back per frame; in-frame stamps (mode 2) stay on the device until the profile is asked
  // for, so that consecutive frames are stamped as they really run, back to back
  if (profiling == 1)
    for (auto& inst : instances) inst.graph->collectProfile(profile);

  if (sync == SyncMode::Blocking || progress)
  {
    device->wait();
    if (progress && progress->cancelled) throw Exception(Error::Cancelled, "execution was cancelled");
  }
}

std::vector<Graph::OpTime> UNetFilter::getProfile()
{
  for (auto& inst : instances)
    if (inst.graph->hasPendingStamps()) inst.graph->collectProfile(profile);
  return profile;
}

// ------------------------------------------------------------------------------------------------
// RTFilter (core/rt_filter.cpp)
// ------------------------------------------------------------------------------------------------
RTFilter::RTFilter(Device* device) : UNetFilter(device)
{
  // weight file stems = the reference's blob names (CMakeLists.txt:55-86, core/rt_filter.cpp:34-60)
  models.hdr           = {"rt_hdr", "rt_hdr_small", nullptr};
  models.hdr_alb       = {"rt_hdr_alb", "rt_hdr_alb_small", nullptr};
  models.hdr_alb_nrm   = {"rt_hdr_alb_nrm", "rt_hdr_alb_nrm_small", nullptr};
  models.hdr_calb_cnrm = {"rt_hdr_calb_cnrm", "rt_hdr_calb_cnrm_small", "rt_hdr_calb_cnrm_large"};
  models.ldr           = {"rt_ldr", "rt_ldr_small", nullptr};
  models.ldr_alb       = {"rt_ldr_alb", "rt_ldr_alb_small", nullptr};
  models.ldr_alb_nrm   = {"rt_ldr_alb_nrm", "rt_ldr_alb_nrm_small", nullptr};
  models.ldr_calb_cnrm = {"rt_ldr_calb_cnrm", "rt_ldr_calb_cnrm_small", nullptr};
  models.alb           = {"rt_alb", nullptr, "rt_alb_large"};
  models.nrm           = {"rt_nrm", nullptr, "rt_nrm_large"};
}

std::shared_ptr<TransferFunction> RTFilter::newTransferFunc()
{
  if (srgb || (!color && normal)) return std::make_shared<TransferFunction>(TransferType::Linear);
  if (hdr) return std::make_shared<TransferFunction>(TransferType::PU);
  return std::make_shared<TransferFunction>(TransferType::SRGB);
}

void RTFilter::setImage(const std::string& name, const Image& image)
{
  if (name == "color") setParam(color, image);
  else if (name == "albedo") setParam(albedo, image);
  else if (name == "normal") setParam(normal, image);
  else if (name == "output") setParam(output, image);
  else warnUnknown(device, name);
  dirty = true;
}

void RTFilter::unsetImage(const std::string& name)
{
  if (name == "color") removeParam(color);
  else if (name == "albedo") removeParam(albedo);
  else if (name == "normal") removeParam(normal);
  else if (name == "output") removeParam(output);
  else warnUnknown(device, name);
  dirty = true;
}

void RTFilter::setInt(const std::string& name, int value)
{
  if (name == "hdr") setParam(hdr, value);
  else if (name == "srgb") setParam(srgb, value);
  else if (name == "cleanAux") setParam(cleanAux, value);
  else UNetFilter::setInt(name, value);
  dirty = true;
}

int RTFilter::getInt(const std::string& name)
{
  if (name == "hdr") return hdr;
  if (name == "srgb") return srgb;
  if (name == "cleanAux") return cleanAux;
  return UNetFilter::getInt(name);
}

// ------------------------------------------------------------------------------------------------
// RTLightmapFilter (core/rtlightmap_filter.cpp)
// ------------------------------------------------------------------------------------------------
RTLightmapFilter::RTLightmapFilter(Device* device) : UNetFilter(device)
{
  hdr = true; // core/rtlightmap_filter.cpp:14
  models.hdr = {"rtlightmap_hdr", nullptr, nullptr};
  models.dir = {"rtlightmap_dir", nullptr, nullptr};
}

std::shared_ptr<TransferFunction> RTLightmapFilter::newTransferFunc()
{
  return std::make_shared<TransferFunction>(hdr ? TransferType::Log : TransferType::Linear);
}

void RTLightmapFilter::setImage(const std::string& name, const Image& image)
{
  if (name == "color") setParam(color, image);
  else if (name == "output") setParam(output, image);
  else warnUnknown(device, name);
  dirty = true;
}

void RTLightmapFilter::unsetImage(const std::string& name)
{
  if (name == "color") removeParam(color);
  else if (name == "output") removeParam(output);
  else warnUnknown(device, name);
  dirty = true;
}

void RTLightmapFilter::setInt(const std::string& name, int value)
{
  if (name == "directional")
  {
    setParam(directional, value);
    hdr = !directional; // core/rtlightmap_filter.cpp:58-62
  }
  else UNetFilter::setInt(name, value);
  dirty = true;
}

int RTLightmapFilter::getInt(const std::string& name)
{
  if (name == "directional") return directional;
  return UNetFilter::getInt(name);
}

} // namespace oidnb200
