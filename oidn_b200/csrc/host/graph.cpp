#include "graph.hpp"
#include "../kernels/common.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <cstring>

namespace oidnb200 {

Graph::Graph(Engine* engine, std::shared_ptr<TensorMap> constTensors)
  : engine(engine), constTensors(std::move(constTensors)) {}

Graph::~Graph() { clear(); }

void Graph::clear()
{
  pairs.clear();
  ops.clear(); nodes.clear(); convs.clear();
  inputProcess.reset(); outputProcess.reset(); outputConv.reset();
  inputNode = outputSrcNode = -1;
  planner.clear();
  planned = finalized = false;
  for (const Stamp& st : stamps) { cudaEventDestroy(static_cast<cudaEvent_t>(st.e0)); cudaEventDestroy(static_cast<cudaEvent_t>(st.e1)); }
  stamps.clear();
  privateByteSize = 0;
  if (weightBuffer) { engine->free(weightBuffer); weightBuffer = nullptr; }
  if (stampBuf) { engine->free(stampBuf); stampBuf = nullptr; }
}

const ConstTensor& Graph::findConst(const std::string& name) const
{
  auto it = constTensors->find(name);
  if (it == constTensors->end()) throw std::invalid_argument("weights blob has no tensor named '" + name + "'");
  return it->second;
}

TensorDesc Graph::getLogicalDesc(Value v) const
{
  const Node& n = nodes.at(v.id);
  return n.upsampled ? TensorDesc{n.stored.C, n.stored.H * 2, n.stored.W * 2} : n.stored;
}

Graph::Value Graph::addInputProcess(const std::string& name, const TensorDesc& dstDesc,
                                    const std::shared_ptr<TransferFunction>& tf, bool hdr, bool snorm)
{
  if (finalized) throw std::logic_error("graph cannot be changed after finalization");
  const int opID = (int)ops.size();
  inputProcess = engine->newInputProcess(dstDesc, tf, hdr, snorm);
  inputProcess->setName(name);
  ops.push_back(inputProcess);
  nodes.push_back(Node{dstDesc, false, planner.newAlloc(opID, dstDesc.byteSize()), opID});
  inputNode = (int)nodes.size() - 1;
  return Value{inputNode};
}

void Graph::addOutputProcess(const std::string& name, Value src, const std::shared_ptr<TransferFunction>& tf,
                             bool hdr, bool snorm)
{
  if (finalized) throw std::logic_error("graph cannot be changed after finalization");
  const Node& n = nodes.at(src.id);
  if (n.upsampled || n.stored.C < 3) throw std::invalid_argument("invalid output process source");
  const int opID = (int)ops.size();
  outputProcess = engine->newOutputProcess(n.stored, tf, hdr, snorm);
  outputProcess->setName(name);
  ops.push_back(outputProcess);
  planner.addDep(opID, n.allocID);
  outputSrcNode = src.id;
  outputConv = std::dynamic_pointer_cast<Conv>(ops.at(n.opID));
}

Graph::Value Graph::addConv(const std::string& name, Value src, Activation activation, PostOp postOp)
{
  return addConvImpl(name, src, Value{-1}, activation, postOp);
}

Graph::Value Graph::addConcatConv(const std::string& name, Value src1, Value src2, Activation activation)
{
  return addConvImpl(name, src1, src2, activation, PostOp::None);
}

Graph::Value Graph::addConvImpl(const std::string& name, Value s1, Value s2, Activation activation, PostOp postOp)
{
  if (finalized) throw std::logic_error("graph cannot be changed after finalization");
  const Node n1 = nodes.at(s1.id);
  const bool concat = s2.id >= 0;
  const TensorDesc l1 = getLogicalDesc(s1);
  ConvDesc d;
  d.src1 = n1.stored;
  d.src1Upsampled = n1.upsampled;
  d.H = l1.H; d.W = l1.W;
  int I2 = 0;
  if (concat)
  {
    const Node& n2 = nodes.at(s2.id);
    if (n2.upsampled || n2.stored.H != l1.H || n2.stored.W != l1.W)
      throw std::invalid_argument("invalid concat+conv source descriptor"); // core/concat_conv.cpp:17
    d.src2 = n2.stored;
    I2 = n2.stored.C;
  }
  // weights: [O, I, 3, 3] "oihw", bias [O] "x" (core/graph.cpp:82-89)
  const ConstTensor& w = findConst(name + ".weight");
  const ConstTensor& b = findConst(name + ".bias");
  if (w.dims.size() != 4 || w.dims[2] != 3 || w.dims[3] != 3 || b.dims.size() != 1 || b.dims[0] != w.dims[0] ||
      w.dims[1] != n1.stored.C + I2)
    throw std::invalid_argument("invalid convolution weight/bias");
  d.outC = w.dims[0];
  d.activation = activation;
  d.postOp = postOp;
  if (postOp == PostOp::Pool && ((d.H & 1) || (d.W & 1)))
    throw std::invalid_argument("invalid pooling source shape"); // core/conv.cpp:27

  const int opID = (int)ops.size();
  auto conv = engine->newConv(d);
  conv->setName(name);
  ops.push_back(conv);
  const TensorDesc dst = conv->getDstDesc();
  nodes.push_back(Node{dst, postOp == PostOp::Upsample, planner.newAlloc(opID, dst.byteSize()), opID});
  planner.addDep(opID, n1.allocID);
  if (concat) planner.addDep(opID, nodes.at(s2.id).allocID);
  // A conv that may run fused with its producer (finalize(): ConvPair) reads the PRODUCER's source while it writes
  // its own destination: that source must stay alive (and un-aliased) through this op.
  if (!concat)
    for (const ConvRecord& pr : convs)
      if (pr.dst == s1.id && pr.src2 < 0)
      {
        const int ca = round_up(pr.conv->getDesc().outC, 16), c1 = pr.conv->getDesc().src1.paddedC();
        if ((ca == 32 || ca == 64) && c1 <= 64 && round_up(d.outC, 16) <= 64 && !pr.conv->getDesc().src1Upsampled &&
            pr.conv->getDesc().postOp == PostOp::None)
          planner.addDep(opID, nodes.at(pr.src1).allocID);
      }

  ConvRecord r;
  r.conv = conv; r.src1 = s1.id; r.src2 = s2.id; r.dst = (int)nodes.size() - 1;
  r.weight = &w; r.bias = &b; r.I1 = n1.stored.C; r.I2 = I2;
  r.weightOffset = round_up(privateByteSize, memoryAlignment);
  r.biasOffset = round_up(r.weightOffset + conv->getWeightByteSize(), memoryAlignment);
  privateByteSize = r.biasOffset + conv->getBiasByteSize();
  convs.push_back(r);
  planned = false;
  return Value{r.dst};
}

size_t Graph::getScratchByteSize()
{
  if (!planned)
  {
    planner.commit();
    planned = true;
  }
  return planner.getByteSize();
}

void Graph::setScratch(void* base, size_t byteSize)
{
  if (byteSize < getScratchByteSize()) throw std::invalid_argument("graph scratch buffer is too small");
  scratchBase = base;
  scratchSize = byteSize;
  finalized = false;
}

static std::vector<uint16_t> asHalfBits(const ConstTensor& t)
{
  std::vector<uint16_t> out(t.count());
  if (t.dtype == 'h')
    memcpy(out.data(), t.data, out.size() * 2);
  else
  {
    const float* f = static_cast<const float*>(t.data);
    for (size_t i = 0; i < out.size(); ++i) out[i] = float_to_half_bits(f[i]);
  }
  return out;
}

void Graph::finalize()
{
  if (!scratchBase && getScratchByteSize() > 0) throw std::logic_error("graph scratch not set");
  engine->makeCurrent();
  auto ptrOf = [&](int node) -> void* {
    return static_cast<uint8_t*>(scratchBase) + planner.getAllocByteOffset(nodes[node].allocID);
  };

  // Weight reorder + pad + upload (core/graph.cpp:113-140), once per graph.
  if (!weightBuffer && privateByteSize > 0)
  {
    std::vector<uint8_t> host(privateByteSize, 0);
    for (const ConvRecord& r : convs)
    {
      const std::vector<uint16_t> w = asHalfBits(*r.weight), b = asHalfBits(*r.bias);
      r.conv->packWeight(w.data(), r.weight->dims[0], r.I1, r.I2, host.data() + r.weightOffset);
      r.conv->packBias(b.data(), r.bias->dims[0], host.data() + r.biasOffset);
    }
    weightBuffer = engine->malloc(privateByteSize);
    checkCuda(cudaMemcpy(weightBuffer, host.data(), privateByteSize, cudaMemcpyHostToDevice), "weight upload");
  }

  for (const ConvRecord& r : convs)
  {
    r.conv->setSrc(ptrOf(r.src1), r.src2 >= 0 ? ptrOf(r.src2) : nullptr);
    r.conv->setWeight(static_cast<uint8_t*>(weightBuffer) + r.weightOffset);
    r.conv->setBias(static_cast<uint8_t*>(weightBuffer) + r.biasOffset);
    r.conv->setDst(ptrOf(r.dst));
    r.conv->finalize();
  }
  if (inputProcess) inputProcess->setDst(ptrOf(inputNode));
  if (outputProcess) outputProcess->setSrc(ptrOf(outputSrcNode));

  // conv -> conv pairs: B is a plain conv reading only A's tensor, and nothing else reads that tensor
  pairs.clear();
  for (size_t i = 0; i + 1 < convs.size(); ++i)
  {
    const ConvRecord& A = convs[i];
    const ConvRecord& B = convs[i + 1];
    if (B.src1 != A.dst || B.src2 >= 0 || A.src2 >= 0) continue;
    int readers = outputSrcNode == A.dst ? 1 : 0;
    for (const ConvRecord& r : convs) readers += (r.src1 == A.dst) + (r.src2 == A.dst);
    if (readers != 1) continue;
    if (!pairs.empty() && pairs.back().opB == nodes[A.dst].opID) continue; // a conv belongs to one pair only
    auto pr = ConvPair::tryCreate(*A.conv, *B.conv);
    if (!pr) continue;
    pairs.push_back(PairRecord{nodes[A.dst].opID, nodes[B.dst].opID, std::move(pr)});
  }
  finalized = true;
}

void Graph::submit()
{
  if (!finalized) throw std::logic_error("graph not finalized");
  engine->makeCurrent();
  // The output process runs inside the last conv's epilogue when the kernel supports the frame's
  // output image; otherwise (or with fuseOutput off) it is the separate pass of the reference.
  const bool fused = outputProcess && outputConv && outputProcess->fuseInto(*outputConv, fuseOutput);
  if (profiling == 2 || stampBuf)
  {
    // in-frame stamps on (mode 2) or to be switched off again
    const size_t n = convs.size() * 2;
    if (profiling == 2)
    {
      if (!stampBuf) stampBuf = engine->malloc((size_t)kStampSlices * n * sizeof(unsigned long long));
      if (stampSlice == kStampSlices) collectStamps();   // all slices used: read them back (waits for the stream)
      stampHost.assign(n, 0ull);
      for (size_t i = 0; i < convs.size(); ++i) stampHost[2 * i] = ~0ull;
      unsigned long long* slice = static_cast<unsigned long long*>(stampBuf) + (size_t)stampSlice * n;
      checkCuda(cudaMemcpyAsync(slice, stampHost.data(), n * sizeof(unsigned long long), cudaMemcpyHostToDevice,
                                static_cast<cudaStream_t>(engine->getStream())), "cudaMemcpyAsync (stamps)");
      for (size_t i = 0; i < convs.size(); ++i)
        oidnb200_conv_set_stamps(convs[i].conv->getHandle(), slice + 2 * i);
      ++stampSlice;
    }
    else
    {
      if (stampSlice > 0) collectStamps();
      for (size_t i = 0; i < convs.size(); ++i) oidnb200_conv_set_stamps(convs[i].conv->getHandle(), nullptr);
      engine->wait(); engine->free(stampBuf); stampBuf = nullptr;
    }
  }
  // a fused pair replaces the launches of its two convs: A is skipped, the pair runs in B's place
  auto submitOp = [&](size_t i) {
    if (fusePairs)
      for (const PairRecord& pr : pairs)
      {
        if ((int)i == pr.opA) return;
        if ((int)i == pr.opB) { pr.pair->submit(); return; }
      }
    ops[i]->submit();
  };
  if (profiling != 1)
  {
    for (size_t i = 0; i < ops.size(); ++i)
    {
      if (!(fused && ops[i] == outputProcess)) submitOp(i);
      if (opCallback) opCallback();
    }
    return;
  }
  cudaStream_t st = static_cast<cudaStream_t>(engine->getStream());
  for (size_t i = 0; i < ops.size(); ++i)
  {
    if (fused && ops[i] == outputProcess) { if (opCallback) opCallback(); continue; }
    cudaEvent_t e0, e1;
    checkCuda(cudaEventCreate(&e0), "cudaEventCreate");
    checkCuda(cudaEventCreate(&e1), "cudaEventCreate");
    cudaEventRecord(e0, st);
    submitOp(i);
    cudaEventRecord(e1, st);
    stamps.push_back(Stamp{(int)i, e0, e1});
    if (opCallback) opCallback();
  }
}

void Graph::collectStamps()
{
  engine->wait();
  const size_t n = convs.size();
  std::vector<unsigned long long> all((size_t)stampSlice * 2 * n);
  if (!all.empty())
    checkCuda(cudaMemcpy(all.data(), stampBuf, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost), "cudaMemcpy (stamps)");
  stampMs.resize(n + 1, 0.); stampLaunches.resize(n + 1, 0);
  for (int sl = 0; sl < stampSlice; ++sl)
  {
    const unsigned long long* st = all.data() + (size_t)sl * 2 * n;
    std::vector<std::pair<unsigned long long, unsigned long long>> iv;
    for (size_t i = 0; i < n; ++i)
    {
      const unsigned long long t0 = st[2 * i], t1 = st[2 * i + 1];
      if (t1 <= t0) continue; // not launched (first conv of a fused pair)
      stampMs[i] += (double)(t1 - t0) * 1e-6;
      stampLaunches[i] += 1;
      iv.emplace_back(t0, t1);
    }
    std::sort(iv.begin(), iv.end());
    unsigned long long total = 0, curEnd = 0;
    for (const auto& v : iv)
    {
      if (v.first >= curEnd) { total += v.second - v.first; curEnd = v.second; }
      else if (v.second > curEnd) { total += v.second - curEnd; curEnd = v.second; }
    }
    stampMs[n] += (double)total * 1e-6;
    stampLaunches[n] += 1;
  }
  stampSlice = 0;
}

void Graph::collectProfile(std::vector<OpTime>& out)
{
  engine->wait();
  if (stampSlice > 0) collectStamps();
  if (!stampMs.empty())
  {
    // in-frame stamp mode: one entry per op (convs filled) + the union entry
    if (out.size() != ops.size() + 1)
    {
      out.clear();
      for (auto& op : ops)
      {
        int kind = 0;
        if (dynamic_cast<InputProcess*>(op.get())) kind = 1;
        else if (dynamic_cast<OutputProcess*>(op.get())) kind = 2;
        out.push_back(OpTime{op->getName(), kind, 0., 0});
      }
      out.push_back(OpTime{"conv_union", 3, 0., 0});
    }
    for (size_t i = 0; i < convs.size(); ++i)
      for (size_t o = 0; o < ops.size(); ++o)
        if (ops[o] == convs[i].conv) { out[o].ms += stampMs[i]; out[o].launches += stampLaunches[i]; }
    out.back().ms += stampMs[convs.size()];
    out.back().launches += stampLaunches[convs.size()];
    stampMs.clear(); stampLaunches.clear();
    return;
  }
  if (out.size() < ops.size())
  {
    out.clear();
    for (auto& op : ops)
    {
      int kind = 0;
      if (dynamic_cast<InputProcess*>(op.get())) kind = 1;
      else if (dynamic_cast<OutputProcess*>(op.get())) kind = 2;
      out.push_back(OpTime{op->getName(), kind, 0., 0});
    }
  }
  for (const Stamp& s : stamps)
  {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(s.e0), static_cast<cudaEvent_t>(s.e1));
    out[s.op].ms += ms;
    out[s.op].launches += 1;
    cudaEventDestroy(static_cast<cudaEvent_t>(s.e0));
    cudaEventDestroy(static_cast<cudaEvent_t>(s.e1));
  }
  stamps.clear();
}

} // namespace oidnb200
