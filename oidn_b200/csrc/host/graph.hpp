// Graph: the op list of one UNet instance on one engine, its weights and its scratch plan.
// Interface follows core/graph.h:27-64 (addInputProcess / addConv / addConcatConv /
// addOutputProcess / getScratchByteSize / setScratch / finalize / submit); the construction is the
// fused B200 design:
//   * PostOp::Pool runs in the producing conv's epilogue (the un-pooled tensor never exists);
//   * PostOp::Upsample produces a *virtual* tensor: it is stored at the conv's own resolution and
//     every consumer reads it through a TMA map with a stride-0 duplication axis;
//   * ConcatConv is one kernel with two K segments reading both sources in place.
#pragma once
#include "arena_planner.hpp"
#include "engine.hpp"
#include "tza.hpp"
#include <vector>

namespace oidnb200 {

class Graph
{
public:
  // A value flowing between ops: stored tensor + whether consumers see it 2x nearest-upsampled.
  struct Value
  {
    int id = -1;
  };

  Graph(Engine* engine, std::shared_ptr<TensorMap> constTensors);
  ~Graph();

  Value addInputProcess(const std::string& name, const TensorDesc& dstDesc,
                        const std::shared_ptr<TransferFunction>& tf, bool hdr, bool snorm);
  void addOutputProcess(const std::string& name, Value src, const std::shared_ptr<TransferFunction>& tf,
                        bool hdr, bool snorm);
  Value addConv(const std::string& name, Value src, Activation activation, PostOp postOp = PostOp::None);
  Value addConcatConv(const std::string& name, Value src1, Value src2, Activation activation);

  size_t getScratchByteSize();               // planned arena size for the intermediate tensors
  size_t getPrivateByteSize() const { return privateByteSize; } // packed weights + biases
  void setScratch(void* base, size_t byteSize);
  void finalize();                           // upload weights, bind pointers, encode tensor maps
  void submit();                             // launch every op in order on the engine's stream
  void clear();

  // Per-op device timing (the reference's compile-time OIDN_MICROBENCH, core/graph.cpp:460-525, as
  // a run-time switch): submit() brackets every op with CUDA events; collectProfile() waits for the
  // stream and adds the elapsed times to `out` (one entry per op, in op order).
  struct OpTime { std::string name; int kind; double ms; int launches; };
  void setProfiling(int mode) { profiling = mode; }
  // mode 2: no events; every conv records when its grid really ran inside the frame (%globaltimer stamps written by
  // the kernel: launches of a frame overlap through programmatic dependent launch, which events around an op
  // serialise away). submit() then waits for the stream and accumulates per conv its interval, plus -- in an extra
  // entry "conv_union" (kind 3) -- the length of the union of all conv intervals: the time the frame spends in convs.
  // Called after every op has been enqueued (also for an output process folded into the last conv), so
  // progress advances per op as in the reference (core/op.cpp:8-22, every op has work amount 1).
  void setOpCallback(std::function<void()> cb) { opCallback = std::move(cb); }
  // Output process folded into the last conv's epilogue when the output image allows it (default on;
  // device parameter "fuseOutput"). Decided per submit(): it depends on the image set for the frame.
  void setFuseOutput(bool on) { fuseOutput = on; }
  // Conv -> conv pairs whose intermediate tensor has no other reader run as one launch (kernels/conv_pair_tc.cu;
  // device parameter "fusePairs", default on): enc_conv0 -> enc_conv1, dec_conv1b -> dec_conv0 and their
  // counterparts in the small / large nets. Bit-identical to the two launches.
  void setFusePairs(bool on) { fusePairs = on; }
  int getNumPairs() const { return (int)pairs.size(); }
  void collectProfile(std::vector<OpTime>& out);
  bool hasPendingStamps() const { return stampSlice > 0; }

  const std::shared_ptr<InputProcess>& getInputProcess() const { return inputProcess; }
  const std::shared_ptr<OutputProcess>& getOutputProcess() const { return outputProcess; }
  int getNumOps() const { return (int)ops.size(); }
  const std::vector<std::shared_ptr<Op>>& getOps() const { return ops; }
  const ArenaPlanner& getPlanner() const { return planner; }
  // logical (consumer-visible) dims of a value
  TensorDesc getLogicalDesc(Value v) const;

private:
  struct Node
  {
    TensorDesc stored;   // what lives in the arena
    bool upsampled;      // consumers see 2H x 2W
    int allocID;
    int opID;
  };
  struct ConvRecord
  {
    std::shared_ptr<Conv> conv;
    int src1, src2, dst;        // node ids (src2 = -1 for plain conv)
    size_t weightOffset, biasOffset;
    const ConstTensor *weight, *bias;
    int I1, I2;
  };

  const ConstTensor& findConst(const std::string& name) const;
  Value addConvImpl(const std::string& name, Value src1, Value src2, Activation activation, PostOp postOp);

  Engine* engine;
  std::shared_ptr<TensorMap> constTensors;
  std::vector<std::shared_ptr<Op>> ops;
  std::vector<Node> nodes;
  std::vector<ConvRecord> convs;
  std::shared_ptr<InputProcess> inputProcess;
  std::shared_ptr<OutputProcess> outputProcess;
  int inputNode = -1, outputSrcNode = -1;
  std::shared_ptr<Conv> outputConv;   // producer of the output process's source, if it is a conv
  bool fuseOutput = true;
  bool fusePairs = true;
  struct PairRecord { int opA, opB; std::unique_ptr<ConvPair> pair; };
  std::vector<PairRecord> pairs;
  ArenaPlanner planner;
  bool planned = false, finalized = false;
  size_t privateByteSize = 0;
  void* scratchBase = nullptr;
  size_t scratchSize = 0;
  void* weightBuffer = nullptr;
  std::function<void()> opCallback;
  int profiling = 0;
  struct Stamp { int op; void* e0; void* e1; };
  std::vector<Stamp> stamps;
  void* stampBuf = nullptr;                       // device: kStampSlices x (2 x uint64 per conv): one slice per submit, so
                                                  // consecutive frames are stamped without a host sync between them
  static constexpr int kStampSlices = 64;
  int stampSlice = 0;                             // slices used since the last collect
  std::vector<unsigned long long> stampHost;      // init pattern / read-back
  std::vector<double> stampMs;                    // per conv (accumulated), last entry: union
  std::vector<int> stampLaunches;
  void collectStamps();
};

} // namespace oidnb200
