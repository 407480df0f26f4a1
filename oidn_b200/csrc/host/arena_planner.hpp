// Scratch-arena planner: places the UNet's intermediate tensors in one buffer so that tensors
// whose lifetimes (first op .. last op) do not intersect may share bytes. Same job as the
// reference's ArenaPlanner (core/arena_planner.cpp:23-152); own algorithm (first-fit over
// lifetime conflicts, largest first). In the fused B200 graph up-sampled and un-pooled tensors
// never exist, so there is no "force adjacent" requirement (that served ConcatConvCHW only).
#pragma once
#include "base.hpp"
#include <vector>

namespace oidnb200 {

class ArenaPlanner
{
public:
  // New allocation produced by op `opID`; returns its id.
  int newAlloc(int opID, size_t byteSize, size_t byteAlignment = memoryAlignment);
  // Op `opID` reads allocation `allocID` (extends its lifetime).
  void addDep(int opID, int allocID);
  void commit();
  void clear();
  size_t getByteSize() const;
  size_t getAllocByteOffset(int allocID) const;
  int numAllocs() const { return (int)allocs.size(); }
  // test hook: true if no two lifetime-overlapping allocations share bytes
  bool validate() const;

private:
  struct Alloc { size_t size, align, offset; int first, last; };
  std::vector<Alloc> allocs;
  size_t total = 0;
  bool dirty = true;
};

} // namespace oidnb200
