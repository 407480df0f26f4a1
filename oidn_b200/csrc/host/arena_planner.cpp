#include "arena_planner.hpp"
#include <algorithm>
#include <numeric>

namespace oidnb200 {

int ArenaPlanner::newAlloc(int opID, size_t byteSize, size_t byteAlignment)
{
  if (byteAlignment == 0 || (byteAlignment & (byteAlignment - 1)))
    throw std::invalid_argument("arena alignment must be a power of two");
  allocs.push_back({byteSize, byteAlignment, 0, opID, opID});
  dirty = true;
  return (int)allocs.size() - 1;
}

void ArenaPlanner::addDep(int opID, int allocID)
{
  if (allocID < 0 || allocID >= (int)allocs.size()) throw std::out_of_range("invalid arena allocation id");
  Alloc& a = allocs[allocID];
  a.first = std::min(a.first, opID);
  a.last = std::max(a.last, opID);
  dirty = true;
}

void ArenaPlanner::commit()
{
  std::vector<int> order(allocs.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return allocs[a].size > allocs[b].size; });
  std::vector<int> placed;
  total = 0;
  for (int id : order)
  {
    Alloc& a = allocs[id];
    // byte ranges of already placed allocations alive at the same time, by offset
    std::vector<std::pair<size_t, size_t>> busy;
    for (int pid : placed)
    {
      const Alloc& p = allocs[pid];
      if (p.first <= a.last && a.first <= p.last) busy.emplace_back(p.offset, p.offset + p.size);
    }
    std::sort(busy.begin(), busy.end());
    size_t off = 0;
    for (const auto& b : busy)
    {
      if (off + a.size <= b.first) break;                 // fits in the gap before this range
      if (b.second > off) off = round_up(b.second, a.align);
    }
    a.offset = off;
    total = std::max(total, off + a.size);
    placed.push_back(id);
  }
  dirty = false;
}

void ArenaPlanner::clear()
{
  allocs.clear();
  total = 0;
  dirty = true;
}

size_t ArenaPlanner::getByteSize() const
{
  if (dirty) throw std::logic_error("arena allocation plan is not committed");
  return total;
}

size_t ArenaPlanner::getAllocByteOffset(int allocID) const
{
  if (dirty) throw std::logic_error("arena allocation plan is not committed");
  return allocs.at(allocID).offset;
}

bool ArenaPlanner::validate() const
{
  for (size_t i = 0; i < allocs.size(); ++i)
    for (size_t j = i + 1; j < allocs.size(); ++j)
    {
      const Alloc &a = allocs[i], &b = allocs[j];
      const bool time = a.first <= b.last && b.first <= a.last;
      const bool space = a.offset < b.offset + b.size && b.offset < a.offset + a.size;
      if (time && space && a.size && b.size) return false;
    }
  return true;
}

} // namespace oidnb200
