// TZA weight container reader (format: training/tza.py:12-108; reader contract core/tza.cpp:27-103).
#pragma once
#include "base.hpp"
#include <map>
#include <vector>

namespace oidnb200 {

// A constant tensor aliasing the user's blob (zero copy, like core/tza.cpp:96-99).
struct ConstTensor
{
  std::vector<int> dims;   // "x": [n]; "oihw": [o,i,h,w]
  std::string layout;      // "x" | "oihw"
  char dtype = 'h';        // 'h' fp16 | 'f' fp32
  const void* data = nullptr;
  size_t count() const { size_t n = 1; for (int d : dims) n *= (size_t)d; return n; }
};

using TensorMap = std::map<std::string, ConstTensor>;

// Throws Exception(InvalidOperation, ...) with the reference's messages on a malformed blob.
std::shared_ptr<TensorMap> parseTZA(const void* buffer, size_t size);

} // namespace oidnb200
