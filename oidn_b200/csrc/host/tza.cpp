#include "tza.hpp"
#include <cstdint>
#include <cstring>

namespace oidnb200 {

namespace {
struct Cursor
{
  const uint8_t* base;
  size_t size, pos;
  void need(size_t n) const
  {
    if (n > size - pos) throw Exception(Error::InvalidOperation, "invalid or corrupted weights blob");
  }
  template <typename T> T read()
  {
    need(sizeof(T));
    T v;
    memcpy(&v, base + pos, sizeof(T));
    pos += sizeof(T);
    return v;
  }
  std::string readString(size_t n)
  {
    need(n);
    std::string s(reinterpret_cast<const char*>(base + pos), n);
    pos += n;
    return s;
  }
};
} // namespace

std::shared_ptr<TensorMap> parseTZA(const void* buffer, size_t size)
{
  if (!buffer || size < 12) throw Exception(Error::InvalidOperation, "invalid or corrupted weights blob");
  Cursor c{static_cast<const uint8_t*>(buffer), size, 0};
  if (c.read<uint16_t>() != 0x41D7) throw Exception(Error::InvalidOperation, "invalid or corrupted weights blob");
  const uint8_t major = c.read<uint8_t>();
  c.read<uint8_t>(); // minor
  if (major != 2) throw Exception(Error::InvalidOperation, "unsupported weights blob version");
  const uint64_t table = c.read<uint64_t>();
  if (table > size) throw Exception(Error::InvalidOperation, "invalid or corrupted weights blob");
  c.pos = (size_t)table;

  auto map = std::make_shared<TensorMap>();
  const uint32_t n = c.read<uint32_t>();
  for (uint32_t i = 0; i < n; ++i)
  {
    ConstTensor t;
    const uint16_t nameLen = c.read<uint16_t>();
    const std::string name = c.readString(nameLen);
    const uint8_t ndims = c.read<uint8_t>();
    // only "x" (rank 1) and "oihw" (rank 4) tensors exist; dims become ints downstream and the element
    // count must not wrap (a crafted blob could otherwise pass the bounds check with a wrapped byte size)
    if (ndims != 1 && ndims != 4) throw Exception(Error::InvalidOperation, "invalid tensor layout");
    size_t count = 1;
    for (int d = 0; d < ndims; ++d)
    {
      const uint32_t v = c.read<uint32_t>();
      if (v > (uint32_t)INT32_MAX || __builtin_mul_overflow(count, (size_t)v, &count))
        throw Exception(Error::InvalidOperation, "invalid or corrupted weights blob");
      t.dims.push_back((int)v);
    }
    t.layout = c.readString(ndims);
    if (!((ndims == 1 && t.layout == "x") || (ndims == 4 && t.layout == "oihw")))
      throw Exception(Error::InvalidOperation, "invalid tensor layout");
    t.dtype = (char)c.read<uint8_t>();
    if (t.dtype != 'f' && t.dtype != 'h') throw Exception(Error::InvalidOperation, "invalid tensor data type");
    const uint64_t off = c.read<uint64_t>();
    size_t bytes = 0;
    if (__builtin_mul_overflow(count, (size_t)(t.dtype == 'h' ? 2 : 4), &bytes))
      throw Exception(Error::InvalidOperation, "invalid or corrupted weights blob");
    if (off > size || bytes > size - off) throw Exception(Error::InvalidOperation, "invalid or corrupted weights blob");
    t.data = c.base + off;
    (*map)[name] = t;
  }
  return map;
}

} // namespace oidnb200
