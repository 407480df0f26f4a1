// Common host-side vocabulary of the B200 backend: error codes, exceptions, pixel formats, image
// descriptors and integer helpers. Values of the enums are the public API's
// (include/OpenImageDenoise/oidn.h:93-102, :239-254, :376-383 in the reference tree) so a device
// module can pass them through unchanged.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>

namespace oidnb200 {

enum class Error : int
{
  None = 0, Unknown = 1, InvalidArgument = 2, InvalidOperation = 3, OutOfMemory = 4,
  UnsupportedHardware = 5, Cancelled = 6
};

class Exception : public std::exception
{
public:
  Exception(Error code, const std::string& msg) : errCode(code), message(msg) {}
  Error code() const noexcept { return errCode; }
  const char* what() const noexcept override { return message.c_str(); }

private:
  Error errCode;
  std::string message;
};

enum class Format : int
{
  Undefined = 0, Float = 1, Float2 = 2, Float3 = 3, Float4 = 4, Half = 257, Half2 = 258, Half3 = 259, Half4 = 260
};

enum class Quality : int { Default = 0, Fast = 4, Balanced = 5, High = 6 };
enum class Storage : int { Undefined = 0, Host = 1, Device = 2, Managed = 3 };
enum class SyncMode { Blocking, Async };
enum class Activation { None, ReLU };
enum class PostOp { None, Pool, Upsample };              // core/conv.h:18-23
enum class TransferType : int { Linear = 0, SRGB = 1, PU = 2, Log = 3 }; // core/color.h:12-18

inline int formatChannels(Format f)
{
  const int v = static_cast<int>(f);
  if (v >= 1 && v <= 4) return v;
  if (v >= 257 && v <= 260) return v - 256;
  return 0;
}
inline bool formatIsHalf(Format f) { return static_cast<int>(f) >= 257; }
inline size_t formatBytes(Format f) { return (size_t)formatChannels(f) * (formatIsHalf(f) ? 2 : 4); }

// A user image (core/image.h:14-120): borrowed pointer + strides. `ptr == nullptr` = not set.
struct Image
{
  void* ptr = nullptr;
  Format format = Format::Undefined;
  int W = 0, H = 0;
  size_t pixelStride = 0, rowStride = 0;

  explicit operator bool() const { return ptr != nullptr; }
  int C() const { return formatChannels(format); }
  const uint8_t* begin() const { return static_cast<const uint8_t*>(ptr); }
  const uint8_t* end() const
  {
    if (!ptr || W == 0 || H == 0) return begin();
    return begin() + (size_t)(H - 1) * rowStride + (size_t)(W - 1) * pixelStride + formatBytes(format);
  }
  // core/image.h "overlaps": byte ranges intersect
  bool overlaps(const Image& o) const { return ptr && o.ptr && begin() < o.end() && o.begin() < end(); }
};

constexpr size_t memoryAlignment = 256; // common/platform.h:263

template <typename T> constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T> constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }
// smallest value >= a that is congruent to c modulo b (common/platform.h:208)
template <typename T> constexpr T round_up(T a, T b, T c) { return ceil_div(a - c, b) * b + c; }
template <typename T> constexpr T gcd_(T a, T b) { return b == 0 ? a : gcd_(b, a % b); }
template <typename T> constexpr T lcm_(T a, T b) { return a / gcd_(a, b) * b; }
template <typename T> constexpr T clamp_(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }

} // namespace oidnb200
