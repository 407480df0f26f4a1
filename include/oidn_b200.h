/* oidn_b200 filter-level C ABI: devices, buffers and the RT / RTLightmap filters.
 *
 * Every entry point mirrors one function of the reference's public C99 API
 * (include/OpenImageDenoise/oidn.h in the reference tree; the line each one stands in for is cited)
 * with the same argument meaning, parameter names ("hdr", "srgb", "cleanAux", "directional",
 * "quality", "maxMemoryMB", "inputScale", "weights", "tileAlignment", "tileOverlap"), enum values
 * (OIDNFormat / OIDNQuality / OIDNStorage / OIDNError numbering) and error behaviour: functions
 * never throw, the first error is stored per device and fetched with oidnb200GetDeviceError.
 * Plain pointers and sizes only. The kernel-level ABI one layer down is oidn_b200_kernels.h.
 */
#ifndef OIDN_B200_H
#define OIDN_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OIDNB200_API __attribute__((visibility("default")))

typedef struct oidnb200_device_t* OIDNB200Device;
typedef struct oidnb200_buffer_t* OIDNB200Buffer;
typedef struct oidnb200_filter_t* OIDNB200Filter;

/* OIDNError (oidn.h:93-102) */
enum
{
  OIDNB200_ERROR_NONE = 0, OIDNB200_ERROR_UNKNOWN = 1, OIDNB200_ERROR_INVALID_ARGUMENT = 2,
  OIDNB200_ERROR_INVALID_OPERATION = 3, OIDNB200_ERROR_OUT_OF_MEMORY = 4,
  OIDNB200_ERROR_UNSUPPORTED_HARDWARE = 5, OIDNB200_ERROR_CANCELLED = 6
};
/* OIDNQuality (oidn.h:376-383) */
enum { OIDNB200_QUALITY_DEFAULT = 0, OIDNB200_QUALITY_FAST = 4, OIDNB200_QUALITY_BALANCED = 5, OIDNB200_QUALITY_HIGH = 6 };
/* OIDNStorage (oidn.h:257-270) */
enum { OIDNB200_STORAGE_UNDEFINED = 0, OIDNB200_STORAGE_HOST = 1, OIDNB200_STORAGE_DEVICE = 2, OIDNB200_STORAGE_MANAGED = 3 };

typedef bool (*OIDNB200ProgressMonitorFunction)(void* userPtr, double n); /* oidn.h:386 */

/* ---- device ------------------------------------------------------------------------------- */
/* oidnGetNumPhysicalDevices restricted to B200-class GPUs (oidn.h:55). */
OIDNB200_API int oidnb200GetNumPhysicalDevices(void);
/* oidnNewCUDADevice(deviceIDs, streams, numPairs) (oidn.h:152-153). streams may be NULL or hold
 * NULL entries (the engine creates its own stream). numPairs > 1 makes a multi-engine device whose
 * tiles are dealt round-robin to the (GPU, stream) pairs (core/unet_filter.cpp:219); distinct GPUs must
 * be peer accessible; the same GPU may appear in several pairs. A frame that does not live in the memory
 * of the GPU that runs a tile (pinned host memory, another GPU) is staged tile by tile with copy engines
 * (device parameter "staging": -1 auto (default), 0 never, 1 always; see oidn_b200/csrc/host/filter.hpp).
 * Returns NULL on error (fetch it with oidnb200GetDeviceError(NULL, ...)). */
OIDNB200_API OIDNB200Device oidnb200NewCUDADevice(const int* deviceIDs, void* const* streams, int numPairs);
/* oidnNewDevice(OIDN_DEVICE_TYPE_CUDA) (oidn.h:128): GPU 0, own stream. */
OIDNB200_API OIDNB200Device oidnb200NewDevice(void);
OIDNB200_API void oidnb200RetainDevice(OIDNB200Device device);                 /* oidn.h:165 */
OIDNB200_API void oidnb200ReleaseDevice(OIDNB200Device device);                /* oidn.h:168 */
OIDNB200_API void oidnb200SetDeviceInt(OIDNB200Device device, const char* name, int value);   /* oidn.h:180 */
OIDNB200_API int oidnb200GetDeviceInt(OIDNB200Device device, const char* name);               /* oidn.h:204 */
/* backend parameter: "weightsDir" = directory holding the reference's weights/<name>.tza files */
OIDNB200_API void oidnb200SetDeviceString(OIDNB200Device device, const char* name, const char* value);
OIDNB200_API void oidnb200CommitDevice(OIDNB200Device device);                 /* oidn.h:229 */
OIDNB200_API void oidnb200SyncDevice(OIDNB200Device device);                   /* oidn.h:232 */
/* oidnGetDeviceError (oidn.h:225): returns and clears the first stored error. device may be NULL. */
OIDNB200_API int oidnb200GetDeviceError(OIDNB200Device device, const char** outMessage);

/* ---- buffers (oidn.h:317-369) ------------------------------------------------------------- */
OIDNB200_API OIDNB200Buffer oidnb200NewBufferWithStorage(OIDNB200Device device, size_t byteSize, int storage);
OIDNB200_API OIDNB200Buffer oidnb200NewBuffer(OIDNB200Device device, size_t byteSize); /* device storage */
OIDNB200_API void* oidnb200GetBufferData(OIDNB200Buffer buffer);
OIDNB200_API size_t oidnb200GetBufferSize(OIDNB200Buffer buffer);
OIDNB200_API void oidnb200ReadBuffer(OIDNB200Buffer buffer, size_t byteOffset, size_t byteSize, void* dstHostPtr);
OIDNB200_API void oidnb200WriteBuffer(OIDNB200Buffer buffer, size_t byteOffset, size_t byteSize, const void* srcHostPtr);
OIDNB200_API void oidnb200ReadBufferAsync(OIDNB200Buffer buffer, size_t byteOffset, size_t byteSize, void* dstHostPtr);
OIDNB200_API void oidnb200WriteBufferAsync(OIDNB200Buffer buffer, size_t byteOffset, size_t byteSize, const void* srcHostPtr);
OIDNB200_API void oidnb200ReleaseBuffer(OIDNB200Buffer buffer);
/* Copy-engine transfer of an image rectangle (rows of widthBytes, row pitches in bytes) in the
 * device's stream order, between any two of: this GPU's memory, a peer GPU's memory (imported
 * buffer / peer-enabled pointer: NVLink), pinned host memory. Used to stage a tile of a frame that
 * lives on another GPU or on the host without occupying SMs (no counterpart in oidn.h, whose
 * buffers only copy linearly: oidnReadBufferAsync / oidnWriteBufferAsync, oidn.h:352-369). */
OIDNB200_API void oidnb200CopyRectAsync(OIDNB200Device device, void* dst, size_t dstPitch, const void* src, size_t srcPitch,
                                        size_t widthBytes, size_t height);
/* Cross-process sharing of a device buffer on one node (the role oidnNewSharedBufferFromFD,
 * oidn.h:326-329, plays for external memory): the owner exports a 64-byte CUDA IPC handle, a peer
 * process imports it as a buffer it may read and write over NVLink. Device-storage buffers only. */
OIDNB200_API void oidnb200GetBufferIpcHandle(OIDNB200Buffer buffer, void* outHandle64);
OIDNB200_API OIDNB200Buffer oidnb200NewSharedBufferFromIpcHandle(OIDNB200Device device, const void* handle64, size_t byteSize);

/* oidnNewSharedBufferFromFD (oidn.h:326-329; devices/cuda/cuda_external_buffer.cpp:9-23, :66-84): imports device
 * memory from a POSIX file descriptor. fdType uses OIDNExternalMemoryTypeFlag's numbering (oidn.h:296-311); like the
 * reference CUDA device only OPAQUE_FD is supported, DMA_BUF is rejected with InvalidArgument. The fd is tried as an
 * external-memory object of another API first (cudaImportExternalMemory, what a Vulkan / OpenGL renderer exports),
 * then as a CUDA allocation exported with cuMemExportToShareableHandle (another CUDA process, or
 * oidnb200GetBufferFD below). On success the buffer owns the fd. */
enum { OIDNB200_EXTERNAL_MEMORY_TYPE_FLAG_NONE = 0, OIDNB200_EXTERNAL_MEMORY_TYPE_FLAG_OPAQUE_FD = 1 << 0,
       OIDNB200_EXTERNAL_MEMORY_TYPE_FLAG_DMA_BUF = 1 << 1 };
OIDNB200_API OIDNB200Buffer oidnb200NewSharedBufferFromFD(OIDNB200Device device, int fdType, int fd, size_t byteSize);
/* Producer side of the same interop (no counterpart in oidn.h, where the renderer's API exports): a device buffer
 * whose memory can be handed to another process or API as an opaque fd. oidnb200GetBufferFD returns a new fd per
 * call (the caller owns it; -1 on error). */
OIDNB200_API OIDNB200Buffer oidnb200NewExportableBuffer(OIDNB200Device device, size_t byteSize);
OIDNB200_API int oidnb200GetBufferFD(OIDNB200Buffer buffer);

/* ---- filters ------------------------------------------------------------------------------ */
OIDNB200_API OIDNB200Filter oidnb200NewFilter(OIDNB200Device device, const char* type); /* "RT" | "RTLightmap", oidn.h:392 */
OIDNB200_API void oidnb200RetainFilter(OIDNB200Filter filter);
OIDNB200_API void oidnb200ReleaseFilter(OIDNB200Filter filter);
/* oidnSetFilterImage (oidn.h:402-407) */
OIDNB200_API void oidnb200SetFilterImage(OIDNB200Filter filter, const char* name, OIDNB200Buffer buffer, int format,
                                         size_t width, size_t height, size_t byteOffset,
                                         size_t pixelByteStride, size_t rowByteStride);
/* oidnSetSharedFilterImage (oidn.h:410-414): borrowed pointer; strides 0 = tightly packed */
OIDNB200_API void oidnb200SetSharedFilterImage(OIDNB200Filter filter, const char* name, void* devPtr, int format,
                                               size_t width, size_t height, size_t byteOffset,
                                               size_t pixelByteStride, size_t rowByteStride);
OIDNB200_API void oidnb200UnsetFilterImage(OIDNB200Filter filter, const char* name);          /* oidn.h:417 */
OIDNB200_API void oidnb200SetSharedFilterData(OIDNB200Filter filter, const char* name, void* hostPtr, size_t byteSize); /* oidn.h:426 */
OIDNB200_API void oidnb200UpdateFilterData(OIDNB200Filter filter, const char* name);          /* oidn.h:430 */
OIDNB200_API void oidnb200UnsetFilterData(OIDNB200Filter filter, const char* name);           /* oidn.h:433 */
OIDNB200_API void oidnb200SetFilterBool(OIDNB200Filter filter, const char* name, bool value); /* oidn.h:442 */
OIDNB200_API bool oidnb200GetFilterBool(OIDNB200Filter filter, const char* name);
OIDNB200_API void oidnb200SetFilterInt(OIDNB200Filter filter, const char* name, int value);
OIDNB200_API int oidnb200GetFilterInt(OIDNB200Filter filter, const char* name);
OIDNB200_API void oidnb200SetFilterFloat(OIDNB200Filter filter, const char* name, float value);
OIDNB200_API float oidnb200GetFilterFloat(OIDNB200Filter filter, const char* name);
OIDNB200_API void oidnb200SetFilterProgressMonitorFunction(OIDNB200Filter filter, OIDNB200ProgressMonitorFunction func, void* userPtr);
OIDNB200_API void oidnb200CommitFilter(OIDNB200Filter filter);        /* oidn.h:501 */
OIDNB200_API void oidnb200ExecuteFilter(OIDNB200Filter filter);       /* oidn.h:504 */
OIDNB200_API void oidnb200ExecuteFilterAsync(OIDNB200Filter filter);  /* oidn.h:507 */

/* Introspection used by the tests and the benchmark (tile grid chosen by the scheduler). */
typedef struct oidnb200_filter_info
{
  int tileH, tileW, tileCountH, tileCountW, tileOverlap, tileAlignment, largeModel, numOps;
  int staged; /* 1: the last frame ran through the tile-staging pipeline (device parameter "staging") */
  size_t memoryBytes;
} oidnb200_filter_info;
OIDNB200_API void oidnb200GetFilterInfo(OIDNB200Filter filter, oidnb200_filter_info* info);

/* Per-op device times (ms, summed over tiles and frames) since the last reset; enabled by the device
 * parameter "profile" = 1. kind: 0 conv, 1 input process, 2 output process. Returns the op count. */
typedef struct oidnb200_op_time
{
  char name[32];
  int kind, launches;
  double ms;
} oidnb200_op_time;
OIDNB200_API int oidnb200GetFilterProfile(OIDNB200Filter filter, oidnb200_op_time* out, int maxOps);
OIDNB200_API void oidnb200ResetFilterProfile(OIDNB200Filter filter);

/* The tile planner alone (no GPU needed): core/unet_filter.cpp:254-335. */
typedef struct oidnb200_tile_plan
{
  int H, W, tileH, tileW, tilePadH, tilePadW, tileCountH, tileCountW, tileAlignment, tileOverlap;
} oidnb200_tile_plan;
OIDNB200_API void oidnb200PlanTiles(int H, int W, int largeModel, int deviceMinAlignment, int numEngines,
                                    long maxTilePixels, oidnb200_tile_plan* plan);
/* This backend's default search (device parameter "tilePolicy" = 1): same geometry rules, but the
 * grid whose tile count is divisible by numUnits (engines x shards) and recomputes the fewest pixels. */
OIDNB200_API void oidnb200PlanTilesMinOverlap(int H, int W, int largeModel, int deviceMinAlignment, int numUnits,
                                              long maxTilePixels, oidnb200_tile_plan* plan);
/* "tilePolicy" = 2 (opt-in): as above, with tile widths costed in whole 128-pixel conv strips per UNet
 * level (a 2064-wide tile is 17 strips, not 16.1): 8K over 8 units becomes 2x4 tiles of 3936x1232
 * instead of 4x2 of 2064x2256. */
OIDNB200_API void oidnb200PlanTilesStripAware(int H, int W, int largeModel, int deviceMinAlignment, int numUnits,
                                              long maxTilePixels, oidnb200_tile_plan* plan);
/* Tile rectangles of a plan, 12 ints each (hSrc,wSrc,hBuf,wBuf,H1,W1,hOutBuf,wOutBuf,hDst,wDst,H2,W2);
 * returns the tile count; writes at most maxTiles tiles. */
OIDNB200_API int oidnb200EnumerateTiles(const oidnb200_tile_plan* plan, int* out, int maxTiles);
/* TZA parser check (core/tza.cpp:27-103): returns the number of tensors or -OIDNError. */
OIDNB200_API int oidnb200ParseTZA(const void* blob, size_t size, const char** outMessage);
/* Arena planner self-check on a synthetic lifetime list (tests): sizes/first/last arrays of n
 * allocations; writes offsets; returns total bytes. */
OIDNB200_API size_t oidnb200PlanArena(int n, const size_t* sizes, const int* firstOp, const int* lastOp, size_t* offsets);

#ifdef __cplusplus
}
#endif
#endif /* OIDN_B200_H */
