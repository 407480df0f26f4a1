// Microbenchmark for the question the round-1 trace left open: what does ONE thread pay to
// issue a tcgen05.mma, as a function of N, of how many warps of the CTA issue concurrently, and of whether
// consecutive MMAs accumulate into the same TMEM tile or into different ones?
//
// One CTA per SM. `issuers` warps each own a private 128 x N fp32 accumulator region (or `tiles` of them, used
// round-robin) and a private pair of shared-memory operand tiles (contents are irrelevant: zeros). Each issuing
// warp's elected lane issues `count` MMAs (M=128, K=16, fp16, both operands from shared memory through 128B-swizzle
// descriptors), reading clock64 before the first and after the last issue ("issue" = the instruction stream is
// free again) and after waiting on the commit mbarrier ("done" = the tensor pipe has retired them).
//
//   probe_mma_issue <N> <issuers 1..4> <tiles 1..> <count>      prints cycles per MMA (issue / done), min over SMs
//
// Build: make -C oidn_b200/csrc probe_mma     Run on the GPU box: tools/bin/probe_mma_issue 96 2 1 64
#include "../oidn_b200/csrc/kernels/ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace oidnb200::ptx;

struct Result
{
  long long issue, done;
};

template <int KSTEPS, int SHIFT>
__device__ __forceinline__ void issue_group(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
  // one staged row's worth of MMAs as the conv kernels issue them: 3 taps x KSTEPS k-steps, offsets are immediates
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k)
      umma_f16(d, adesc + (uint64_t)(k * 2 + (SHIFT ? s * KSTEPS * 2 : 0)), bdesc + (uint64_t)(k * 2), idesc, (s | k) ? 1u : acc);
}

__global__ void __launch_bounds__(128, 1)
mma_issue_kernel(int N, int issuers, int tiles, int count, int row_bytes, int shift, Result* out)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  // layout: [0,64) barriers (one per warp), [64,68) tmem pointer, from 1024: per warp A tile (128 rows x 128 B) and
  // B tile (256 rows x 128 B)
  const uint32_t bar = sbase + 8 * warp;
  const uint32_t a_tile = sbase + 1024 + (uint32_t)warp * (16384 + 32768);
  const uint32_t b_tile = a_tile + 16384;
  for (int i = threadIdx.x; i < (int)((16384 + 32768) * 4 / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(sgen + 1024)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0)
  {
    for (int w = 0; w < 4; ++w) mbar_init(sbase + 8 * w, 1);
    fence_mbar_init();
  }
  if (warp == 0)
  {
    tmem_alloc(sbase + 64, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + 64);

  if (warp < issuers)
  {
    const bool leader = elect_one();
    // this warp's accumulator columns: 512 / issuers columns, cut into `tiles` regions of N columns
    const uint32_t cols = 512u / (uint32_t)issuers;
    const uint32_t t0 = tmem_base + (uint32_t)warp * cols;
    const uint32_t idesc = umma_idesc_f16((uint32_t)N);
    // row_bytes 128/64/32 = the swizzle mode of the layer's K chunk (64/32/16 channels); shift: the A operand starts
    // 0, 1, 2 rows into the tile in turn (the conv kernels' horizontal taps are such shifted views of one staged row)
    const uint64_t adesc = umma_desc(a_tile, (uint32_t)row_bytes, 0), bdesc = umma_desc(b_tile, (uint32_t)row_bytes, 0);
    const int ksteps = row_bytes / 32;
    long long c0 = 0, c1 = 0, c2 = 0;
    __syncwarp();
    if (leader)
    {
      c0 = clock64();
      const int groups = count / (3 * ksteps);
      for (int g = 0; g < groups; ++g)
      {
        const uint32_t d = t0 + (uint32_t)(g % tiles) * (uint32_t)N;
        const uint32_t acc = g >= tiles ? 1u : 0u;
        if (ksteps == 4) { if (shift) issue_group<4, 1>(d, adesc, bdesc, idesc, acc); else issue_group<4, 0>(d, adesc, bdesc, idesc, acc); }
        else if (ksteps == 2) { if (shift) issue_group<2, 1>(d, adesc, bdesc, idesc, acc); else issue_group<2, 0>(d, adesc, bdesc, idesc, acc); }
        else { if (shift) issue_group<1, 1>(d, adesc, bdesc, idesc, acc); else issue_group<1, 0>(d, adesc, bdesc, idesc, acc); }
      }
      c1 = clock64();
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0, 1);
    c2 = clock64();
    if (leader)
    {
      out[blockIdx.x * 4 + warp].issue = c1 - c0;
      out[blockIdx.x * 4 + warp].done = c2 - c0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
  {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main(int argc, char** argv)
{
  if (argc < 5)
  {
    printf("usage: probe_mma_issue <N 16..256, multiple of 16> <issuers 1..4> <tiles >= 1> <count> [row_bytes 128|64|32] [shift 0|1]\n");
    return 1;
  }
  const int N = atoi(argv[1]), issuers = atoi(argv[2]), tiles = atoi(argv[3]);
  int count = atoi(argv[4]);
  const int row_bytes = argc > 5 ? atoi(argv[5]) : 128, shift = argc > 6 ? atoi(argv[6]) : 0;
  if (N < 16 || N > 256 || N % 16 || issuers < 1 || issuers > 4 || tiles < 1 || tiles * N > 512 / issuers || count < 1)
  {
    printf("bad arguments (tiles * N must fit 512 / issuers TMEM columns)\n");
    return 1;
  }
  count = count / (3 * (row_bytes / 32)) * (3 * (row_bytes / 32));   // whole rows of 3 taps x k-steps
  if (count < 1) { printf("count too small\n"); return 1; }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = 1024 + 1024 + 4 * (16384 + 32768);
  cudaFuncSetAttribute(mma_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  Result* d = nullptr;
  cudaMalloc(&d, sizeof(Result) * sms * 4);
  for (int rep = 0; rep < 3; ++rep)
  {
    cudaMemset(d, 0, sizeof(Result) * sms * 4);
    mma_issue_kernel<<<sms, 128, smem>>>(N, issuers, tiles, count, row_bytes, shift, d);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
    {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      return 2;
    }
  }
  std::vector<Result> h(sms * 4);
  cudaMemcpy(h.data(), d, sizeof(Result) * sms * 4, cudaMemcpyDeviceToHost);
  long long best_issue = -1, best_done = -1, worst_done = 0;
  for (int b = 0; b < sms; ++b)
    for (int w = 0; w < issuers; ++w)
    {
      const Result& r = h[b * 4 + w];
      if (best_issue < 0 || r.issue < best_issue) best_issue = r.issue;
      if (best_done < 0 || r.done < best_done) best_done = r.done;
      if (r.done > worst_done) worst_done = r.done;
    }
  const double floor_cycles = 128.0 * N / 256.0;   // tensor-pipe floor per MMA (B300_MICROARCH.md), one issuer
  printf("N=%d issuers=%d tiles=%d count=%d rowB=%d shift=%d | per MMA: issue %.1f cycles, done %.1f (min) .. %.1f (max) cycles | pipe floor %.1f x %d issuers = %.1f\n",
         N, issuers, tiles, count, row_bytes, shift, (double)best_issue / count, (double)best_done / count, (double)worst_done / count,
         floor_cycles, issuers, floor_cycles * issuers);
  cudaFree(d);
  return 0;
}
