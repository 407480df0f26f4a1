// Harness for the fused conv pair (kernels/conv_pair_tc.cu): B(A(x)) as one launch against the two single-conv
// launches on the same buffers -- must be bit-identical -- and the time of both.
//   probe_pair <H> <W> <C1A> <CoutA> <CoutB> <poolB> [iters]
#include "../include/oidn_b200_kernels.h"
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static uint32_t lcg = 12345u;
static float frand() { lcg = lcg * 1664525u + 1013904223u; return (lcg >> 8) * (1.0f / 16777216.0f); }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define AB(x) do { if (x) { printf("ABI error at line %d: %s\n", __LINE__, oidnb200_last_error()); return 3; } } while (0)

static int make_conv(const oidnb200_conv_desc& d, int I, int O, oidnb200_conv** conv, void** dw, void** db)
{
  AB(oidnb200_conv_create(&d, conv));
  std::vector<uint16_t> w((size_t)O * I * 9), b(O);
  const float ws = sqrtf(2.f / (9.f * I));
  for (auto& v : w) { __half t = __float2half((frand() * 2.f - 1.f) * 1.7f * ws); v = *(uint16_t*)&t; }
  for (auto& v : b) { __half t = __float2half(frand() * 0.1f); v = *(uint16_t*)&t; }
  std::vector<uint8_t> pw(oidnb200_conv_weight_bytes(*conv)), pb(oidnb200_conv_bias_bytes(*conv));
  AB(oidnb200_conv_pack_weights(*conv, w.data(), O, I, 0, pw.data()));
  AB(oidnb200_conv_pack_bias(*conv, b.data(), O, pb.data()));
  CK(cudaMalloc(dw, pw.size())); CK(cudaMalloc(db, pb.size()));
  CK(cudaMemcpy(*dw, pw.data(), pw.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(*db, pb.data(), pb.size(), cudaMemcpyHostToDevice));
  return 0;
}

int main(int argc, char** argv)
{
  if (argc < 7) { printf("usage: probe_pair H W C1A CoutA CoutB poolB [iters]\n"); return 1; }
  const int H = atoi(argv[1]), W = atoi(argv[2]), C1 = atoi(argv[3]), CA = atoi(argv[4]), CB = atoi(argv[5]), pool = atoi(argv[6]);
  const int iters = argc > 7 ? atoi(argv[7]) : 0;
  oidnb200_conv_desc da{}, db{};
  da.H = H; da.W = W; da.C1 = C1; da.C2 = 0; da.Cout = CA; da.relu = 1;
  db.H = H; db.W = W; db.C1 = CA; db.C2 = 0; db.Cout = CB; db.relu = 1; db.post_op = pool;
  oidnb200_conv *a = nullptr, *b = nullptr;
  void *wa, *ba, *wb, *bb;
  const int IA = C1 > 9 ? C1 - 3 : C1, OA = CA, OB = CB > 3 ? CB - 1 : CB;
  if (int rc = make_conv(da, IA, OA, &a, &wa, &ba)) return rc;
  if (int rc = make_conv(db, CA, OB, &b, &wb, &bb)) return rc;

  const size_t n_in = (size_t)H * W * C1, n_mid = (size_t)H * W * CA;
  size_t n_out = (size_t)H * W * CB; if (pool) n_out /= 4;
  std::vector<__half> hin(n_in);
  for (size_t i = 0; i < n_in; ++i) hin[i] = __float2half((int)(i % C1) < IA ? frand() : 0.f);
  void *din, *dmid, *dref, *dout;
  CK(cudaMalloc(&din, n_in * 2)); CK(cudaMalloc(&dmid, n_mid * 2)); CK(cudaMalloc(&dref, n_out * 2)); CK(cudaMalloc(&dout, n_out * 2));
  CK(cudaMemcpy(din, hin.data(), n_in * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dref, 0, n_out * 2)); CK(cudaMemset(dout, 0xFF, n_out * 2));

  // two launches -> dref
  AB(oidnb200_conv_bind(a, din, nullptr, wa, ba, dmid));
  AB(oidnb200_conv_bind(b, dmid, nullptr, wb, bb, dref));
  AB(oidnb200_conv_launch(a, 0)); AB(oidnb200_conv_launch(b, 0));
  CK(cudaDeviceSynchronize());
  // one launch -> dout
  oidnb200_conv_pair* pr = nullptr;
  if (oidnb200_conv_pair_create(a, b, &pr)) { printf("pair not supported: %s\n", oidnb200_last_error()); return 4; }
  oidnb200_conv_info info; oidnb200_conv_pair_get_info(pr, &info);
  printf("cfg H=%d W=%d %d->%d->%d pool=%d | grid=%d smem=%d streams=%d NA=%d NM=%d RA=%d RB=%d RC=%d strips=%d\n", H, W, C1, CA, CB, pool,
         info.grid, info.smem_bytes, info.nstreams, info.nstages, info.ngroups, info.nchunks, info.ring_slots, info.rows_per_item, info.nstrips);
  AB(oidnb200_conv_bind(b, dmid, nullptr, wb, bb, dout));
  AB(oidnb200_conv_pair_bind(pr));
  AB(oidnb200_conv_pair_launch(pr, 0));
  CK(cudaDeviceSynchronize());

  std::vector<uint16_t> o(n_out), r(n_out);
  CK(cudaMemcpy(o.data(), dout, n_out * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r.data(), dref, n_out * 2, cudaMemcpyDeviceToHost));
  size_t bad = 0, first = (size_t)-1;
  for (size_t i = 0; i < n_out; ++i) if (o[i] != r[i]) { if (!bad) first = i; ++bad; }
  printf("RESULT %s bad=%zu/%zu", bad ? "FAIL" : "PASS (bit-identical)", bad, n_out);
  if (bad)
  {
    const int Wd = pool ? W / 2 : W; const size_t px = first / CB;
    printf(" first bad: y=%zu x=%zu c=%zu got=%g ref=%g", px / Wd, px % Wd, first % CB,
           __half2float(*(__half*)&o[first]), __half2float(*(__half*)&r[first]));
    // how the errors are distributed
    size_t byx[8] = {0}; for (size_t i = 0; i < n_out; ++i) if (o[i] != r[i]) { const size_t x = (i / CB) % Wd; byx[(x * 8) / Wd]++; }
    printf(" | by x-octile:"); for (int k = 0; k < 8; ++k) printf(" %zu", byx[k]);
  }
  printf("\n");
  if (iters > 0 && !bad && getenv("PROBE_TRACE"))
  {
    unsigned long long* dtr; CK(cudaMalloc(&dtr, 352 * 8)); CK(cudaMemset(dtr, 0, 352 * 8));
    oidnb200_conv_set_trace(b, dtr);
    oidnb200_conv_pair_launch(pr, 0); CK(cudaDeviceSynchronize());
    unsigned long long tr[352]; CK(cudaMemcpy(tr, dtr, sizeof(tr), cudaMemcpyDeviceToHost));
    const char* roles[22] = {"TMA0", "MMAA0", "TMA1", "MMAA1", "AEPI0.0", "AEPI0.1", "AEPI0.2", "AEPI0.3", "AEPI1.0", "AEPI1.1", "AEPI1.2", "AEPI1.3",
                             "BEPI0.0", "BEPI0.1", "BEPI0.2", "BEPI0.3", "BEPI1.0", "BEPI1.1", "BEPI1.2", "BEPI1.3", "MMAB0", "MMAB1"};
    printf("TRACE (cycles/CTA; waits: 1 A-stage-empty 3 accA-empty 4 A-stage-full 5 accA-full 6 mid-full 7 accB-empty 8 mid-empty 9/10 accB-full; 11 = A issue, 15 = B issue)\n");
    for (int w = 0; w < 22; ++w)
    {
      if ((w >= 4 && w < 20 && (w & 3)) || !tr[w * 16]) continue;
      printf("  %-8s total %9.0f |", roles[w], (double)tr[w * 16] / info.grid);
      for (int t = 1; t < 16; ++t) if (tr[w * 16 + t]) printf(" t%d %5.1f%%", t, 100.0 * tr[w * 16 + t] / (double)tr[w * 16]);
      printf("\n");
    }
    oidnb200_conv_set_trace(b, nullptr);
  }
  if (iters > 0 && !bad)
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms2, ms1;
    for (int i = 0; i < 3; ++i) { oidnb200_conv_launch(a, 0); oidnb200_conv_launch(b, 0); }
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) { oidnb200_conv_launch(a, 0); oidnb200_conv_launch(b, 0); }
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms2, e0, e1); ms2 /= iters;
    for (int i = 0; i < 3; ++i) oidnb200_conv_pair_launch(pr, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) oidnb200_conv_pair_launch(pr, 0);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms1, e0, e1); ms1 /= iters;
    printf("TIME two launches %.4f ms, fused pair %.4f ms (%.2fx)\n", ms2, ms1, ms2 / ms1);
  }
  return bad ? 10 : 0;
}
