// Hardware probe / unit harness for the tcgen05 conv: runs one configuration per process (a trap
// poisons the context), compares against the SIMT witness and prints timing.
//   probe_conv <H> <W> <C1> <C2> <Cout> <post_op> <src1_up> <shift_mode> [iters]
#include "../include/oidn_b200_kernels.h"
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

static uint32_t lcg = 12345u;
static float frand() { lcg = lcg * 1664525u + 1013904223u; return (lcg >> 8) * (1.0f / 16777216.0f); }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv)
{
  if (argc < 9) { printf("usage\n"); return 1; }
  oidnb200_conv_desc d{};
  d.H = atoi(argv[1]); d.W = atoi(argv[2]); d.C1 = atoi(argv[3]); d.C2 = atoi(argv[4]);
  d.Cout = atoi(argv[5]); d.post_op = atoi(argv[6]); d.src1_upsampled = atoi(argv[7]);
  d.shift_mode = atoi(argv[8]); d.relu = 1;
  const int iters = argc > 9 ? atoi(argv[9]) : 0;

  oidnb200_conv* conv = nullptr;
  if (oidnb200_conv_create(&d, &conv)) { printf("create failed: %s\n", oidnb200_last_error()); return 3; }
  oidnb200_conv_info info; oidnb200_conv_get_info(conv, &info);
  printf("cfg H=%d W=%d C1=%d C2=%d Cout=%d post=%d up=%d mode=%d | grid=%d smem=%d groups=%d CoutG=%d chunks=%d stages=%d R=%d RC=%d streams=%d\n",
         d.H, d.W, d.C1, d.C2, d.Cout, d.post_op, d.src1_upsampled, d.shift_mode, info.grid, info.smem_bytes,
         info.ngroups, info.cout_group, info.nchunks, info.nstages, info.ring_slots, info.rows_per_item, info.nstreams);

  const int H1 = d.src1_upsampled ? d.H / 2 : d.H, W1 = d.src1_upsampled ? d.W / 2 : d.W;
  const size_t n1 = (size_t)H1 * W1 * d.C1, n2 = (size_t)d.H * d.W * d.C2;
  size_t nout = (size_t)d.H * d.W * d.Cout;
  if (d.post_op == 1) nout /= 4;
  if (d.post_op == 2) nout *= 4;
  std::vector<__half> h1(n1), h2(n2 ? n2 : 1);
  for (auto& v : h1) v = __float2half(frand());
  for (auto& v : h2) v = __float2half(frand());
  const int I1 = d.C1 - 3 > 0 ? d.C1 - 3 : d.C1, I2 = d.C2 ? d.C2 - 5 : 0, O = d.Cout - 1 > 0 ? d.Cout - 1 : d.Cout;
  // zero the padded channels of the sources like the real pipeline does
  for (size_t i = 0; i < n1 / d.C1; ++i) for (int c = I1; c < d.C1; ++c) h1[i * d.C1 + c] = __float2half(0.f);
  for (size_t i = 0; d.C2 && i < n2 / d.C2; ++i) for (int c = I2; c < d.C2; ++c) h2[i * d.C2 + c] = __float2half(0.f);
  std::vector<uint16_t> w((size_t)O * (I1 + I2) * 9), b(O);
  const float ws = sqrtf(2.f / (9.f * (I1 + I2)));
  for (auto& v : w) { __half t = __float2half((frand() * 2.f - 1.f) * 1.7f * ws); v = *(uint16_t*)&t; }
  for (auto& v : b) { __half t = __float2half(frand() * 0.1f); v = *(uint16_t*)&t; }
  std::vector<uint8_t> pw(oidnb200_conv_weight_bytes(conv)), pb(oidnb200_conv_bias_bytes(conv));
  if (oidnb200_conv_pack_weights(conv, w.data(), O, I1, I2, pw.data())) { printf("pack failed: %s\n", oidnb200_last_error()); return 3; }
  oidnb200_conv_pack_bias(conv, b.data(), O, pb.data());

  void *d1, *d2 = nullptr, *dw, *db, *dout, *dref, *dscr;
  CK(cudaMalloc(&d1, n1 * 2)); if (n2) CK(cudaMalloc(&d2, n2 * 2));
  CK(cudaMalloc(&dw, pw.size())); CK(cudaMalloc(&db, pb.size()));
  CK(cudaMalloc(&dout, nout * 2)); CK(cudaMalloc(&dref, nout * 2)); CK(cudaMalloc(&dscr, (size_t)d.H * d.W * d.Cout * 2));
  CK(cudaMemcpy(d1, h1.data(), n1 * 2, cudaMemcpyHostToDevice));
  if (n2) CK(cudaMemcpy(d2, h2.data(), n2 * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, pw.data(), pw.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, pb.data(), pb.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, nout * 2)); CK(cudaMemset(dref, 0, nout * 2));

  if (oidnb200_conv_bind(conv, d1, d2, dw, db, dref)) { printf("bind failed: %s\n", oidnb200_last_error()); return 4; }
  if (oidnb200_conv_launch_simt(conv, dscr, 0)) { printf("simt failed: %s\n", oidnb200_last_error()); return 4; }
  CK(cudaDeviceSynchronize());
  if (oidnb200_conv_bind(conv, d1, d2, dw, db, dout)) { printf("bind failed: %s\n", oidnb200_last_error()); return 4; }
  if (oidnb200_conv_launch(conv, 0)) { printf("launch failed: %s\n", oidnb200_last_error()); return 5; }
  CK(cudaDeviceSynchronize());

  std::vector<__half> o(nout), r(nout);
  CK(cudaMemcpy(o.data(), dout, nout * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r.data(), dref, nout * 2, cudaMemcpyDeviceToHost));
  double maxd = 0, maxr = 0; size_t bad = 0, firstbad = (size_t)-1;
  for (size_t i = 0; i < nout; ++i)
  {
    const float a = __half2float(o[i]), c = __half2float(r[i]);
    const double df = fabs((double)a - c);
    if (!(df <= 2e-3 + 4e-3 * fabs(c))) { if (!bad) firstbad = i; ++bad; }
    if (df > maxd || df != df) maxd = df;
    if (fabs(c) > maxr) maxr = fabs(c);
  }
  printf("RESULT %s maxdiff=%.5g maxref=%.4g bad=%zu/%zu", bad ? "FAIL" : "PASS", maxd, maxr, bad, nout);
  if (bad)
  {
    const int Cd = d.Cout; const size_t px = firstbad / Cd;
    int Wd = d.W; if (d.post_op == 1) Wd /= 2; if (d.post_op == 2) Wd *= 2;
    printf(" first bad: y=%zu x=%zu c=%zu got=%g ref=%g", px / Wd, px % Wd, firstbad % Cd,
           __half2float(o[firstbad]), __half2float(r[firstbad]));
  }
  printf("\n");

  if (iters > 0 && !bad && getenv("PROBE_TRACE"))
  {
    unsigned long long* dtr; CK(cudaMalloc(&dtr, 192 * 8)); CK(cudaMemset(dtr, 0, 192 * 8));
    oidnb200_conv_set_trace(conv, dtr);
    oidnb200_conv_bind(conv, d1, d2, dw, db, dout);
    oidnb200_conv_launch(conv, 0); CK(cudaDeviceSynchronize());
    unsigned long long tr[192]; CK(cudaMemcpy(tr, dtr, sizeof(tr), cudaMemcpyDeviceToHost));
    const char* roles[12] = {"TMA0", "MMA0", "TMA1", "MMA1", "EPI0.0", "EPI0.1", "EPI0.2", "EPI0.3", "EPI1.0", "EPI1.1", "EPI1.2", "EPI1.3"};
    printf("TRACE (cycles summed over %d CTAs; tags: 1 A-empty 2 weights 3 tmem-empty 4 A-full 5 tmem-full 6 tmem-full(pool) 7 store-read; MMA 8 issue; EPI 8 tmem-ld 9 math+sts 10 fences 11 arrive+store)\n", info.grid);
    for (int w = 0; w < 12; ++w)
    {
      printf("  %-7s total %10.0f/CTA |", roles[w], (double)tr[w * 16] / info.grid);
      for (int t = 1; t < 16; ++t) if (tr[w * 16 + t]) printf(" t%d %5.1f%%", t, 100.0 * tr[w * 16 + t] / (double)tr[w * 16]);
      printf("\n");
    }
    oidnb200_conv_set_trace(conv, nullptr);
    oidnb200_conv_bind(conv, d1, d2, dw, db, dout);
  }
  if (iters > 0 && !bad)
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) oidnb200_conv_launch(conv, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) oidnb200_conv_launch(conv, 0);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    const double flop = 2.0 * 9 * (d.C1 + d.C2) * d.Cout * (double)d.H * d.W;
    printf("TIME %.4f ms  %.1f TFLOP/s (padded channels)\n", ms, flop / ms * 1e-9);
  }
  return bad ? 10 : 0;
}
