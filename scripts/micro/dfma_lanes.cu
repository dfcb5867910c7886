// Does the FP64 pipe skip inactive half-warps?  Same number of warp-level DFMA instructions,
// different active-lane masks.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b, unsigned mask) {
  double v[8];
  for (int i = 0; i < 8; i++) v[i] = (threadIdx.x + i) * 1e-3;
  if ((mask >> (threadIdx.x & 31)) & 1u) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fma(v[i], a, b);
      }
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += v[i];
  if (s == 123.456) out[0] = s;
}
int main() {
  double *d; cudaMalloc(&d, 8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned masks[] = {0xffffffffu, 0x0000ffffu, 0x000000ffu, 0x00000001u, 0x00010001u, 0x55555555u, 0x0f0f0f0fu};
  const char *names[] = {"all32", "lanes0-15", "lanes0-7", "lane0", "lanes{0,16}", "even lanes", "0f0f0f0f"};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 8; warps >= 1; warps /= 8)
  for (int m = 0; m < 7; m++) {
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      k<<<sms * 8, 32 * warps>>>(d, 4096, 0.999999, 1e-9, masks[m]);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    double winst = 32.0 * 4096 * sms * 8 * warps;  // warp-level DFMA instructions
    printf("warps/block %d  %-12s %.3f ms  %.2f warp-DFMA/clk/SM (at 1.965 GHz)\n", warps, names[m], best, winst / (best * 1e-3) / 1.965e9 / sms);
  }
  return 0;
}
