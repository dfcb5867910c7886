// Micro-benchmark for the round-2 question (VERDICT.md "next round" item 3): the three serial sweeps of one interior-point
// iteration of the kinematic solver (backward Riccati, forward roll-out of the step, adjoint recursion) with ONE SCENARIO
// PER LANE instead of one per warp.  Stage records are structure-of-arrays across scenarios in global memory
// ([field][stage][scenario], scenario fastest: a warp's access to one field of one stage is one 256-byte line), read with
// plain coalesced loads, optionally one stage ahead (software prefetch into registers).
//
// The arithmetic is the shipped sweep's (csrc/mpcb_kernel.cuh: riccati_backward / riccati_forward / adjoint, plain
// obstacle rows: 6 + 1 Hessian entries, 6 Jacobian entries), on synthetic records that keep F_uu positive definite.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o lane_sweep lane_sweep.cu
//   ./lane_sweep [scenarios] [warps_per_block] [blocks_per_sm_cap] [reps]
//
// Output: scenario-stages per second per SM for the three sweeps together, to be compared with the shipped kernel's
// 11 M stage-iterations/s/SM in its sweeps (708 k solves/s x 28.1 iterations x 51 stages / 148 SMs / 0.62 of the time).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int N = 50, S = N + 1;
// record fields per stage
enum { CDEF = 0, JAC = 4, HXX = 10, HUX = 16, GX = 17, HUU = 21, EE = 23, GU = 25, TK = 27, NIN = 29,
       KX = 29, KW = 37, KK = 41, DX = 43, DU = 47, LAMP = 49, NF = 53 };

__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}

// element (field f, stage k) of scenario s
#define AT(f, k) rec[((size_t)(f) * S + (k)) * B + s]

__global__ void fill_kernel(double *rec, size_t B) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  unsigned h = (unsigned)s * 2654435761u + 12345u;
  auto rnd = [&]() { h = h * 1664525u + 1013904223u; return (double)(h >> 8) * (1.0 / 16777216.0) - 0.5; };
  for (int k = 0; k <= N; k++) {
    for (int i = 0; i < 4; i++) AT(CDEF + i, k) = 1e-2 * rnd();
    for (int i = 0; i < 6; i++) AT(JAC + i, k) = 0.2 * rnd();
    AT(HXX + 0, k) = 1.0 + 0.1 * rnd(); AT(HXX + 1, k) = 0.05 * rnd(); AT(HXX + 2, k) = 2.0 + 0.1 * rnd();
    AT(HXX + 3, k) = 1.5 + 0.1 * rnd(); AT(HXX + 4, k) = 0.05 * rnd(); AT(HXX + 5, k) = 1.2 + 0.1 * rnd();
    AT(HUX, k) = 0.05 * rnd();
    for (int i = 0; i < 4; i++) AT(GX + i, k) = rnd();
    AT(HUU + 0, k) = 5.0 + rnd(); AT(HUU + 1, k) = 4.0 + rnd();
    AT(EE + 0, k) = 1.0 + 0.2 * rnd(); AT(EE + 1, k) = 0.5 + 0.2 * rnd();
    for (int i = 0; i < 2; i++) { AT(GU + i, k) = rnd(); AT(TK + i, k) = 0.1 * rnd(); }
  }
}

template <bool PREFETCH>
__global__ void __launch_bounds__(256) lane_sweep_kernel(double *__restrict__ rec, size_t B, int reps, double T, int *notok) {
  const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  int bad = 0;
  for (int rep = 0; rep < reps; rep++) {
    // ---------------- backward Riccati sweep (value function on [x; previous control])
    double p00 = AT(HXX + 0, N), p01 = AT(HXX + 1, N), p11 = AT(HXX + 2, N), p22 = AT(HXX + 3, N);
    double p23 = AT(HXX + 4, N), p33 = AT(HXX + 5, N), p02 = 0, p03 = 0, p12 = 0, p13 = 0;
    double px0 = AT(GX + 0, N), px1 = AT(GX + 1, N), px2 = AT(GX + 2, N), px3 = AT(GX + 3, N);
    double w0d = 0, w0a = 0, w1d = 0, w1a = 0, w2d = 0, w2a = 0, w3d = 0, w3a = 0;
    double qdd = 0, qda = 0, qaa = 0, pwd = 0, pwa = 0;
    double in[NIN], nx[NIN];
    if (PREFETCH) {
#pragma unroll
      for (int f = 0; f < NIN; f++) nx[f] = f < 4 ? AT(f, N) : AT(f, N - 1);  // CDEF of stage k+1, the rest of stage k
    }
#pragma unroll 1
    for (int k = N - 1; k >= 0; k--) {
      if (PREFETCH) {
#pragma unroll
        for (int f = 0; f < NIN; f++) in[f] = nx[f];
        if (k > 0) {
#pragma unroll
          for (int f = 0; f < NIN; f++) nx[f] = f < 4 ? AT(f, k) : AT(f, k - 1);
        }
      } else {
#pragma unroll
        for (int f = 0; f < NIN; f++) in[f] = f < 4 ? AT(f, k + 1) : AT(f, k);
      }
      const double a02 = in[JAC + 0], a03 = in[JAC + 1], a12 = in[JAC + 2], a13 = in[JAC + 3], a23 = in[JAC + 4], b2 = in[JAC + 5];
      const double Ed = in[EE + 0], Ea = in[EE + 1], td = in[TK + 0], ta = in[TK + 1];
      const double b0 = -in[CDEF + 0], b1 = -in[CDEF + 1], b2_ = -in[CDEF + 2], b3 = -in[CDEF + 3];
      const double m02 = p02 + a02 * p00 + a12 * p01;
      const double m12 = p12 + a02 * p01 + a12 * p11;
      const double m22 = p22 + a02 * p02 + a12 * p12;
      const double m32 = p23 + a02 * p03 + a12 * p13;
      const double m03 = p03 + a03 * p00 + a13 * p01 + a23 * p02;
      const double m13 = p13 + a03 * p01 + a13 * p11 + a23 * p12;
      const double m23 = p23 + a03 * p02 + a13 * p12 + a23 * p22;
      const double m33 = p33 + a03 * p03 + a13 * p13 + a23 * p23;
      const double f00 = in[HXX + 0] + p00, f01 = in[HXX + 1] + p01, f11 = in[HXX + 2] + p11;
      const double f02 = m02, f03 = m03, f12 = m12, f13 = m13;
      const double f22 = in[HXX + 3] + m22 + a02 * m02 + a12 * m12;
      const double f23 = in[HXX + 4] + m23 + a02 * m03 + a12 * m13;
      const double f33 = in[HXX + 5] + m33 + a03 * m03 + a13 * m13 + a23 * m23;
      const double ud0 = b2 * p02 + w0d, ud1 = b2 * p12 + w1d;
      const double ud2 = b2 * m22 + w2d + a02 * w0d + a12 * w1d;
      const double ud3 = in[HUX] + b2 * m23 + w3d + a03 * w0d + a13 * w1d + a23 * w2d;
      const double ua0 = T * p03 + w0a, ua1 = T * p13 + w1a;
      const double ua2 = T * m32 + w2a + a02 * w0a + a12 * w1a;
      const double ua3 = T * m33 + w3a + a03 * w0a + a13 * w1a + a23 * w2a;
      const double Fdd = in[HUU + 0] + Ed + qdd + b2 * (b2 * p22 + 2.0 * w2d);
      const double Fda = qda + b2 * (T * p23) + b2 * w2a + T * w3d;
      const double Faa = in[HUU + 1] + Ea + qaa + T * (T * p33 + 2.0 * w3a);
      const double Pb0 = px0 + p00 * b0 + p01 * b1 + p02 * b2_ + p03 * b3;
      const double Pb1 = px1 + p01 * b0 + p11 * b1 + p12 * b2_ + p13 * b3;
      const double Pb2 = px2 + p02 * b0 + p12 * b1 + p22 * b2_ + p23 * b3;
      const double Pb3 = px3 + p03 * b0 + p13 * b1 + p23 * b2_ + p33 * b3;
      const double fx0 = in[GX + 0] + Pb0, fx1 = in[GX + 1] + Pb1;
      const double fx2 = in[GX + 2] + Pb2 + a02 * Pb0 + a12 * Pb1;
      const double fx3 = in[GX + 3] + Pb3 + a03 * Pb0 + a13 * Pb1 + a23 * Pb2;
      const double fud = in[GU + 0] + td + pwd + b2 * Pb2 + w0d * b0 + w1d * b1 + w2d * b2_ + w3d * b3;
      const double fua = in[GU + 1] + ta + pwa + T * Pb3 + w0a * b0 + w1a * b1 + w2a * b2_ + w3a * b3;
      const double det = Fdd * Faa - Fda * Fda;
      if (!(Fdd > 0.0) || !(det > 0.0)) bad++;
      const double id = fast_rcp(det);
      const double idd = Faa * id, ida = -Fda * id, iaa = Fdd * id;
      const double kd0 = -(idd * ud0 + ida * ua0), kd1 = -(idd * ud1 + ida * ua1), kd2 = -(idd * ud2 + ida * ua2), kd3 = -(idd * ud3 + ida * ua3);
      const double ka0 = -(ida * ud0 + iaa * ua0), ka1 = -(ida * ud1 + iaa * ua1), ka2 = -(ida * ud2 + iaa * ua2), ka3 = -(ida * ud3 + iaa * ua3);
      const double wdd = idd * Ed, wda = ida * Ea, wad = ida * Ed, waa = iaa * Ea;
      const double kkd = -(idd * fud + ida * fua), kka = -(ida * fud + iaa * fua);
      AT(KX + 0, k) = kd0; AT(KX + 1, k) = kd1; AT(KX + 2, k) = kd2; AT(KX + 3, k) = kd3;
      AT(KX + 4, k) = ka0; AT(KX + 5, k) = ka1; AT(KX + 6, k) = ka2; AT(KX + 7, k) = ka3;
      AT(KW + 0, k) = wdd; AT(KW + 1, k) = wda; AT(KW + 2, k) = wad; AT(KW + 3, k) = waa;
      AT(KK + 0, k) = kkd; AT(KK + 1, k) = kka;
      p00 = f00 + ud0 * kd0 + ua0 * ka0;
      p11 = f11 + ud1 * kd1 + ua1 * ka1;
      p22 = f22 + ud2 * kd2 + ua2 * ka2;
      p33 = f33 + ud3 * kd3 + ua3 * ka3;
      p01 = f01 + ud0 * kd1 + ua0 * ka1;
      p02 = f02 + ud0 * kd2 + ua0 * ka2;
      p03 = f03 + ud0 * kd3 + ua0 * ka3;
      p12 = f12 + ud1 * kd2 + ua1 * ka2;
      p13 = f13 + ud1 * kd3 + ua1 * ka3;
      p23 = f23 + ud2 * kd3 + ua2 * ka3;
      w0d = ud0 * wdd + ua0 * wad; w0a = ud0 * wda + ua0 * waa;
      w1d = ud1 * wdd + ua1 * wad; w1a = ud1 * wda + ua1 * waa;
      w2d = ud2 * wdd + ua2 * wad; w2a = ud2 * wda + ua2 * waa;
      w3d = ud3 * wdd + ua3 * wad; w3a = ud3 * wda + ua3 * waa;
      px0 = fx0 + ud0 * kkd + ua0 * kka;
      px1 = fx1 + ud1 * kkd + ua1 * kka;
      px2 = fx2 + ud2 * kkd + ua2 * kka;
      px3 = fx3 + ud3 * kkd + ua3 * kka;
      qdd = Ed - Ed * wdd;
      qda = -0.5 * (Ed * wda + Ea * wad);
      qaa = Ea - Ea * waa;
      pwd = -td - Ed * kkd;
      pwa = -ta - Ea * kka;
    }
    // ---------------- forward sweep
    double d0 = -AT(CDEF + 0, 0), d1 = -AT(CDEF + 1, 0), d2 = -AT(CDEF + 2, 0), d3 = -AT(CDEF + 3, 0);
    double vd = 0, va = 0;
    AT(DX + 0, 0) = d0; AT(DX + 1, 0) = d1; AT(DX + 2, 0) = d2; AT(DX + 3, 0) = d3;
#pragma unroll 1
    for (int k = 0; k < N; k++) {
      double g[14], j[6], c[4];
#pragma unroll
      for (int f = 0; f < 14; f++) g[f] = AT(KX + f, k);
#pragma unroll
      for (int f = 0; f < 6; f++) j[f] = AT(JAC + f, k);
#pragma unroll
      for (int f = 0; f < 4; f++) c[f] = AT(CDEF + f, k + 1);
      const double ud = g[12] + g[0] * d0 + g[1] * d1 + g[2] * d2 + g[3] * d3 + g[8] * vd + g[9] * va;
      const double ua = g[13] + g[4] * d0 + g[5] * d1 + g[6] * d2 + g[7] * d3 + g[10] * vd + g[11] * va;
      const double n0 = d0 + j[0] * d2 + j[1] * d3 - c[0];
      const double n1 = d1 + j[2] * d2 + j[3] * d3 - c[1];
      const double n2 = d2 + j[4] * d3 + j[5] * ud - c[2];
      const double n3 = d3 + T * ua - c[3];
      AT(DU + 0, k) = ud; AT(DU + 1, k) = ua;
      AT(DX + 0, k + 1) = n0; AT(DX + 1, k + 1) = n1; AT(DX + 2, k + 1) = n2; AT(DX + 3, k + 1) = n3;
      d0 = n0; d1 = n1; d2 = n2; d3 = n3; vd = ud; va = ua;
    }
    // ---------------- adjoint sweep (stage residual r_k formed on the fly)
    double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
#pragma unroll 1
    for (int k = N; k >= 0; k--) {
      double h[7], gx[4], dx[4], j[6];
#pragma unroll
      for (int f = 0; f < 7; f++) h[f] = AT(HXX + f, k);
#pragma unroll
      for (int f = 0; f < 4; f++) { gx[f] = AT(GX + f, k); dx[f] = AT(DX + f, k); }
      const double ud = k < N ? AT(DU + 0, k) : 0.0;
      const double r0 = gx[0] + h[0] * dx[0] + h[1] * dx[1];
      const double r1 = gx[1] + h[1] * dx[0] + h[2] * dx[1];
      const double r2 = gx[2] + h[3] * dx[2] + h[4] * dx[3];
      const double r3 = gx[3] + h[4] * dx[2] + h[5] * dx[3] + h[6] * ud;
      double n0, n1, n2, n3;
      if (k == N) { n0 = -r0; n1 = -r1; n2 = -r2; n3 = -r3; }
      else {
#pragma unroll
        for (int f = 0; f < 6; f++) j[f] = AT(JAC + f, k);
        n0 = l0 - r0;
        n1 = l1 - r1;
        n2 = l2 + j[0] * l0 + j[2] * l1 - r2;
        n3 = l3 + j[1] * l0 + j[3] * l1 + j[4] * l2 - r3;
      }
      AT(LAMP + 0, k) = n0; AT(LAMP + 1, k) = n1; AT(LAMP + 2, k) = n2; AT(LAMP + 3, k) = n3;
      l0 = n0; l1 = n1; l2 = n2; l3 = n3;
    }
  }
  if (bad) atomicAdd(notok, bad);
}

int main(int argc, char **argv) {
  size_t B = argc > 1 ? atoll(argv[1]) : 37888;
  int tpb = argc > 2 ? atoi(argv[2]) * 32 : 128;
  int reps = argc > 3 ? atoi(argv[3]) : 20;
  int dev = 0, sms = 0;
  CHECK(cudaGetDevice(&dev));
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *rec;
  int *notok;
  size_t bytes = (size_t)NF * S * B * sizeof(double);
  CHECK(cudaMalloc(&rec, bytes));
  CHECK(cudaMalloc(&notok, sizeof(int)));
  CHECK(cudaMemset(notok, 0, sizeof(int)));
  fill_kernel<<<(unsigned)((B + 255) / 256), 256>>>(rec, B);
  CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int pf = 0; pf < 2; pf++) {
    unsigned grid = (unsigned)((B + tpb - 1) / tpb);
    for (int w = 0; w < 2; w++) {
      CHECK(cudaEventRecord(e0));
      if (pf) lane_sweep_kernel<true><<<grid, tpb>>>(rec, B, reps, 0.1, notok);
      else lane_sweep_kernel<false><<<grid, tpb>>>(rec, B, reps, 0.1, notok);
      CHECK(cudaEventRecord(e1));
      CHECK(cudaEventSynchronize(e1));
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int bad = 0;
    CHECK(cudaMemcpy(&bad, notok, sizeof(int), cudaMemcpyDeviceToHost));
    double stages = (double)B * reps * S;
    // bytes moved per scenario-stage: backward 29 in + 14 out, forward 24 in + 6 out, adjoint 22 in + 4 out
    double gb = stages * (29 + 14 + 24 + 6 + 22 + 4) * 8 / 1e9;
    printf("scenarios %zu (%.0f MB of records), block %d, prefetch %d: %.3f ms for %d sweeps -> %.1f M scenario-stages/s/SM, %.0f GB/s of record traffic, "
           "%.2f M sweep-triples/s (not pos. def.: %d)\n",
           B, bytes / 1e6, tpb, pf, ms, reps, stages / (ms * 1e-3) / sms / 1e6, gb / (ms * 1e-3), (double)B * reps / (ms * 1e-3) / 1e6, bad);
  }
  return 0;
}
