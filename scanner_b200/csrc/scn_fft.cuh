// scn_fft.cuh -- register-resident radix-2/4/8/16 DFT butterflies and the Stockham pass
// plan used by the fused spectrum-sense kernel (sm_100a).
//
// Semantics follow the reference's FFT wrapper, fft.cpp:4-25: forward transform
// (exponent sign -1), unnormalised, natural-order output, size N = 2^LOG2N.
//
// Layout: every thread owns PTS = 16 complex points of one transform, T = N/16 threads
// per transform.  A transform is a sequence of Stockham autosort passes of radix
// 16,16,..,R_last (R_last in {2,4,8,16}); between passes the points are exchanged through
// a padded shared-memory tile.  In every pass thread t reads points t + q*T (q = 0..15),
// so global/shared reads are always unit-stride across a warp, and after the last pass
// thread t holds output bins t + q*T.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scn {

constexpr int kPts = 16;   // complex points per thread

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by -i : (x + iy)(-i) = y - ix
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

constexpr float kSqrtHalf = 0.70710678118654752440f;
constexpr float kC16_1 = 0.92387953251128675613f;   // cos(pi/8)
constexpr float kS16_1 = 0.38268343236508977173f;   // sin(pi/8)

// a * W8^1 = a * (1 - i)/sqrt(2)
__device__ __forceinline__ float2 mul_w8_1(float2 a) {
  return make_float2((a.x + a.y) * kSqrtHalf, (a.y - a.x) * kSqrtHalf);
}
// a * W8^3 = a * (-1 - i)/sqrt(2)
__device__ __forceinline__ float2 mul_w8_3(float2 a) {
  return make_float2((a.y - a.x) * kSqrtHalf, -(a.x + a.y) * kSqrtHalf);
}
// a * W16^1 = a * (c - i s)
__device__ __forceinline__ float2 mul_w16_1(float2 a) {
  return make_float2(a.x * kC16_1 + a.y * kS16_1, a.y * kC16_1 - a.x * kS16_1);
}
// a * W16^3 = a * (s - i c)
__device__ __forceinline__ float2 mul_w16_3(float2 a) {
  return make_float2(a.x * kS16_1 + a.y * kC16_1, a.y * kS16_1 - a.x * kC16_1);
}
// a * W16^9 = -a * W16^1
__device__ __forceinline__ float2 mul_w16_9(float2 a) {
  return make_float2(-(a.x * kC16_1 + a.y * kS16_1), a.x * kS16_1 - a.y * kC16_1);
}

// In-place radix-2 on v[O], v[O+S].
template <int S, int O>
__device__ __forceinline__ void dft2(float2 (&v)[kPts]) {
  float2 a = v[O], b = v[O + S];
  v[O] = cadd(a, b);
  v[O + S] = csub(a, b);
}

// In-place radix-4, natural order, on v[O + r*S], r = 0..3.
template <int S, int O>
__device__ __forceinline__ void dft4(float2 (&v)[kPts]) {
  float2 x0 = v[O], x1 = v[O + S], x2 = v[O + 2 * S], x3 = v[O + 3 * S];
  float2 y0 = cadd(x0, x2), y1 = csub(x0, x2);
  float2 y2 = cadd(x1, x3), y3 = mul_mi(csub(x1, x3));
  v[O] = cadd(y0, y2);
  v[O + S] = cadd(y1, y3);
  v[O + 2 * S] = csub(y0, y2);
  v[O + 3 * S] = csub(y1, y3);
}

// In-place radix-8, natural order, on v[O + r*S], r = 0..7.
// r = 4*r1 + r0, q = q0 + 2*q1:  radix-2 over r1, twiddle W8^(r0*q0), radix-4 over r0.
template <int S, int O>
__device__ __forceinline__ void dft8(float2 (&v)[kPts]) {
  float2 a[8];
#pragma unroll
  for (int r0 = 0; r0 < 4; r0++) {
    float2 lo = v[O + r0 * S], hi = v[O + (r0 + 4) * S];
    a[r0] = cadd(lo, hi);        // q0 = 0
    a[r0 + 4] = csub(lo, hi);    // q0 = 1
  }
  a[5] = mul_w8_1(a[5]);
  a[6] = mul_mi(a[6]);
  a[7] = mul_w8_3(a[7]);
#pragma unroll
  for (int q0 = 0; q0 < 2; q0++) {
    float2 x0 = a[4 * q0], x1 = a[4 * q0 + 1], x2 = a[4 * q0 + 2], x3 = a[4 * q0 + 3];
    float2 y0 = cadd(x0, x2), y1 = csub(x0, x2);
    float2 y2 = cadd(x1, x3), y3 = mul_mi(csub(x1, x3));
    // output q = q0 + 2*q1
    v[O + (q0 + 0) * S] = cadd(y0, y2);
    v[O + (q0 + 2) * S] = cadd(y1, y3);
    v[O + (q0 + 4) * S] = csub(y0, y2);
    v[O + (q0 + 6) * S] = csub(y1, y3);
  }
}

// In-place radix-16, natural order, on v[r], r = 0..15.
// r = 4*r1 + r0, q = q0 + 4*q1:  radix-4 over r1, twiddle W16^(r0*q0), radix-4 over r0.
__device__ __forceinline__ void dft16(float2 (&v)[kPts]) {
  float2 a[16];
#pragma unroll
  for (int r0 = 0; r0 < 4; r0++) {
    float2 x0 = v[r0], x1 = v[r0 + 4], x2 = v[r0 + 8], x3 = v[r0 + 12];
    float2 y0 = cadd(x0, x2), y1 = csub(x0, x2);
    float2 y2 = cadd(x1, x3), y3 = mul_mi(csub(x1, x3));
    a[r0] = cadd(y0, y2);          // q0 = 0
    a[r0 + 4] = cadd(y1, y3);      // q0 = 1
    a[r0 + 8] = csub(y0, y2);      // q0 = 2
    a[r0 + 12] = csub(y1, y3);     // q0 = 3
  }
  // a[r0 + 4*q0] *= W16^(r0*q0)
  a[5] = mul_w16_1(a[5]);              // 1*1
  a[6] = mul_w8_1(a[6]);               // 2*1 -> W16^2
  a[7] = mul_w16_3(a[7]);              // 3*1
  a[9] = mul_w8_1(a[9]);               // 1*2
  a[10] = mul_mi(a[10]);               // 2*2 -> W16^4
  a[11] = mul_w8_3(a[11]);             // 3*2 -> W16^6
  a[13] = mul_w16_3(a[13]);            // 1*3
  a[14] = mul_w8_3(a[14]);             // 2*3 -> W16^6
  a[15] = mul_w16_9(a[15]);            // 3*3 -> W16^9
#pragma unroll
  for (int q0 = 0; q0 < 4; q0++) {
    float2 x0 = a[4 * q0], x1 = a[4 * q0 + 1], x2 = a[4 * q0 + 2], x3 = a[4 * q0 + 3];
    float2 y0 = cadd(x0, x2), y1 = csub(x0, x2);
    float2 y2 = cadd(x1, x3), y3 = mul_mi(csub(x1, x3));
    v[q0] = cadd(y0, y2);
    v[q0 + 4] = cadd(y1, y3);
    v[q0 + 8] = csub(y0, y2);
    v[q0 + 12] = csub(y1, y3);
  }
}

// ---- pass plan ---------------------------------------------------------------------
// Pass p has radix 16 while at least 4 bits remain, else the remainder.
__host__ __device__ constexpr int num_passes(int log2n) { return (log2n + 3) / 4; }
__host__ __device__ constexpr int pass_log2r(int log2n, int p) {
  return (log2n - 4 * p) >= 4 ? 4 : (log2n - 4 * p);
}
// Complex twiddles stored for pass p >= 1: (16/R)*(R-1) per thread.
__host__ __device__ constexpr int pass_tw_per_thread(int log2n, int p) {
  return 16 - (16 >> pass_log2r(log2n, p));
}
__host__ __device__ constexpr int pass_tw_offset(int log2n, int p) {   // in units of T complex
  int off = 0;
  for (int i = 1; i < p; i++) off += pass_tw_per_thread(log2n, i);
  return off;
}
__host__ __device__ constexpr int total_tw_per_thread(int log2n) {
  return pass_tw_offset(log2n, num_passes(log2n));
}

// Padded index into the exchange tile: one float2 of padding per 16 (keeps both the
// stride-16 writes of pass 0 and the unit-stride accesses of every other pass
// bank-conflict free for 64-bit accesses).
__device__ __forceinline__ int xpad(int idx) { return idx + (idx >> 4); }
__host__ __device__ constexpr int xch_elems(int n) { return n + (n >> 4); }

// Radix-R butterflies of one pass on the thread's 16 points (register slot q = m + r*M).
template <int LOG2R>
__device__ __forceinline__ void pass_butterflies(float2 (&v)[kPts]) {
  if constexpr (LOG2R == 4) {
    dft16(v);
  } else if constexpr (LOG2R == 3) {
    dft8<2, 0>(v);
    dft8<2, 1>(v);
  } else if constexpr (LOG2R == 2) {
    dft4<4, 0>(v); dft4<4, 1>(v); dft4<4, 2>(v); dft4<4, 3>(v);
  } else {
    dft2<8, 0>(v); dft2<8, 1>(v); dft2<8, 2>(v); dft2<8, 3>(v);
    dft2<8, 4>(v); dft2<8, 5>(v); dft2<8, 6>(v); dft2<8, 7>(v);
  }
}

// Twiddle multiply for pass P (P >= 1): slot q = m + r*M (r >= 1) *= W_{Ns*R}^{k*r},
// k = (t + m*T) mod Ns.  Table layout: tw[(off + m*(R-1) + (r-1)) * T + t].
template <int LOG2N, int P>
__device__ __forceinline__ void pass_twiddle(float2 (&v)[kPts], const float2* __restrict__ tw, int t) {
  constexpr int LOG2R = pass_log2r(LOG2N, P);
  constexpr int R = 1 << LOG2R, M = 16 / R, T = (1 << LOG2N) / 16;
  constexpr int OFF = pass_tw_offset(LOG2N, P);
#pragma unroll
  for (int m = 0; m < M; m++) {
#pragma unroll
    for (int r = 1; r < R; r++) {
      float2 w = __ldg(&tw[(OFF + m * (R - 1) + (r - 1)) * T + t]);
      v[m + r * M] = cmul(v[m + r * M], w);
    }
  }
}

// Scatter of pass P's outputs into the exchange tile (Stockham autosort index):
// butterfly j = t + m*T, k = j mod Ns, j0 = (j - k)*R + k, output r -> j0 + r*Ns.
// With the 1-in-16 padding the padded address is LINEAR in r:
//   Ns == 1 : xpad(16 j + r)     = 17 j + r
//   Ns >= 16: xpad(j0 + r Ns)    = xpad(j0) + r (Ns + Ns/16)
// so every store is base + immediate.
template <int LOG2N, int P>
__device__ __forceinline__ void pass_scatter(const float2 (&v)[kPts], float2* __restrict__ xch, int t) {
  constexpr int LOG2R = pass_log2r(LOG2N, P);
  constexpr int R = 1 << LOG2R, M = 16 / R, T = (1 << LOG2N) / 16;
  constexpr int LOG2NS = 4 * P;
  constexpr int NS = 1 << LOG2NS;
  constexpr int RSTRIDE = (NS == 1) ? 1 : (NS + NS / 16);
#pragma unroll
  for (int m = 0; m < M; m++) {
    const int j = t + m * T;
    const int k = j & (NS - 1);
    const int j0 = ((j - k) << LOG2R) + k;
    float2* base = xch + xpad(j0);
#pragma unroll
    for (int r = 0; r < R; r++) base[r * RSTRIDE] = v[m + r * M];
  }
}

// Gather for the next pass: point t + q*T, i.e. xpad(t) + q*(T + T/16): base + immediate.
template <int LOG2N>
__device__ __forceinline__ void pass_gather(float2 (&v)[kPts], const float2* __restrict__ xch, int t) {
  constexpr int T = (1 << LOG2N) / 16;
  const float2* base = xch + xpad(t);
#pragma unroll
  for (int q = 0; q < kPts; q++) v[q] = base[q * (T + T / 16)];
}

}  // namespace scn
