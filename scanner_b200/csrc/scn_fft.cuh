// scn_fft.cuh -- register-resident radix-2/4/8/16 DFT butterflies on packed fp32x2 math and the
// Stockham pass plan used by the fused spectrum-sense kernel (sm_100a).
//
// Semantics follow the reference's FFT wrapper, fft.cpp:4-25: forward transform (exponent sign
// -1), unnormalised, natural-order output, size N = 2^LOG2N.
//
// Blackwell specifics: a complex value is a float2 in an aligned register pair and every complex
// add / subtract / rotate / multiply is one or two packed FADD2 / FFMA2 / FMUL2 instructions
// (add.f32x2 / fma.rn.f32x2 / mul.f32x2, sm_100+).  The SASS operand modifiers make the complex
// idioms free: `.LO_HI` swaps the halves, `.NP` negates one half, `.F32` broadcasts a scalar, so
//   a + (-i) d        = FFMA2(d.LO_HI.NP, 1, a)                         (1 instruction)
//   v * (c + i s)     = FMUL2(v, c.F32) ; FFMA2(-v.LO_HI.NP, s.F32, .)   (2 instructions)
// which halves the issue slots of the transform relative to scalar FADD/FMUL/FFMA.
//
// Layout: every thread owns 16 complex points of one transform, T = N/16 threads per transform.
// Passes: radix R0 = 2^(LOG2N mod 4) first (16 when that is 1), then radix 16 throughout.
//   pass 0 : thread t owns butterflies j = M0*t + m (M0 = 16/R0 consecutive columns), so each of
//            its R0 rows is a run of M0 CONSECUTIVE samples -> one vector load per row;
//            register slot q = m + r*M0.
//   pass p : thread t owns butterfly j = t; it gathers points t + q*T from the exchange tile.
// After the last pass thread t holds output bins t + q*T (q = slot).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scn {

// Raw IQ is read exactly once, through the read-only path.  (Loading it WITHOUT allocating in L1 --
// ld.global.nc.L1::no_allocate, to keep the stream from evicting the window / twiddle tables -- was measured on B200:
// no gain for the 1- and 2-byte kinds, 3 % slower for fp32 IQ at N = 4096 / 8192; removed.  What fixed the table
// misses there was halving the table footprint and, later, landing the raw stream by TMA: scn_p64.cuh.)
__device__ __forceinline__ unsigned short ldg_stream(const unsigned short* p) { return __ldg(p); }
__device__ __forceinline__ unsigned int ldg_stream(const unsigned int* p) { return __ldg(p); }
__device__ __forceinline__ uint2 ldg_stream(const uint2* p) { return __ldg(p); }
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) { return __ldg(p); }
__device__ __forceinline__ float2 ldg_stream(const float2* p) { return __ldg(p); }

constexpr int kPts = 16;   // complex points per thread

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// v * t, t = (re, im):  v*t.re + (i v)*t.im,  i v = (-v.y, v.x)
__device__ __forceinline__ float2 cmul(float2 v, float2 t) {
  const float2 r = __fmul2_rn(v, make_float2(t.x, t.x));
  return __ffma2_rn(make_float2(-v.y, v.x), make_float2(t.y, t.y), r);
}
// a + (-i) d  and  a - (-i) d          ((-i) d = (d.y, -d.x))
__device__ __forceinline__ float2 add_mi(float2 a, float2 d) {
  return __ffma2_rn(make_float2(d.y, d.x), make_float2(1.0f, -1.0f), a);
}
__device__ __forceinline__ float2 sub_mi(float2 a, float2 d) {
  return __ffma2_rn(make_float2(d.y, d.x), make_float2(-1.0f, 1.0f), a);
}

constexpr float kSqrtHalf = 0.70710678118654752440f;
constexpr float kC16 = 0.92387953251128675613f;   // cos(pi/8)
constexpr float kS16 = 0.38268343236508977173f;   // sin(pi/8)

// Radix-4 on x0..x3 -> natural order outputs.  X2MI: x2 carries a pending factor (-i).
template <bool X2MI>
__device__ __forceinline__ void bfly4(float2 x0, float2 x1, float2 x2, float2 x3, float2& o0, float2& o1,
                                      float2& o2, float2& o3) {
  const float2 y0 = X2MI ? add_mi(x0, x2) : cadd(x0, x2);
  const float2 y1 = X2MI ? sub_mi(x0, x2) : csub(x0, x2);
  const float2 y2 = cadd(x1, x3);
  const float2 d = csub(x1, x3);
  o0 = cadd(y0, y2);
  o1 = add_mi(y1, d);
  o2 = csub(y0, y2);
  o3 = sub_mi(y1, d);
}

// In-place radix-2 on v[O], v[O+S].
template <int S, int O>
__device__ __forceinline__ void dft2(float2 (&v)[kPts]) {
  const float2 a = v[O], b = v[O + S];
  v[O] = cadd(a, b);
  v[O + S] = csub(a, b);
}

// In-place radix-4, natural order, on v[O + r*S].
template <int S, int O>
__device__ __forceinline__ void dft4(float2 (&v)[kPts]) {
  bfly4<false>(v[O], v[O + S], v[O + 2 * S], v[O + 3 * S], v[O], v[O + S], v[O + 2 * S], v[O + 3 * S]);
}

// In-place radix-8, natural order, on v[O + r*S].
// r = 4*r1 + r0, q = q0 + 2*q1: radix-2 over r1, twiddle W8^(r0*q0), radix-4 over r0.
template <int S, int O>
__device__ __forceinline__ void dft8(float2 (&v)[kPts]) {
  float2 a[8];
#pragma unroll
  for (int r0 = 0; r0 < 4; r0++) {
    const float2 lo = v[O + r0 * S], hi = v[O + (r0 + 4) * S];
    a[r0] = cadd(lo, hi);        // q0 = 0
    a[r0 + 4] = csub(lo, hi);    // q0 = 1
  }
  a[5] = cmul(a[5], make_float2(kSqrtHalf, -kSqrtHalf));     // W8^1
  a[7] = cmul(a[7], make_float2(-kSqrtHalf, -kSqrtHalf));    // W8^3      (a[6] * W8^2 = -i: folded below)
  bfly4<false>(a[0], a[1], a[2], a[3], v[O], v[O + 2 * S], v[O + 4 * S], v[O + 6 * S]);
  bfly4<true>(a[4], a[5], a[6], a[7], v[O + S], v[O + 3 * S], v[O + 5 * S], v[O + 7 * S]);
}

// In-place radix-16, natural order, on v[r], r = 0..15.
// r = 4*r1 + r0, q = q0 + 4*q1: radix-4 over r1, twiddle W16^(r0*q0), radix-4 over r0.
__device__ __forceinline__ void dft16(float2 (&v)[kPts]) {
  float2 a[16];
#pragma unroll
  for (int r0 = 0; r0 < 4; r0++)
    bfly4<false>(v[r0], v[r0 + 4], v[r0 + 8], v[r0 + 12], a[r0], a[r0 + 4], a[r0 + 8], a[r0 + 12]);
  // a[r0 + 4*q0] *= W16^(r0*q0);  W16^4 = -i on a[10] is folded into the second-stage butterfly
  a[5] = cmul(a[5], make_float2(kC16, -kS16));               // W16^1
  a[6] = cmul(a[6], make_float2(kSqrtHalf, -kSqrtHalf));     // W16^2
  a[7] = cmul(a[7], make_float2(kS16, -kC16));               // W16^3
  a[9] = cmul(a[9], make_float2(kSqrtHalf, -kSqrtHalf));     // W16^2
  a[11] = cmul(a[11], make_float2(-kSqrtHalf, -kSqrtHalf));  // W16^6
  a[13] = cmul(a[13], make_float2(kS16, -kC16));             // W16^3
  a[14] = cmul(a[14], make_float2(-kSqrtHalf, -kSqrtHalf));  // W16^6
  a[15] = cmul(a[15], make_float2(-kC16, kS16));             // W16^9
  bfly4<false>(a[0], a[1], a[2], a[3], v[0], v[4], v[8], v[12]);
  bfly4<false>(a[4], a[5], a[6], a[7], v[1], v[5], v[9], v[13]);
  bfly4<true>(a[8], a[9], a[10], a[11], v[2], v[6], v[10], v[14]);
  bfly4<false>(a[12], a[13], a[14], a[15], v[3], v[7], v[11], v[15]);
}

// ---- pass plan ---------------------------------------------------------------------
__host__ __device__ constexpr int num_passes(int log2n) { return (log2n + 3) / 4; }
__host__ __device__ constexpr int pass_log2r(int log2n, int p) {
  return (p == 0 && (log2n % 4) != 0) ? (log2n % 4) : 4;
}
// log2 of Ns(p) = product of the radices of the passes before p
__host__ __device__ constexpr int pass_log2ns(int log2n, int p) {
  int s = 0;
  for (int i = 0; i < p; i++) s += pass_log2r(log2n, i);
  return s;
}
// Twiddles: passes p >= 1 are radix 16 with 15 factors per thread; table layout
// tw[((p-1)*15 + (r-1)) * T + t] = exp(-2 pi i * (t mod Ns_p) * r / (16 Ns_p)).
__host__ __device__ constexpr int total_tw_per_thread(int log2n) { return 15 * (num_passes(log2n) - 1); }

// Padded index into the exchange tile: one float2 of padding per 16.  Keeps the stride-16
// scatter of pass 0, the grouped scatters of pass 1 (Ns = 2/4/8) and every unit-stride access
// bank-conflict free for 64-bit accesses, and makes every address base + immediate.
__device__ __forceinline__ int xpad(int idx) { return idx + (idx >> 4); }
__host__ __device__ constexpr int xch_elems(int n) { return n + (n >> 4); }

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

// Pass-0 scatter: butterfly j = M0*t + m writes outputs r to j*R0 + r = 16t + m*R0 + r,
// i.e. padded 17t + (m*R0 + r): sixteen consecutive float2 per thread.
template <int LOG2N>
__device__ __forceinline__ void pass0_scatter(const float2 (&v)[kPts], float2* __restrict__ xch, int t) {
  constexpr int LOG2R = pass_log2r(LOG2N, 0);
  constexpr int R = 1 << LOG2R, M = 16 / R;
  float2* base = xch + 17 * t;
#pragma unroll
  for (int m = 0; m < M; m++)
#pragma unroll
    for (int r = 0; r < R; r++) base[m * R + r] = v[m + r * M];
}

// Pass-p (p >= 1, radix 16) scatter: k = t mod Ns, j0 = (t - k)*16 + k, output r -> j0 + r*Ns.
// Padded address is xpad(j0) + (r*Ns + ((r*Ns) >> 4)) for every Ns = 2^s (no carry out of the low
// 4 bits because k < Ns): base + immediate.
template <int LOG2N, int P>
__device__ __forceinline__ void pass_scatter(const float2 (&v)[kPts], float2* __restrict__ xch, int t) {
  constexpr int NS = 1 << pass_log2ns(LOG2N, P);
  const int k = t & (NS - 1);
  float2* base = xch + xpad(((t - k) << 4) + k);
#pragma unroll
  for (int r = 0; r < 16; r++) base[r * NS + ((r * NS) >> 4)] = v[r];
}

// Gather for pass p >= 1: point t + q*T, padded xpad(t) + q*(T + T/16).
template <int LOG2N>
__device__ __forceinline__ void pass_gather(float2 (&v)[kPts], const float2* __restrict__ xch, int t) {
  constexpr int T = (1 << LOG2N) / 16;
  const float2* base = xch + xpad(t);
#pragma unroll
  for (int q = 0; q < kPts; q++) v[q] = base[q * (T + T / 16)];
}

// Same 15 factors from the first one by a depth-4 product tree (w^2, w^4, w^8, then products):
// 14 packed complex multiplies instead of 14 loads; <= 4 roundings deep.
template <int LOG2N, int P>
__device__ __forceinline__ void power_twiddles(float2 (&w)[15], const float2* __restrict__ tw, int t) {
  constexpr int T = (1 << LOG2N) / 16;
  const float2 w1 = __ldg(&tw[((P - 1) * 15) * T + t]);
  const float2 w2 = cmul(w1, w1), w4 = cmul(w2, w2), w8 = cmul(w4, w4);
  const float2 w3 = cmul(w2, w1), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
  w[0] = w1; w[1] = w2; w[2] = w3; w[3] = w4; w[4] = w5; w[5] = w6; w[6] = w7; w[7] = w8;
  w[8] = cmul(w8, w1); w[9] = cmul(w8, w2); w[10] = cmul(w8, w3); w[11] = cmul(w8, w4);
  w[12] = cmul(w8, w5); w[13] = cmul(w8, w6); w[14] = cmul(w8, w7);
}

// Accuracy-first variant: six table values (w^1..w^4, w^8, w^12) and nine single products
// w^(4a+b) = w^(4a) * w^b: one rounding deep, 6 loads + 9 multiplies instead of 15 loads.
template <int LOG2N, int P>
__device__ __forceinline__ void product_twiddles(float2 (&w)[15], const float2* __restrict__ tw, int t) {
  constexpr int T = (1 << LOG2N) / 16;
  const float2* base = tw + ((P - 1) * 15) * T + t;
  const float2 w1 = __ldg(base), w2 = __ldg(base + T), w3 = __ldg(base + 2 * T), w4 = __ldg(base + 3 * T);
  const float2 w8 = __ldg(base + 7 * T), w12 = __ldg(base + 11 * T);
  w[0] = w1; w[1] = w2; w[2] = w3; w[3] = w4;
  w[4] = cmul(w4, w1); w[5] = cmul(w4, w2); w[6] = cmul(w4, w3);
  w[7] = w8;
  w[8] = cmul(w8, w1); w[9] = cmul(w8, w2); w[10] = cmul(w8, w3);
  w[11] = w12;
  w[12] = cmul(w12, w1); w[13] = cmul(w12, w2); w[14] = cmul(w12, w3);
}

__device__ __forceinline__ void apply_twiddles(float2 (&v)[kPts], const float2 (&w)[15]) {
#pragma unroll
  for (int r = 1; r < 16; r++) v[r] = cmul(v[r], w[r - 1]);
}

}  // namespace scn
