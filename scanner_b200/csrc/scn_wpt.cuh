// scn_wpt.cuh -- warp-per-transform variant of the fused spectrum-sense kernel for N = 2048, int8 IQ,
// K = 1 (BASELINE.json configs[1], the headline workload).
//
// One WARP owns one transform: every lane holds 64 complex points, N = 32 x 64:
//   pass 0: radix-32 on two adjacent columns per lane (rows are 4-byte loads of 2 int8 IQ samples),
//   ONE exchange through a warp-private shared-memory tile (__syncwarp only -- no CTA barrier anywhere),
//   pass 1: radix-64 with twiddles W_2048^(lane*r), outputs lane + 32 q: coalesced stores.
// Compared with the 16-points-per-thread family (scn_kernel.cuh) this halves the shared-memory
// traffic (one 16 B/sample exchange instead of two), removes all block barriers, and makes the DC
// sum, the mask assembly and the hit-record ranks warp-local (REDUX / ballot only).
// Same arithmetic contract; parity is checked by the same tests.
#pragma once
#include "scn_fft.cuh"
#include "scn_kernel.cuh"
#include "scn_wconst.cuh"

namespace scn {

constexpr int kWptN = 2048;
constexpr int kWptWarpsPerCta = 2;
// tile padding: 1 float2 per 64 -> lane stride 65 elements: conflict-free 64-bit scatter, unit-stride gather
__host__ __device__ constexpr int wpt_tile_elems() { return kWptN + (kWptN / 64); }
// The next transform's 4 KB of int8 IQ is pulled into L2 by one bulk prefetch per warp (cp.async.bulk.prefetch.L2)
// right after the conversion and loaded (L2 hits) at the top of its own iteration.  Measured alternatives on B200
// (profiles/README.md), all removed: prefetching it into 32 registers one transform ahead (453 vs 458 Gsamples/s),
// staging it with one TMA bulk copy per transform (405: 32 extra LDS per lane and a tighter register allocation),
// issuing the loads after the pass-0 scatter (423), keeping the two warps of a CTA in lockstep for the instruction
// cache (+0.7 %, noise), 5-6 CTAs per SM at 168 registers (371; 417 without the register prefetch).
constexpr size_t kWptWarpBytes = sizeof(float2) * size_t(wpt_tile_elems());
static_assert(kWptWarpBytes % 16 == 0, "warp region must keep 16-byte alignment");
constexpr size_t kWptSmemBytes = kWptWarpBytes * kWptWarpsPerCta;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int K>
__device__ __forceinline__ float2 mul_w64(float2 a) {     // a * W64^K with the trivial cases folded
  constexpr int k = K & 63;
  if constexpr (k == 0) return a;
  else if constexpr (k == 16) return make_float2(a.y, -a.x);      // -i
  else if constexpr (k == 32) return make_float2(-a.x, -a.y);
  else if constexpr (k == 48) return make_float2(-a.y, a.x);      // +i
  else return cmul(a, make_float2(kW64re[k], kW64im[k]));
}

// Radix-8 of x[0..7] in place, natural order (same decomposition as dft8 in scn_fft.cuh).
__device__ __forceinline__ void r8(float2 (&x)[8]) {
  float2 a[8];
#pragma unroll
  for (int r0 = 0; r0 < 4; r0++) { a[r0] = cadd(x[r0], x[r0 + 4]); a[r0 + 4] = csub(x[r0], x[r0 + 4]); }
  a[5] = cmul(a[5], make_float2(kSqrtHalf, -kSqrtHalf));
  a[7] = cmul(a[7], make_float2(-kSqrtHalf, -kSqrtHalf));
  bfly4<false>(a[0], a[1], a[2], a[3], x[0], x[2], x[4], x[6]);
  bfly4<true>(a[4], a[5], a[6], a[7], x[1], x[3], x[5], x[7]);
}

// Radix-32 in place on v[O .. O+31] (input slot r = natural index).  32 = 4 x 8:
// radix-4 over r1 (slots r0 + 8 r1), twiddle W32^(r0 q0), radix-8 over r0 (slots 8 q0 + r0).
// Result: slot 8 q0 + q1 holds output q = q0 + 4 q1.
template <int O>
__device__ __forceinline__ void dft32_inplace(float2 (&v)[64]) {
#pragma unroll
  for (int r0 = 0; r0 < 8; r0++)
    bfly4<false>(v[O + r0], v[O + r0 + 8], v[O + r0 + 16], v[O + r0 + 24],
                 v[O + r0], v[O + r0 + 8], v[O + r0 + 16], v[O + r0 + 24]);
  // slot r0 + 8 q0 *= W32^(r0 q0) = W64^(2 r0 q0)
#define SCN_TW32(R0, Q0) v[O + R0 + 8 * Q0] = mul_w64<2 * R0 * Q0>(v[O + R0 + 8 * Q0]);
#define SCN_TW32_ROW(Q0) SCN_TW32(1, Q0) SCN_TW32(2, Q0) SCN_TW32(3, Q0) SCN_TW32(4, Q0) SCN_TW32(5, Q0) SCN_TW32(6, Q0) SCN_TW32(7, Q0)
  SCN_TW32_ROW(1) SCN_TW32_ROW(2) SCN_TW32_ROW(3)
#undef SCN_TW32_ROW
#undef SCN_TW32
#pragma unroll
  for (int q0 = 0; q0 < 4; q0++) {
    float2 x[8];
#pragma unroll
    for (int r0 = 0; r0 < 8; r0++) x[r0] = v[O + 8 * q0 + r0];
    r8(x);
#pragma unroll
    for (int q1 = 0; q1 < 8; q1++) v[O + 8 * q0 + q1] = x[q1];
  }
}
__host__ __device__ constexpr int dft32_out_index(int slot) { return (slot >> 3) + 4 * (slot & 7); }

// Radix-64 in place on v[0..63]: 64 = 8 x 8; slot 8 q0 + q1 holds output q = q0 + 8 q1.
template <int R0>
__device__ __forceinline__ void tw64_row(float2 (&v)[64]) {   // slots R0 + 8 q0, q0 = 1..7
  v[R0 + 8] = mul_w64<R0 * 1>(v[R0 + 8]);   v[R0 + 16] = mul_w64<R0 * 2>(v[R0 + 16]);
  v[R0 + 24] = mul_w64<R0 * 3>(v[R0 + 24]); v[R0 + 32] = mul_w64<R0 * 4>(v[R0 + 32]);
  v[R0 + 40] = mul_w64<R0 * 5>(v[R0 + 40]); v[R0 + 48] = mul_w64<R0 * 6>(v[R0 + 48]);
  v[R0 + 56] = mul_w64<R0 * 7>(v[R0 + 56]);
}
__device__ __forceinline__ void dft64_inplace(float2 (&v)[64]) {
#pragma unroll
  for (int r0 = 0; r0 < 8; r0++) {
    float2 x[8];
#pragma unroll
    for (int r1 = 0; r1 < 8; r1++) x[r1] = v[r0 + 8 * r1];
    r8(x);
#pragma unroll
    for (int q0 = 0; q0 < 8; q0++) v[r0 + 8 * q0] = x[q0];
  }
  tw64_row<1>(v); tw64_row<2>(v); tw64_row<3>(v); tw64_row<4>(v); tw64_row<5>(v); tw64_row<6>(v); tw64_row<7>(v);
#pragma unroll
  for (int q0 = 0; q0 < 8; q0++) {
    float2 x[8];
#pragma unroll
    for (int r0 = 0; r0 < 8; r0++) x[r0] = v[8 * q0 + r0];
    r8(x);
#pragma unroll
    for (int q1 = 0; q1 < 8; q1++) v[8 * q0 + q1] = x[q1];
  }
}
__host__ __device__ constexpr int dft64_out_index(int slot) { return (slot >> 3) + 8 * (slot & 7); }

// Twiddle table for this variant: tww[(r-1) * 32 + lane] = exp(-2 pi i lane r / 2048), r = 1..63 (host: scn_api.cu).
#ifndef SCN_WPT_MINCTAS
#define SCN_WPT_MINCTAS 4
#endif


template <bool DC>
__global__ void __launch_bounds__(32 * kWptWarpsPerCta, SCN_WPT_MINCTAS)
spectrum_sense_wpt_kernel(const KernelParams p) {
  constexpr int N = kWptN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float2* tile = reinterpret_cast<float2*>(smem_raw) + size_t(warp) * wpt_tile_elems();
  const uint32_t half = N / 2;
  const uint32_t gw = blockIdx.x * kWptWarpsPerCta + warp;          // global warp id
  const uint32_t nw = gridDim.x * kWptWarpsPerCta;

  // candidate test of process.cpp:46-53 for FFT bin j (evaluated lazily: only slots with a raw hit)
  auto is_candidate = [&](uint32_t j) -> bool {
    const uint32_t i = j ^ half;
    return !(j < p.dc_ignore || (N - j) < p.dc_ignore) && !(i < (half - p.use_window) || i > (half + p.use_window));
  };
  // Transform indices: the first two of a warp are static (gw, gw + nw); every later one comes from the launch's
  // work counter (WorkQueue, scn_kernel.cuh), fetched one transform ahead of its loads so the atomic's latency
  // hides behind an FFT.  A warp that starts late or runs slowly (another kernel holding its SM) just takes fewer.
  WorkQueue wq(p.work, nw, p.always_zero);
  uint32_t raw[32];                                                  // row r: samples 2*lane, 2*lane+1 (+ 64 r)
  uint32_t s_cur = gw, s_next = gw + nw;
  if (s_cur >= p.n_spectra) { if (lane == 0) wq.retire(); return; }

  while (true) {
    // raw samples of this transform.  They are NOT held in registers one transform ahead (32 registers per lane that
    // the FFT's schedule can use better: +1 %, profiles/r02zm_wpt_l2_prefetch_ab.txt); instead the copy engine was asked
    // one transform ago to pull these 4 KB into L2 (see below), so the loads are L2 hits.
    {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(p.raw + size_t(s_cur) * N * 2) + lane;
#pragma unroll
      for (int r = 0; r < 32; r++) raw[r] = ldg_stream(src + 32 * r);
    }
    // ---- DC (warp-local), convert + window ---------------------------------------------------------------
    float2 negc = make_float2(-(kMagic + 128.0f), -(kMagic + 128.0f));
    if constexpr (DC) {
      int si = 0, sq = 0;
#pragma unroll
      for (int r = 0; r < 32; r++) { si = __dp4a(int(raw[r]), 0x00010001, si); sq = __dp4a(int(raw[r]), 0x01000100, sq); }
      si = __reduce_add_sync(0xffffffffu, si);
      sq = __reduce_add_sync(0xffffffffu, sq);
      const int dci = int(unsigned(si) >> 11), dcq = int(unsigned(sq) >> 11);   // unsigned division by N (utility.cpp:49-50)
      // |dc| <= 2^21 always holds for int8 sums over 2048 samples except through the unsigned quirk, where
      // dc < 2^32 / 2048 = 2^21: the magic-number path is exact in every case.
      negc = make_float2(-(kMagic + 128.0f + float(dci)), -(kMagic + 128.0f + float(dcq)));
    }
    float2 v[64];
#pragma unroll
    for (int r = 0; r < 32; r++) {
      const float2 w2 = __ldg(reinterpret_cast<const float2*>(p.window) + lane + 32 * r);   // taps 2*lane, 2*lane+1 (+64 r)
      const uint32_t x = raw[r] ^ 0x80808080u;
      const float2 a = make_float2(__uint_as_float(__byte_perm(x, kMagicBits, 0x7650)),
                                   __uint_as_float(__byte_perm(x, kMagicBits, 0x7651)));
      const float2 b = make_float2(__uint_as_float(__byte_perm(x, kMagicBits, 0x7652)),
                                   __uint_as_float(__byte_perm(x, kMagicBits, 0x7653)));
      v[r] = __fmul2_rn(__fadd2_rn(a, negc), make_float2(w2.x, w2.x));          // column 0, row r
      v[32 + r] = __fmul2_rn(__fadd2_rn(b, negc), make_float2(w2.y, w2.y));     // column 1, row r
    }
    // ---- next transform: one bulk L2 prefetch of its 4 KB (UBLKPF.L2, no LSU work, no registers); the index after
    // it is requested from the work counter ------------------------------------------------------------------------
    const bool has_next = s_next < p.n_spectra;
    uint32_t ticket = 0;
    if (has_next && lane == 0) {
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.raw + size_t(s_next) * N * 2), "r"(uint32_t(N * 2)) : "memory");
      ticket = wq.take();
    }

    // ---- pass 0: radix-32 on both columns; scatter (Stockham: butterfly j = 2 lane + c -> 32 j + q) ------------
    dft32_inplace<0>(v);
    dft32_inplace<32>(v);
    __syncwarp();                                  // previous gather of this tile is complete
    {
      float2* base = tile + 65 * lane;             // 64 lane + lane padding
#pragma unroll
      for (int c = 0; c < 2; c++)
#pragma unroll
        for (int s = 0; s < 32; s++) base[32 * c + dft32_out_index(s)] = v[32 * c + s];
    }
    __syncwarp();
    // ---- pass 1: gather lane + 32 r, twiddle W_2048^(lane r), radix-64 -------------------------------------------
#pragma unroll
    for (int r = 0; r < 64; r++) v[r] = tile[lane + 32 * r + (r >> 1)];
    {
      const float2* tw = p.twiddles + lane;
      // 14 table values + 49 single products; all 63 from the (L1-resident, 16 KB) table measured 0.6 % slower
      float2 wb[8];                                // w^1 .. w^7
#pragma unroll
      for (int b = 1; b < 8; b++) { wb[b] = __ldg(tw + (b - 1) * 32); v[b] = cmul(v[b], wb[b]); }
#pragma unroll
      for (int a = 1; a < 8; a++) {
        const float2 wa = __ldg(tw + (8 * a - 1) * 32);               // w^(8a)
        v[8 * a] = cmul(v[8 * a], wa);
#pragma unroll
        for (int b = 1; b < 8; b++) v[8 * a + b] = cmul(v[8 * a + b], cmul(wa, wb[b]));
      }
    }
    dft64_inplace(v);

    // ---- power, dB, spectrum out, detection ---------------------------------------------------------------------
    __syncwarp();                                  // gather done: the tile is free to stash the dB of raw hits
    float* out = p.spectra ? p.spectra + size_t(s_cur) * N + lane : nullptr;
    float* stash = reinterpret_cast<float*>(tile);
    uint32_t hb_lo = 0, hb_hi = 0;                 // hit bits by slot s
#pragma unroll
    for (int s = 0; s < 64; s++) {
      const float2 sq2 = __fmul2_rn(v[s], v[s]);
      const float db = kDbPerLog2 * __log2f(__fadd_rn(sq2.x, sq2.y));
      if (out) out[32 * dft64_out_index(s)] = db;
      if (db > p.threshold) {                      // strict >, NaN never hits (process.cpp:54); rare
        stash[lane + 32 * s] = db;
        if (s < 32) hb_lo |= 1u << s; else hb_hi |= 1u << (s - 32);
      }
    }
    // (a two-level test -- running maximum per group of 8 slots, per-slot compares only in a group above the threshold
    //  -- executes ~80 fewer instructions per transform and measured 1.2 % SLOWER at the bench's hit rate; removed)
    uint32_t w_lo = 0, w_hi = 0;                   // mask words `lane` and `lane + 32` of this spectrum
    uint32_t total = 0;
    uint32_t any_lo = __reduce_or_sync(0xffffffffu, hb_lo), any_hi = __reduce_or_sync(0xffffffffu, hb_hi);
    if ((any_lo | any_hi) != 0u) {
      // word of output q is q ^ 32: bins lane + 32 q <-> shifted index (lane + 32 q) ^ 1024
      uint32_t rem_lo = any_lo, rem_hi = any_hi;
      while ((rem_lo | rem_hi) != 0u) {            // warp-uniform loop over the slots that have a hit
        int s;
        if (rem_lo) { s = __ffs(rem_lo) - 1; rem_lo &= rem_lo - 1; } else { s = 32 + __ffs(rem_hi) - 1; rem_hi &= rem_hi - 1; }
        const int qq = (s >> 3) + 8 * (s & 7);
        uint32_t mine = (s < 32) ? (hb_lo >> s) & 1u : (hb_hi >> (s - 32)) & 1u;
        if (mine && !is_candidate(uint32_t(lane) + 32u * qq)) {        // out of band / DC hole: not a hit
          mine = 0u;
          if (s < 32) hb_lo &= ~(1u << s); else hb_hi &= ~(1u << (s - 32));
        }
        const uint32_t b = __ballot_sync(0xffffffffu, mine);
        const int word = qq ^ 32;
        if (lane == (word & 31)) { if (word < 32) w_lo = b; else w_hi = b; }
      }
      total = __reduce_add_sync(0xffffffffu, __popc(w_lo) + __popc(w_hi));
      if (p.hits != nullptr) {
        // exclusive prefix over the 64 words in word order: words 0..31 are w_lo of lanes 0..31, 32..63 w_hi
        uint32_t inc_lo = __popc(w_lo), inc_hi = __popc(w_hi);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t a = __shfl_up_sync(0xffffffffu, inc_lo, o), c = __shfl_up_sync(0xffffffffu, inc_hi, o);
          if (lane >= o) { inc_lo += a; inc_hi += c; }
        }
        const uint32_t sum_lo = __shfl_sync(0xffffffffu, inc_lo, 31);
        const uint32_t ex_lo = inc_lo - __popc(w_lo), ex_hi = sum_lo + inc_hi - __popc(w_hi);
        rem_lo = any_lo; rem_hi = any_hi;
        while ((rem_lo | rem_hi) != 0u) {
          int s;
          if (rem_lo) { s = __ffs(rem_lo) - 1; rem_lo &= rem_lo - 1; } else { s = 32 + __ffs(rem_hi) - 1; rem_hi &= rem_hi - 1; }
          const int q = (s >> 3) + 8 * (s & 7);
          const int word = q ^ 32;
          const uint32_t before = __shfl_sync(0xffffffffu, word < 32 ? ex_lo : ex_hi, word & 31);
          const uint32_t wbits = __shfl_sync(0xffffffffu, word < 32 ? w_lo : w_hi, word & 31);
          const uint32_t mine = (s < 32) ? (hb_lo >> s) & 1u : (hb_hi >> (s - 32)) & 1u;
          if (mine) {
            const uint32_t rank = before + __popc(wbits & ((1u << lane) - 1u));
            if (rank < p.hit_cap) {
              scn_hit h;
              h.bin = (uint32_t(lane) + 32u * q) ^ half;
              h.power_db = stash[lane + 32 * s];
              p.hits[size_t(s_cur) * p.hit_cap + rank] = h;
            }
          }
        }
      }
    }
    if (p.masks != nullptr) {
      p.masks[size_t(s_cur) * 64 + lane] = w_lo;
      p.masks[size_t(s_cur) * 64 + 32 + lane] = w_hi;
    }
    if (p.counts != nullptr && lane == 0) p.counts[s_cur] = total;

    if (!has_next) break;
    s_cur = s_next;
    // keep the broadcast of the ticket HERE, behind the epilogue's stores (see WorkQueue::take)
    asm volatile("" : "+r"(ticket) : : "memory");
    s_next = __shfl_sync(0xffffffffu, ticket, 0) + wq.first_dynamic();
  }
  if (lane == 0) wq.retire();
}

}  // namespace scn
