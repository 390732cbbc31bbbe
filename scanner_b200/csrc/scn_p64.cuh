// scn_p64.cuh -- 64-points-per-thread variants of the fused kernel for N = 4096 and N = 8192
// (8192 is the reference's default --count, scan.cpp:85; BASELINE.json configs[2] is N = 4096 with
// 64-FFT averaging, configs[3] is fp32 IQ at N = 8192).
//
// T = N/64 threads own one transform (2 or 4 warps per CTA):
//   N = 4096 = 64 x 64    : radix-64 on column t, ONE exchange, radix-64 with twiddles W_4096^(t r)
//   N = 8192 = 64 x 64 x 2: the same two passes (twiddles W_4096^((t mod 64) r)); threads t and t ^ 64 then hold the
//                           even- / odd-column halves Z_0, Z_1 of the same 128-point row, so the last radix-2 is a
//                           PAIRWISE swap of 32 points each through 32 KB of the tile, then 32 butterflies
//                           Z_0[k] +- W_8192^(kk + 64 k) Z_1[k] per thread
// In both cases a warp's lanes hold 32 consecutive bins in every register slot: coalesced stores, warp-owned mask words.  Against the 16-points-per-thread plans
// ([16,16,16] / [2,16,16,16]) this saves one exchange, keeps CTAs small (several independent CTAs per SM
// drift apart, so one CTA's exchange overlaps another's butterflies) and runs the radix-64 stages on
// packed fp32x2 math with W64 twiddles as immediates (scn_wpt.cuh).
//
// Registers: the 64 points take 128, K > 1 adds 64 accumulators -- nothing is left for a register
// prefetch.  For the integer kinds the NEXT raw buffer (8..32 KB) is therefore prefetched by ONE bulk
// async copy (TMA: cp.async.bulk + mbarrier complete_tx, issued by one thread right after the first
// exchange barrier) into a shared-memory staging buffer; threads pick their column out of it with LDS,
// and the next buffer's DC sums (packed dot products over the staged words, REDUX) ride on a barrier
// that is needed anyway.  fp32 IQ (32 / 64 KB per transform) loads its column directly.
// Same arithmetic contract and tests as the generic family.
#pragma once
#include "scn_wpt.cuh"

namespace scn {

template <int LOG2N>
struct P64Geometry {
  static_assert(LOG2N == 12 || LOG2N == 13, "N = 4096 or 8192");
  static constexpr int N = 1 << LOG2N;
  static constexpr int T = N / 64;                       // threads per transform == threads per CTA
  static constexpr int WARPS = T / 32;
  static constexpr int WORDS = N / 32;
  static constexpr int TILE_ELEMS = N + N / 64;          // one float2 of padding per 64
  static constexpr size_t kTileBytes = sizeof(float2) * size_t(TILE_ELEMS);
  static constexpr size_t kMaskBytes = sizeof(uint32_t) * WORDS * 2;          // ping-pong by spectrum parity
  static constexpr size_t kRedBytes = sizeof(int32_t) * 2 * WARPS * 2;        // DC partials, ping-pong
  static constexpr size_t kWorkBytes = 16;                                    // next-spectrum index, ping-pong
  static constexpr size_t kStageOffset = kTileBytes + kMaskBytes + kWorkBytes + kRedBytes + 16;   // + mbarrier slot
  static_assert(kStageOffset % 16 == 0, "staging buffer must be 16-byte aligned for the bulk copy");
};
template <int LOG2N, int KIND>
constexpr size_t p64_smem_bytes() {
  using G = P64Geometry<LOG2N>;
  return KIND == SCN_KIND_FLOAT_COMPLEX ? G::kStageOffset      // tile | masks | work | (red) | mbarrier
                                        : G::kStageOffset + size_t(G::N) * KindTraits<KIND>::kBytes;
}
// twiddle tables (host: scn_api.cu, layout 2): twA[(r-1)*64 + k] = exp(-2 pi i k r / 4096), r = 1..63, k < 64;
//                                             twB[m*64 + kk]    = exp(-2 pi i (kk + 64 m) / 8192), m < 32, kk < 64
// (threads t >= 64 need W^(kk + 64 (m + 32)) = the same entry times -i, applied for free in the butterfly -- with the
//  mirrored window taps the per-CTA table working set is 39 KB instead of 78 KB, inside the L1 that two 66 KB CTAs
//  leave; ncu had the tables hitting L1 only 34 % of the time before)
constexpr int kP64TwAElems = 63 * 64;
constexpr int kP64TwBElems = 32 * 64;

#ifndef SCN_P64_MINCTAS
#define SCN_P64_MINCTAS 2
#endif
template <int LOG2N>
constexpr int p64_min_ctas() { return LOG2N == 12 ? 2 * SCN_P64_MINCTAS : SCN_P64_MINCTAS; }   // 255 registers

template <int LOG2N, int KIND, bool DC, bool AVG>
__global__ void __launch_bounds__(P64Geometry<LOG2N>::T, p64_min_ctas<LOG2N>())
spectrum_sense_p64_kernel(const KernelParams p) {
  using G = P64Geometry<LOG2N>;
  constexpr int N = G::N, T = G::T;
  constexpr bool kStaged = KIND != SCN_KIND_FLOAT_COMPLEX;
  constexpr bool kDC = DC && kStaged;
  constexpr uint32_t kRawBytes = uint32_t(N) * KindTraits<KIND>::kBytes;
  constexpr int GSTRIDE = T + T / 64;                    // padded stride of the gather t + T r
  static_assert(KIND == SCN_KIND_FLOAT_COMPLEX || KIND == SCN_KIND_BYTE_COMPLEX || KIND == SCN_KIND_SHORT_COMPLEX,
                "interleaved kinds only");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  uint32_t* smask = reinterpret_cast<uint32_t*>(smem_raw + G::kTileBytes);                      // [2][WORDS]
  uint32_t* swork = reinterpret_cast<uint32_t*>(smem_raw + G::kTileBytes + G::kMaskBytes);     // [2]
  int32_t* sred = reinterpret_cast<int32_t*>(smem_raw + G::kTileBytes + G::kMaskBytes + G::kWorkBytes);   // [2][WARPS][2]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + G::kTileBytes + G::kMaskBytes + G::kWorkBytes + G::kRedBytes);
  const unsigned char* stage = smem_raw + G::kStageOffset;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t half = N / 2;
  const float2* twA = p.twiddles;
  const float2* twB = p.twiddles + kP64TwAElems;
  const uint32_t K = AVG ? p.averaging : 1u;
  uint32_t spar = 0, phase = 0, tpar = 0;
  int dci = 0, dcq = 0;

  // window tap of sample t + T r.  N = 8192 only (its tables overflow L1, N = 4096's do not and its K > 1 variant
  // has no register to spare): taps of the upper half come from the mirrored address when the table is symmetric.
  constexpr bool kMirror = LOG2N == 13;
  const float* wlo = p.window + t;
  const float* whi = (kMirror && p.win_mirror) ? p.window + (T - 1 - t) + 31 * T : p.window + t + 32 * T;
  const int wstep = (kMirror && p.win_mirror) ? -T : T;
  auto wtap = [&](int r) -> float {
    if constexpr (kMirror) return r < 32 ? __ldg(wlo + T * r) : __ldg(whi + wstep * (r - 32));
    else return __ldg(wlo + T * r);
  };

  auto is_candidate = [&](uint32_t j) -> bool {          // process.cpp:46-53
    const uint32_t i = j ^ half;
    return !(j < p.dc_ignore || (N - j) < p.dc_ignore) && !(i < (half - p.use_window) || i > (half + p.use_window));
  };
  // FFT bin held in slot x of v[] after the last pass.  N = 4096: t + 64 q(x).  N = 8192 (row kk = t mod 64,
  // column parity hi = t / 64): slots 0..31 hold X[kk + 64 (m + 32 hi)], slots 32..63 the same + 4096.
  // Either way the lanes of a warp hold 32 CONSECUTIVE bins in a slot, i.e. exactly one mask word.
  const uint32_t bin_base = LOG2N == 13 ? uint32_t(t & 63) + 2048u * uint32_t(t >> 6) : uint32_t(t);
  // bin of slot x relative to bin_base (a compile-time constant for a compile-time x)
  auto bin_rel = [](int x) -> uint32_t {
    return LOG2N == 13 ? 64u * uint32_t(x & 31) + 4096u * uint32_t(x >> 5) : uint32_t(T) * uint32_t(dft64_out_index(x));
  };
  auto bin_of = [&](int x) -> uint32_t { return bin_base + bin_rel(x); };
  // mask word (shifted index >> 5) of slot x for this warp: lane 0's bin is bin_base - lane
  const uint32_t word_base = (bin_base - uint32_t(lane)) >> 5;
  auto word_of = [&](int x) -> uint32_t { return (word_base + (bin_rel(x) >> 5)) ^ (half >> 5); };

  // int32 sums of I and Q over the staged buffer (utility.cpp:44-48): this thread's words t + T i
  auto staged_sums = [&](int& si, int& sq) {
    si = 0; sq = 0;
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(stage);
#pragma unroll 8
    for (int i = 0; i < int(kRawBytes / 4 / T); i++) {
      const int wv = int(w32[t + T * i]);
      if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) { si = __dp4a(wv, 0x00010001, si); sq = __dp4a(wv, 0x01000100, sq); }
      else { si = __dp2a_lo(wv, 0x00000001, si); sq = __dp2a_lo(wv, 0x00000100, sq); }
    }
    si = __reduce_add_sync(0xffffffffu, si);
    sq = __reduce_add_sync(0xffffffffu, sq);
  };
  auto finish_dc = [&](const int32_t* red, int& odci, int& odcq) {
    int si = 0, sq = 0;
#pragma unroll
    for (int w = 0; w < G::WARPS; w++) { si += red[2 * w]; sq += red[2 * w + 1]; }
    odci = int(unsigned(si) >> LOG2N);       // unsigned division by N (utility.cpp:49-50)
    odcq = int(unsigned(sq) >> LOG2N);
  };

  // ---- tile stream of this CTA: (spectrum, k), k = 0..K-1; spectra blockIdx.x and blockIdx.x + gridDim.x are
  // static, later ones come from the launch's work counter (WorkQueue, scn_kernel.cuh) ----
  WorkQueue wq(p.work, gridDim.x, p.always_zero);
  uint32_t s = blockIdx.x, s_after = blockIdx.x + gridDim.x, s_after2 = 0, k = 0, ticket = 0;
  if (s >= p.n_spectra) {
    if (t == 0) wq.retire();
    return;
  }
  if constexpr (kStaged) {
    if (t == 0) {
      mbar_init(bar, 1);
      mbar_expect_tx(bar, kRawBytes);
      bulk_g2s(const_cast<unsigned char*>(stage), p.raw + size_t(s) * K * kRawBytes, kRawBytes, bar);
    }
    __syncthreads();
    if constexpr (kDC) {
      mbar_wait(bar, 0);
      int si, sq;
      staged_sums(si, sq);
      if (lane == 0) { sred[2 * warp] = si; sred[2 * warp + 1] = sq; }
      __syncthreads();
      finish_dc(sred, dci, dcq);
      tpar = 1;
    }
  }

  // fp32 IQ, K = 1: the imaginary halves of v[] are dead once the powers are formed, so the first 32 points of the
  // NEXT buffer are loaded into those 64 registers before the epilogue (dB, stores, detection) and their latency
  // hides behind it; the other 32 points load at the top of the next tile as before.
#ifndef SCN_P64_PREFETCH
#define SCN_P64_PREFETCH 1
#endif
  // fp32 IQ, SCN_P64_LAND: the exchange tile doubles as the landing zone of ONE TMA bulk copy of the next raw
  // buffer (32 / 64 KB), issued by one thread as soon as every thread has finished its last gather from the tile
  // (N = 4096: before the second radix-64, so the copy overlaps half of the transform and the whole epilogue;
  // N = 8192: before the epilogue).  Threads then read their column with LDS.  The raw stream no longer passes
  // through L1 (the window / twiddle tables stay resident) and no registers are tied up by a prefetch.
#ifndef SCN_P64_LAND
#define SCN_P64_LAND 1
#endif
  constexpr bool kLand = SCN_P64_LAND && !kStaged;
  // N = 8192 lands in two halves (two arrivals per mbarrier phase): samples 4096.. as soon as the last full-tile
  // gather is done -- the pairwise swap of the last radix-2 only needs the first 32 KB of the tile -- and samples
  // 0..4095 after the swap, under the epilogue.
  constexpr uint32_t kLandParts = (LOG2N == 13) ? 2u : 1u;
  constexpr uint32_t kPartBytes = kRawBytes / kLandParts;
  auto land_part = [&](size_t buffer, uint32_t part) {       // one thread; the tile region is no longer read
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, kPartBytes);
    bulk_g2s(reinterpret_cast<unsigned char*>(tile) + size_t(part) * kPartBytes,
             p.raw + buffer * kRawBytes + size_t(part) * kPartBytes, kPartBytes, bar);
  };
  if constexpr (kLand) {
    if (t == 0) {
      mbar_init(bar, kLandParts);
      for (uint32_t part = 0; part < kLandParts; part++) land_part(size_t(s) * K, part);
    }
    __syncthreads();
  }
  constexpr bool kPrefetch = SCN_P64_PREFETCH && !kStaged && !AVG && !kLand;
#ifndef SCN_P64_PREFETCH_13
#define SCN_P64_PREFETCH_13 16
#endif
  constexpr int kPre = !kPrefetch ? 0 : (LOG2N == 12 ? 32 : SCN_P64_PREFETCH_13);   // points prefetched
  float2 nxt[kPrefetch ? kPre : 1];
  if constexpr (kPrefetch) {
    const float2* src = reinterpret_cast<const float2*>(p.raw) + size_t(s) * N + t;
#pragma unroll
    for (int r = 0; r < kPre; r++) nxt[r] = ldg_stream(src + T * r);
  }

  float acc[AVG ? 64 : 1];
  while (true) {
    uint32_t ns = s, nk = k + 1;
    if (nk == K) { nk = 0; ns = s_after; }
    const bool has_next = ns < p.n_spectra;
    if (k == 0 && t == 0) ticket = wq.take();          // the spectrum after s_after: requested now, parked in a
                                                       // register, handed to the CTA behind the epilogue barrier
    const bool epilogue_tile = (k == K - 1);
    const size_t buf_index = size_t(s) * K + k;

    // ---- load / convert + window ----------------------------------------------------------------------------
    float2 v[64];
    if constexpr (kLand) {
      mbar_wait(bar, phase);
      phase ^= 1u;
      const float2* land = tile + t;
#pragma unroll
      for (int r = 0; r < 64; r++) v[r] = land[T * r];
#pragma unroll
      for (int r = 0; r < 64; r++) {
        const float w = wtap(r);
        v[r] = __fmul2_rn(v[r], make_float2(w, w));
      }
      __syncthreads();                                 // every column is in registers: the tile may be scattered into
    } else if constexpr (!kStaged) {
      const float2* src = reinterpret_cast<const float2*>(p.raw) + buf_index * N + t;
#pragma unroll
      for (int r = 0; r < 64; r++) v[r] = (kPrefetch && r < kPre) ? nxt[r] : ldg_stream(src + T * r);
#pragma unroll
      for (int r = 0; r < 64; r++) {
        const float w = wtap(r);
        v[r] = __fmul2_rn(v[r], make_float2(w, w));
      }
    } else {
      // convert + scale + window from the staged buffer (utility.cpp:52-55, process.cpp:28-34): magic-number
      // placement (PRMT), one exact FADD2 for magic + bias + dc, one FMUL2 for the pre-scaled window tap;
      // |dc| <= 2^32/N <= 2^20 even through the unsigned-division quirk, so this path is always exact.
      mbar_wait(bar, phase);
      phase ^= 1u;
      constexpr float kOff = (KIND == SCN_KIND_BYTE_COMPLEX) ? 128.0f : 32768.0f;
      const float2 negc = make_float2(-(kMagic + kOff + float(dci)), -(kMagic + kOff + float(dcq)));
#pragma unroll
      for (int r = 0; r < 64; r++) {
        uint32_t bi, bq;
        if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
          const uint32_t x = uint32_t(reinterpret_cast<const unsigned short*>(stage)[t + T * r]) ^ 0x8080u;
          bi = __byte_perm(x, kMagicBits, 0x7650);
          bq = __byte_perm(x, kMagicBits, 0x7651);
        } else {
          const uint32_t x = reinterpret_cast<const uint32_t*>(stage)[t + T * r] ^ 0x80008000u;
          bi = __byte_perm(x, kMagicBits, 0x7610);
          bq = __byte_perm(x, kMagicBits, 0x7632);
        }
        const float w = wtap(r);
        v[r] = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(bi), __uint_as_float(bq)), negc), make_float2(w, w));
      }
    }
    // ---- pass 0: radix-64 over r; scatter 64 t + q (padded 65 t + q) -----------------------------------------
    dft64_inplace(v);
    {
      float2* base = tile + 65 * t;
#pragma unroll
      for (int x = 0; x < 64; x++) base[dft64_out_index(x)] = v[x];
    }
    __syncthreads();
    if constexpr (kStaged) {
      // every thread has consumed the staged buffer (it passed the barrier after its conversion): refill it
      if (has_next && t == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, kRawBytes);
        bulk_g2s(const_cast<unsigned char*>(stage), p.raw + (size_t(ns) * K + nk) * kRawBytes, kRawBytes, bar);
      }
    }
    // ---- pass 1: gather t + T r, twiddle W_4096^(kk r), radix-64 --------------------------------------------------
    {
      const float2* base = tile + t + (t >> 6);
#pragma unroll
      for (int r = 0; r < 64; r++) v[r] = base[GSTRIDE * r];
    }
    if constexpr (kLand && LOG2N == 12) {
      __syncthreads();                                 // last gather done: the tile is free for the next raw buffer
      if (has_next && t == 0) land_part(size_t(ns) * K + nk, 0);
    }
    const int kk = t & 63;
    {
      const float2* tw = twA + kk;
      float2 wb[8];
#pragma unroll
      for (int b = 1; b < 8; b++) { wb[b] = __ldg(tw + (b - 1) * 64); v[b] = cmul(v[b], wb[b]); }
#pragma unroll
      for (int a = 1; a < 8; a++) {
        const float2 wa = __ldg(tw + (8 * a - 1) * 64);
        v[8 * a] = cmul(v[8 * a], wa);
#pragma unroll
        for (int b = 1; b < 8; b++) v[8 * a + b] = cmul(v[8 * a + b], cmul(wa, wb[b]));
      }
    }
    dft64_inplace(v);
    if constexpr (LOG2N == 13) {
      // ---- pass 2: the 128-point row kk = (64-point DFT of its even columns, thread kk) + (odd columns, thread
      // kk + 64): X[kk + 64 k] = Z_0[k] + W_8192^(kk + 64 k) Z_1[k], X[.. + 4096] = Z_0[k] - (same), k < 64.
      // Thread kk takes k < 32, thread kk + 64 takes k >= 32: each sends the partner the 32 points it needs
      // through the first 32 KB of the tile, layout [m][thread] (conflict-free both ways). ------------------------
      const uint32_t hi = uint32_t(t) >> 6;            // warp-uniform
      __syncthreads();                                 // every thread has finished its gather of exchange 1
      if constexpr (kLand) {
        if (has_next && t == 0) land_part(size_t(ns) * K + nk, 1);   // upper 32 KB of the tile is free from here on
      }
      {
        float2* dst = tile + (t ^ 64);
        if (hi == 0) {
#pragma unroll
          for (int x = 0; x < 64; x++) if (dft64_out_index(x) >= 32) dst[T * (dft64_out_index(x) - 32)] = v[x];   // Z_0[32 + m]
        } else {
#pragma unroll
          for (int x = 0; x < 64; x++) if (dft64_out_index(x) < 32) dst[T * dft64_out_index(x)] = v[x];           // Z_1[m]
        }
      }
      __syncthreads();
      {
        const float2* src = tile + t;
        const float2* tw = twB + kk;                   // W_8192^(kk + 64 m); threads t >= 64 need it times -i
        float2 o[64];
        if (hi == 0) {
#pragma unroll
          for (int m = 0; m < 32; m++) {
            const float2 e = v[8 * (m & 7) + (m >> 3)];                  // own Z_0[m]
            const float2 b = cmul(src[T * m], __ldg(tw + 64 * m));       // W Z_1[m]
            o[m] = cadd(e, b);
            o[32 + m] = csub(e, b);
          }
        } else {
#pragma unroll
          for (int m = 0; m < 32; m++) {
            const float2 b = cmul(v[8 * (m & 7) + (m >> 3) + 4], __ldg(tw + 64 * m));   // own Z_1[32 + m] times W^(kk + 64 m)
            const float2 e = src[T * m];                                  // Z_0[32 + m]
            o[m] = add_mi(e, b);                                          // ... times -i = W^(kk + 64 (m + 32))
            o[32 + m] = sub_mi(e, b);
          }
        }
#pragma unroll
        for (int x = 0; x < 64; x++) v[x] = o[x];
      }
      if constexpr (kLand) {
        __syncthreads();                               // swap region read: the lower 32 KB may land now
        if (has_next && t == 0) land_part(size_t(ns) * K + nk, 0);
      }
    }

    // ---- power, K-averaging (fp32, buffer order) -----------------------------------------------------
#pragma unroll
    for (int x = 0; x < 64; x++) {
      const float2 sq2 = __fmul2_rn(v[x], v[x]);       // fl(re*re), fl(im*im): no FMA contraction
      float pw = __fadd_rn(sq2.x, sq2.y);
      if constexpr (AVG) pw = acc[x] = (k == 0) ? pw : __fadd_rn(acc[x], pw);
      v[x].x = pw;
    }
    if constexpr (kPrefetch) {
      if (has_next) {
        const float2* src = reinterpret_cast<const float2*>(p.raw) + size_t(ns) * N + t;
#pragma unroll
        for (int r = 0; r < kPre; r++) nxt[r] = ldg_stream(src + T * r);
      }
    }

    if (!epilogue_tile) {
      // K > 1, not the last buffer of the spectrum: the next buffer's DC reduction, and (every variant) the
      // guarantee that no thread scatters the next tile before all have gathered this one, need one barrier
      if constexpr (kDC) {
        if (has_next) {
          mbar_wait(bar, phase);
          int si, sq;
          staged_sums(si, sq);
          if (lane == 0) { sred[2 * G::WARPS * tpar + 2 * warp] = si; sred[2 * G::WARPS * tpar + 2 * warp + 1] = sq; }
        }
      }
      __syncthreads();
      if constexpr (kDC) {
        if (has_next) { finish_dc(sred + 2 * G::WARPS * tpar, dci, dcq); tpar ^= 1u; }
      }
    } else {
      // ---- dB, spectrum out, detection (slot x <-> FFT bin bin_of(x)) ----------------------------------------------
      uint32_t* sm = smask + spar * G::WORDS;
      float* out = p.spectra ? p.spectra + size_t(s) * N + bin_base : nullptr;
      bool anyraw = false;
#pragma unroll
      for (int x = 0; x < 64; x++) {
        const float pbar = AVG ? __fmul_rn(v[x].x, p.inv_averaging) : v[x].x;
        const float db = kDbPerLog2 * __log2f(pbar);
        v[x].x = db;
        if (out) out[bin_rel(x)] = db;
        anyraw = anyraw || (db > p.threshold);          // strict >, NaN never hits (process.cpp:54)
      }
      // this warp owns mask words word_of(x), x = 0..63: zero them (two per lane), then fill on demand
      sm[word_of(lane)] = 0u;
      sm[word_of(lane + 32)] = 0u;
      uint32_t hb_lo = 0, hb_hi = 0;                   // hit bits by slot x
      const bool warp_any = __any_sync(0xffffffffu, anyraw);
      if (warp_any) {
#pragma unroll
        for (int x = 0; x < 64; x++) {
          if (v[x].x > p.threshold) { if (x < 32) hb_lo |= 1u << x; else hb_hi |= 1u << (x - 32); }
        }
        // candidate test only for the (few) raw hits of this lane
        for (uint32_t rest = hb_lo; rest; rest &= rest - 1) {
          const int x = __ffs(rest) - 1;
          if (!is_candidate(bin_of(x))) hb_lo &= ~(1u << x);
        }
        for (uint32_t rest = hb_hi; rest; rest &= rest - 1) {
          const int x = __ffs(rest) - 1;
          if (!is_candidate(bin_of(x + 32))) hb_hi &= ~(1u << x);
        }
        __syncwarp();
        uint32_t rem_lo = __reduce_or_sync(0xffffffffu, hb_lo), rem_hi = __reduce_or_sync(0xffffffffu, hb_hi);
        while ((rem_lo | rem_hi) != 0u) {
          int x;
          if (rem_lo) { x = __ffs(rem_lo) - 1; rem_lo &= rem_lo - 1; } else { x = 32 + __ffs(rem_hi) - 1; rem_hi &= rem_hi - 1; }
          const uint32_t mine = (x < 32) ? (hb_lo >> x) & 1u : (hb_hi >> (x - 32)) & 1u;
          const uint32_t b = __ballot_sync(0xffffffffu, mine);
          if (lane == 0) sm[word_of(x)] = b;
        }
      }
      if constexpr (kDC) {
        if (has_next) {                                // next buffer's DC sums ride on this barrier
          mbar_wait(bar, phase);
          int si, sq;
          staged_sums(si, sq);
          if (lane == 0) { sred[2 * G::WARPS * tpar + 2 * warp] = si; sred[2 * G::WARPS * tpar + 2 * warp + 1] = sq; }
        }
      }
      if (t == 0) swork[spar] = ticket + wq.first_dynamic();
      __syncthreads();
      s_after2 = swork[spar];
      if constexpr (kDC) {
        if (has_next) { finish_dc(sred + 2 * G::WARPS * tpar, dci, dcq); tpar ^= 1u; }
      }
      // warp 0: mask words out (coalesced) + hit count
      if (warp == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int c = 0; c < G::WORDS; c += 32) {
          const uint32_t mw = sm[c + lane];
          if (p.masks != nullptr) p.masks[size_t(s) * G::WORDS + c + lane] = mw;
          total += __popc(mw);
        }
        total = __reduce_add_sync(0xffffffffu, total);
        if (lane == 0 && p.counts != nullptr) p.counts[s] = total;
      }
      // hit records in ascending shifted bin: rank = hits in earlier words + hits in lower lanes of my word
      if (p.hits != nullptr && warp_any) {
        const uint32_t any_lo = __reduce_or_sync(0xffffffffu, hb_lo), any_hi = __reduce_or_sync(0xffffffffu, hb_hi);
#pragma unroll
        for (int x = 0; x < 64; x++) {
          const bool any_x = (x < 32) ? ((any_lo >> x) & 1u) : ((any_hi >> (x - 32)) & 1u);
          if (any_x) {                                 // warp-uniform, rare
            const uint32_t word = word_of(x);
            uint32_t before = 0;
            for (uint32_t y = lane; y < word; y += 32) before += __popc(sm[y]);
            before = __reduce_add_sync(0xffffffffu, before);
            const uint32_t mine = (x < 32) ? (hb_lo >> x) & 1u : (hb_hi >> (x - 32)) & 1u;
            const uint32_t b = __ballot_sync(0xffffffffu, mine);
            if (mine) {
              const uint32_t rank = before + __popc(b & ((1u << lane) - 1u));
              if (rank < p.hit_cap) {
                scn_hit h;
                h.bin = bin_of(x) ^ half;
                h.power_db = v[x].x;
                p.hits[size_t(s) * p.hit_cap + rank] = h;
              }
            }
          }
        }
      }
      spar ^= 1u;
      // (the next scatter into `tile` is safe: every thread passed the barrier above after its last gather;
      //  the mask words ping-pong by spectrum parity)
    }

    if (!has_next) break;
    if (nk == 0) s_after = s_after2;
    s = ns; k = nk;
  }
  if (t == 0) wq.retire();
}

}  // namespace scn
