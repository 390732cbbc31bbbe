// scn_p64.cuh -- 64-points-per-thread variant of the fused kernel for N = 8192, K = 1: fp32 IQ
// (BASELINE.json configs[3]: Airspy-style fp32 stream, 8192-pt FFT, threshold detect) and int8 / int16
// interleaved IQ (8192 is the reference's default --count, scan.cpp:85).
//
// Integer kinds: all 128 registers hold the points, so the NEXT transform's raw buffer (16 / 32 KB) is
// prefetched by ONE bulk async copy (TMA: cp.async.bulk + mbarrier complete_tx) into a shared-memory
// staging buffer while the current transform is in the FFT; threads then pick their column out of it
// with LDS.  The DC sums of the next transform (packed dot products over the staged words) ride on the
// epilogue barrier.  fp32 IQ (64 KB per transform: no room to stage) loads its column directly.
//
// 128 threads own one transform, N = 64 x 64 x 2:
//   pass 0: radix-64 in registers on column t (rows t + 128 r: coalesced 8-byte loads, window fused),
//   exchange, pass 1: radix-64 with twiddles W_4096^((t mod 64) r), exchange,
//   pass 2: 32 radix-2 butterflies with twiddles W_8192^(t + 128 c); outputs t + 128 q: coalesced stores.
// Two exchanges instead of the three of the 16-points-per-thread plan [2,16,16,16], 4 warps per CTA
// instead of 16 (two CTAs per SM drift apart, so one CTA's exchange overlaps the other's butterflies),
// and the radix-64 stages run on packed fp32x2 math with W64 twiddles as immediates (scn_wpt.cuh).
// Same arithmetic contract and tests as the generic family.
#pragma once
#include "scn_wpt.cuh"

namespace scn {

constexpr int kP64N = 8192;
constexpr int kP64Threads = 128;
__host__ __device__ constexpr int p64_tile_elems() { return kP64N + kP64N / 64; }
constexpr int kP64Words = kP64N / 32;
constexpr size_t kP64TileBytes = sizeof(float2) * size_t(p64_tile_elems());
constexpr size_t kP64MaskBytes = sizeof(uint32_t) * kP64Words * 2;
// layout: [tile][mask x2][dc partials 2 x 4 warps x 2][mbarrier][staging]
constexpr size_t kP64RedBytes = sizeof(int32_t) * 2 * 4 * 2;
constexpr size_t kP64StageOffset = kP64TileBytes + kP64MaskBytes + kP64RedBytes + 16;
template <int KIND> constexpr size_t p64_smem_bytes() {
  return KIND == SCN_KIND_FLOAT_COMPLEX ? kP64TileBytes + kP64MaskBytes
                                        : kP64StageOffset + size_t(kP64N) * KindTraits<KIND>::kBytes;
}
static_assert(kP64StageOffset % 16 == 0, "staging buffer must be 16-byte aligned for the bulk copy");
// twiddle tables (host: scn_api.cu): twA[(r-1)*64 + k] = exp(-2 pi i k r / 4096), r = 1..63, k < 64;
//                                    twB[c*128 + t]    = exp(-2 pi i (t + 128 c) / 8192), c < 32, t < 128
constexpr int kP64TwAElems = 63 * 64;
constexpr int kP64TwBElems = 32 * 128;

#ifndef SCN_P64_MINCTAS
#define SCN_P64_MINCTAS 2
#endif
template <int KIND, bool DC>
__global__ void __launch_bounds__(kP64Threads, SCN_P64_MINCTAS)
spectrum_sense_p64_kernel(const KernelParams p) {
  constexpr int N = kP64N, T = kP64Threads;
  constexpr bool kStaged = KIND != SCN_KIND_FLOAT_COMPLEX;
  constexpr bool kDC = DC && kStaged;
  constexpr uint32_t kRawBytes = uint32_t(N) * KindTraits<KIND>::kBytes;
  static_assert(KIND == SCN_KIND_FLOAT_COMPLEX || KIND == SCN_KIND_BYTE_COMPLEX || KIND == SCN_KIND_SHORT_COMPLEX,
                "interleaved kinds only");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tile = reinterpret_cast<float2*>(smem_raw);
  uint32_t* smask = reinterpret_cast<uint32_t*>(smem_raw + kP64TileBytes);   // [2][256]
  int32_t* sred = reinterpret_cast<int32_t*>(smem_raw + kP64TileBytes + kP64MaskBytes);   // [2][4][2]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kP64TileBytes + kP64MaskBytes + kP64RedBytes);
  const unsigned char* stage = smem_raw + kP64StageOffset;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t half = N / 2;
  const float2* twA = p.twiddles;
  const float2* twB = p.twiddles + kP64TwAElems;
  uint32_t spar = 0, phase = 0, tpar = 0;
  int dci = 0, dcq = 0;

  // int32 sums of I and Q over the staged buffer (utility.cpp:44-48): this thread's words t + 128 i
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
    const int si = red[0] + red[2] + red[4] + red[6], sq = red[1] + red[3] + red[5] + red[7];
    odci = int(unsigned(si) >> 13);          // unsigned division by N = 8192 (utility.cpp:49-50)
    odcq = int(unsigned(sq) >> 13);
  };

  if (blockIdx.x >= p.n_spectra) return;
  if constexpr (kStaged) {
    if (t == 0) {
      mbar_init(bar, 1);
      mbar_expect_tx(bar, kRawBytes);
      bulk_g2s(const_cast<unsigned char*>(stage), p.raw + size_t(blockIdx.x) * kRawBytes, kRawBytes, bar);
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

  auto is_candidate = [&](uint32_t j) -> bool {          // process.cpp:46-53
    const uint32_t i = j ^ half;
    return !(j < p.dc_ignore || (N - j) < p.dc_ignore) && !(i < (half - p.use_window) || i > (half + p.use_window));
  };

  for (uint32_t s = blockIdx.x; s < p.n_spectra; s += gridDim.x) {
    // ---- load + window (process.cpp:28-34): v[r] = x[t + 128 r] * w[t + 128 r] ---------------------------
    float2 v[64];
    const uint32_t s_next = s + gridDim.x;
    const bool has_next = s_next < p.n_spectra;
    if constexpr (!kStaged) {
      const float2* src = reinterpret_cast<const float2*>(p.raw) + size_t(s) * N + t;
#pragma unroll
      for (int r = 0; r < 64; r++) v[r] = __ldg(src + T * r);
#pragma unroll
      for (int r = 0; r < 64; r++) {
        const float w = __ldg(p.window + t + T * r);
        v[r] = __fmul2_rn(v[r], make_float2(w, w));
      }
    } else {
      // convert + scale + window from the staged buffer (utility.cpp:52-55, process.cpp:28-34): magic-number
      // placement (PRMT), one exact FADD2 for magic + bias + dc, one FMUL2 for the pre-scaled window tap;
      // |dc| < 2^19 for N = 8192 even through the unsigned-division quirk, so this path is always exact.
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
        const float w = __ldg(p.window + t + T * r);
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
        bulk_g2s(const_cast<unsigned char*>(stage), p.raw + size_t(s_next) * kRawBytes, kRawBytes, bar);
      }
    }
    // ---- pass 1: gather t + 128 r, twiddle W_4096^(k r), radix-64, scatter j0 + 64 q ---------------------------
    {
      const float2* base = tile + t + (t >> 6);
#pragma unroll
      for (int r = 0; r < 64; r++) v[r] = base[130 * r];            // 128 r + 2 r padding
    }
    const int k = t & 63;
    {
      const float2* tw = twA + k;
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
    __syncthreads();                                   // every thread has finished its gather of exchange 1
    {
      const int j0 = ((t - k) << 6) + k;
      float2* base = tile + j0 + (j0 >> 6);
#pragma unroll
      for (int x = 0; x < 64; x++) base[65 * dft64_out_index(x)] = v[x];
    }
    __syncthreads();
    // ---- pass 2: 32 radix-2 butterflies j = t + 128 c: inputs j and j + 4096, twiddle W_8192^j ----------------------
    {
      const float2* base = tile + t + (t >> 6);
#pragma unroll
      for (int c = 0; c < 32; c++) {
        const float2 a = base[130 * c];
        const float2 b = cmul(base[130 * c + 4096 + 64], __ldg(twB + c * T + t));
        v[c] = cadd(a, b);                             // bin t + 128 c
        v[32 + c] = csub(a, b);                        // bin t + 128 (c + 32)
      }
    }
    // ---- power, dB, spectrum out, detection (slot q <-> FFT bin t + 128 q) ---------------------------------------
    uint32_t* sm = smask + spar * kP64Words;
    float* out = p.spectra ? p.spectra + size_t(s) * N + t : nullptr;
    bool anyraw = false;
#pragma unroll
    for (int q = 0; q < 64; q++) {
      const float2 sq2 = __fmul2_rn(v[q], v[q]);
      const float db = kDbPerLog2 * __log2f(__fadd_rn(sq2.x, sq2.y));
      v[q].x = db;
      if (out) out[T * q] = db;
      anyraw = anyraw || (db > p.threshold);
    }
    // this warp owns mask words (warp + 4 q) ^ 128, q = 0..63: zero them (two per lane), then fill on demand
    sm[(warp + 4 * lane) ^ 128] = 0u;
    sm[(warp + 4 * (lane + 32)) ^ 128] = 0u;
    uint32_t hb_lo = 0, hb_hi = 0;
    const bool warp_any = __any_sync(0xffffffffu, anyraw);
    if (warp_any) {
#pragma unroll
      for (int q = 0; q < 64; q++) {
        if (v[q].x > p.threshold) { if (q < 32) hb_lo |= 1u << q; else hb_hi |= 1u << (q - 32); }
      }
      // candidate test only for the (few) raw hits of this lane
      for (uint32_t rest = hb_lo; rest; rest &= rest - 1) {
        const int q = __ffs(rest) - 1;
        if (!is_candidate(uint32_t(t) + T * q)) hb_lo &= ~(1u << q);
      }
      for (uint32_t rest = hb_hi; rest; rest &= rest - 1) {
        const int q = __ffs(rest) - 1;
        if (!is_candidate(uint32_t(t) + T * (q + 32))) hb_hi &= ~(1u << q);
      }
      __syncwarp();
      uint32_t rem_lo = __reduce_or_sync(0xffffffffu, hb_lo), rem_hi = __reduce_or_sync(0xffffffffu, hb_hi);
      while ((rem_lo | rem_hi) != 0u) {
        int q;
        if (rem_lo) { q = __ffs(rem_lo) - 1; rem_lo &= rem_lo - 1; } else { q = 32 + __ffs(rem_hi) - 1; rem_hi &= rem_hi - 1; }
        const uint32_t mine = (q < 32) ? (hb_lo >> q) & 1u : (hb_hi >> (q - 32)) & 1u;
        const uint32_t b = __ballot_sync(0xffffffffu, mine);
        if (lane == 0) sm[(warp + 4 * q) ^ 128] = b;
      }
    }
    if constexpr (kDC) {
      if (has_next) {                                  // next transform's DC sums ride on this barrier
        mbar_wait(bar, phase);
        int si, sq;
        staged_sums(si, sq);
        if (lane == 0) { sred[8 * tpar + 2 * warp] = si; sred[8 * tpar + 2 * warp + 1] = sq; }
      }
    }
    __syncthreads();
    if constexpr (kDC) {
      if (has_next) { finish_dc(sred + 8 * tpar, dci, dcq); tpar ^= 1u; }
    }
    // warp 0: mask words out (coalesced) + hit count
    if (warp == 0) {
      uint32_t total = 0;
#pragma unroll
      for (int c = 0; c < kP64Words; c += 32) {
        const uint32_t mw = sm[c + lane];
        if (p.masks != nullptr) p.masks[size_t(s) * kP64Words + c + lane] = mw;
        total += __popc(mw);
      }
      total = __reduce_add_sync(0xffffffffu, total);
      if (lane == 0 && p.counts != nullptr) p.counts[s] = total;
    }
    // hit records in ascending shifted bin: rank = hits in earlier words + hits in lower lanes of my word
    if (p.hits != nullptr && warp_any) {
      uint32_t rem_lo = __reduce_or_sync(0xffffffffu, hb_lo), rem_hi = __reduce_or_sync(0xffffffffu, hb_hi);
#pragma unroll
      for (int q = 0; q < 64; q++) {
        const bool any_q = (q < 32) ? ((rem_lo >> q) & 1u) : ((rem_hi >> (q - 32)) & 1u);
        if (any_q) {                                   // warp-uniform, rare
          const uint32_t word = (warp + 4 * q) ^ 128;
          uint32_t before = 0;
          for (uint32_t x = lane; x < word; x += 32) before += __popc(sm[x]);
          before = __reduce_add_sync(0xffffffffu, before);
          const uint32_t mine = (q < 32) ? (hb_lo >> q) & 1u : (hb_hi >> (q - 32)) & 1u;
          const uint32_t b = __ballot_sync(0xffffffffu, mine);
          if (mine) {
            const uint32_t rank = before + __popc(b & ((1u << lane) - 1u));
            if (rank < p.hit_cap) {
              scn_hit h;
              h.bin = (uint32_t(t) + T * q) ^ half;
              h.power_db = v[q].x;
              p.hits[size_t(s) * p.hit_cap + rank] = h;
            }
          }
        }
      }
    }
    spar ^= 1u;
    // (the next transform's scatter into `tile` is safe: every thread passed the barrier above after its
    //  pass-2 gather; the mask words ping-pong by spectrum parity)
  }
}

}  // namespace scn
