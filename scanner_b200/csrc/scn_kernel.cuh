// scn_kernel.cuh -- the fused spectrum-sense kernel (sm_100a).
//
// One persistent CTA walks "groups" of F spectra (F = transforms resident in one CTA); for each
// spectrum it streams K raw IQ buffers through
//   load -> [DC sum] -> convert+scale+window -> FFT -> |X|^2 -> accumulate
// entirely in registers/shared memory, then emits dB, the detection mask, the hit count and the
// compact hit records.  Each raw sample crosses HBM exactly once.
//
// Reference arithmetic restated (file:line under the reference tree):
//   conversion      utility.cpp:34-56 (int8), :58-84 (int16 interleaved), :9-32 (int16 split)
//   DC quirk        utility.cpp:49-50  (int32 /= uint32  => unsigned division)
//   window          process.cpp:28-34  (one fp32 multiply per component)
//   FFT             fft.cpp:20-25      (forward, unnormalised)
//   dB              utility.cpp:91-97  (10*log2(sqrt(re^2+im^2))/log2(10))
//   detection       process.cpp:46-61  (fftshift index, DC hole, used band, strict >)
//
// Software pipeline of one tile (= one raw buffer of one transform group):
//   convert raw(t) -> issue loads raw(t+1) -> pass 0 -> exchange -> ... -> [DC partial sums of
//   raw(t+1) into smem] -> last exchange barrier -> [dc(t+1) from smem] -> last pass -> power.
// The next tile's HBM latency hides behind this tile's FFT, and the DC reduction rides on an
// exchange barrier that is needed anyway.  Barriers per tile: (passes - 1) + 1 per spectrum.
#pragma once
#include "scn_fft.cuh"
#include "../../include/scanner_b200.h"

// Twiddle strategy of a radix-16 pass (15 factors per thread).  Two schemes survived the A/B runs on B200
// (profiles/README.md; plain table loads, factors hoisted into registers for the whole launch, and mixed schemes were
// measured and removed):
//   kTwProducts: six loads (w^1..4, w^8, w^12) + nine single products -- one rounding deep, as accurate as 15 table
//                loads and 4 % faster because the L1 pipe is a co-bottleneck; every size below 2^14
//   kTwTree:     one load + a 14-multiply product tree (1.5x the rms error); N = 2^14 only, whose 1024-thread CTAs
//                run at 64 registers and spill with anything wider (+10-13 %)
// Other knobs used to explore the occupancy / register trade (defaults = measured best):
#ifndef SCN_XBUFS_MAXLOG2
#define SCN_XBUFS_MAXLOG2 13   // ping-pong exchange tiles up to this size, single tile (two barriers) above
#endif
#ifndef SCN_WINREG_MAXLOG2
#define SCN_WINREG_MAXLOG2 13  // window taps live in registers up to this size, L1 loads per tile above
#endif

namespace scn {

struct KernelParams {
  const uint8_t* __restrict__ raw;       // n_spectra * K buffers
  const float* __restrict__ window;      // N taps, pre-scaled by 1/max for integer kinds
  const float2* __restrict__ twiddles;   // pass tables, see pass_twiddle()
  float* __restrict__ spectra;           // nullable, [n_spectra][N]
  uint32_t* __restrict__ masks;          // nullable, [n_spectra][N/32]
  uint32_t* __restrict__ counts;         // nullable, [n_spectra]
  scn_hit* __restrict__ hits;            // nullable, [n_spectra][hit_cap]
  uint32_t hit_cap;
  uint32_t n_spectra;
  uint32_t averaging;                    // K
  float inv_averaging;                   // 1/K
  float threshold;
  uint32_t use_window;
  uint32_t dc_ignore;
  // Row mode (second step of the four-step path for N > 2^14, scn_large.cu): the "spectra" are the
  // 2^rpb_shift rows of each intermediate buffer and the kernel stops after |X|^2 / averaging,
  // writing fp32 power [row][N] to power_out.  rpb_shift == 0 and power_out == nullptr otherwise.
  uint32_t rpb_shift;
  float* __restrict__ power_out;
  // 1 when window[n] == window[N-1-n] bit for bit (every gr-fft window is): kernels whose tables overflow L1
  // (scn_p64.cuh at N = 8192) then read taps n >= N/2 from the mirrored address, halving the table footprint.
  uint32_t win_mirror;
  // Work counter of this launch (WorkQueue below): [0] tiles handed out beyond the static ones, [1] takers retired.
  uint32_t* __restrict__ work;
  uint32_t always_zero;                  // 0; see WorkQueue::take
};

// Dynamic tile distribution for the persistent kernels.  A "taker" (a warp in scn_wpt.cuh, a CTA elsewhere) owns
// tiles taker, taker + takers statically -- so its first loads need no round trip -- and takes every later tile
// from one global counter, one tile AHEAD of the loads that need the index (the atomic's latency hides behind a
// transform).  Why: with a static stride a CTA that starts late or shares its SM with another kernel (the NCCL
// all-gather of the previous batch's records, bench.py) delays the whole launch by its entire share; with the
// counter it simply takes fewer tiles.  The last taker to retire zeroes both words, so the counter needs no reset
// between launches; a context rotates over several counters so launches on different streams never share one.
struct WorkQueue {
  uint32_t* w;
  uint32_t takers, zero;
  __device__ __forceinline__ WorkQueue(uint32_t* work, uint32_t n_takers, uint32_t always_zero)
      : w(work), takers(n_takers), zero(always_zero) {}
  // atomicAdd() on an address ptxas can prove warp-uniform is lowered to the warp-aggregated form (elect + ATOMG +
  // SHFL of the result to the active lanes) -- also from inline PTX, also as atom.inc, also behind an empty asm --
  // and that SHFL sits right behind the ATOMG: the warp then waits out the whole L2 round trip (10 % of all stall
  // samples of the warp-per-transform kernel in ncu).  `zero` is a kernel parameter that is always 0: adding
  // (zero & threadIdx.x) makes the address lane-dependent as far as ptxas can tell, so it emits a plain ATOMG whose
  // result lands in a register nobody reads until the caller needs the index.
  // (Kernel time did not move in the A/B -- the other warp of the scheduler filled the gap -- but the stall is gone.)
  // take() returns the RAW counter value; the caller adds first_dynamic() where it consumes the ticket (an add right
  // here would wait for the atomic just like the SHFL did).
  __device__ __forceinline__ uint32_t take() { return atomicAdd(w + (zero & threadIdx.x), 1u); }
  __device__ __forceinline__ uint32_t first_dynamic() const { return 2u * takers; }
  // called once per taker, after its last take() has returned
  __device__ __forceinline__ void retire() {
    if (atomicAdd(w + 1, 1u) == takers - 1u) { atomicExch(w, 0u); atomicExch(w + 1, 0u); }
  }
};

// dB = 10*log2(sqrt(p))/log2(10) = (5/log2(10)) * log2(p)
constexpr float kDbPerLog2 = 1.5051499783199060f;
// 1.5 * 2^23: float(kMagic | u) == 12582912 + u for 0 <= u < 2^22
constexpr uint32_t kMagicBits = 0x4B400000u;
constexpr float kMagic = 12582912.0f;

template <int KIND> struct KindTraits;
template <> struct KindTraits<SCN_KIND_BYTE_COMPLEX> { static constexpr int kBytes = 2; static constexpr bool kInt = true; };
template <> struct KindTraits<SCN_KIND_SHORT> { static constexpr int kBytes = 4; static constexpr bool kInt = true; };
template <> struct KindTraits<SCN_KIND_SHORT_COMPLEX> { static constexpr int kBytes = 4; static constexpr bool kInt = true; };
template <> struct KindTraits<SCN_KIND_FLOAT_COMPLEX> { static constexpr int kBytes = 8; static constexpr bool kInt = false; };

// Geometry shared by host and device.
template <int LOG2N>
struct Geometry {
  static constexpr int N = 1 << LOG2N;
  static constexpr int T = N / kPts;                              // threads per transform
  static constexpr int F = (T >= 128) ? 1 : (128 / T);            // transforms per CTA
  static constexpr int THREADS = T * F;
  static constexpr int WORDS = N / 32;                            // mask words per spectrum
  static constexpr int WARPS = THREADS / 32;
  static constexpr int WARPS_PER_FFT = (T >= 32) ? T / 32 : 1;
  static constexpr int FFTS_PER_WARP = (T >= 32) ? 1 : 32 / T;
  // Two exchange tiles (ping-pong: one barrier per exchange) whenever they leave room for
  // several CTAs per SM; the largest size falls back to one tile and two barriers.
  static constexpr int XBUFS = (LOG2N <= SCN_XBUFS_MAXLOG2) ? 2 : 1;
  static constexpr size_t kXchTile = sizeof(float2) * size_t(xch_elems(N)) * F;
  static constexpr size_t kXchBytes = kXchTile * XBUFS;
  static constexpr size_t kMaskBytes = sizeof(uint32_t) * size_t(WORDS) * F * 2;   // ping-pong by spectrum parity
  static constexpr int RED_SLOTS = (WARPS > F) ? WARPS : F;                        // (si, sq) pairs per parity
  static constexpr size_t kRedBytes = sizeof(int32_t) * 2 * RED_SLOTS * 2;         // ping-pong by tile parity
  static constexpr size_t kWorkBytes = 16;                                         // next-group index, ping-pong
  static constexpr size_t kSmemBytes = kXchBytes + kMaskBytes + kRedBytes + kWorkBytes;
  // register budget: 128/thread up to 512-thread CTAs
// (resident-CTA target per kernel variant: see min_ctas() below)
};

// ---- raw tile of one thread ----------------------------------------------------------------------
// Row r of pass 0 is M0 consecutive samples starting at sample M0*t + r*(N/R0).
template <int LOG2N, int KIND>
struct RawTile {
  static constexpr int N = 1 << LOG2N;
  static constexpr int LOG2R0 = pass_log2r(LOG2N, 0);
  static constexpr int R0 = 1 << LOG2R0, M0 = 16 / R0;
  static constexpr int kBytes = KindTraits<KIND>::kBytes;
  static constexpr bool kSplit = KIND == SCN_KIND_SHORT;
  // bytes of one contiguous run: M0 samples (or M0 int16 for each half of the split layout)
  static constexpr int RUN = kSplit ? M0 * 2 : M0 * kBytes;
  static constexpr int WPR = (RUN + 3) / 4;                       // 32-bit words per run
  static constexpr int RUNS = kSplit ? 2 : 1;
  static constexpr int WORDS = R0 * RUNS * WPR;                   // raw registers per thread
  uint32_t w[WORDS];

  template <int NB>
  static __device__ __forceinline__ void load_run(const uint8_t* __restrict__ p, uint32_t* dst) {
    if constexpr (NB == 2) {
      dst[0] = ldg_stream(reinterpret_cast<const unsigned short*>(p));
    } else if constexpr (NB == 4) {
      dst[0] = ldg_stream(reinterpret_cast<const unsigned int*>(p));
    } else if constexpr (NB == 8) {
      const uint2 v = ldg_stream(reinterpret_cast<const uint2*>(p));
      dst[0] = v.x; dst[1] = v.y;
    } else {
#pragma unroll
      for (int i = 0; i < NB / 16; i++) {
        const uint4 v = ldg_stream(reinterpret_cast<const uint4*>(p) + i);
        dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
      }
    }
  }

  __device__ __forceinline__ void load(const uint8_t* __restrict__ buf, int t) {
#pragma unroll
    for (int r = 0; r < R0; r++) {
      const int s0 = M0 * t + r * (N / R0);
      if constexpr (kSplit) {
        load_run<RUN>(buf + size_t(s0) * 2, &w[(r * 2) * WPR]);
        load_run<RUN>(buf + size_t(N) * 2 + size_t(s0) * 2, &w[(r * 2 + 1) * WPR]);
      } else {
        load_run<RUN>(buf + size_t(s0) * kBytes, &w[r * WPR]);
      }
    }
  }

  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < WORDS; i++) w[i] = 0u;
  }

  // int32 sums of I and Q over this thread's 16 samples (utility.cpp:44-48), packed dot products.
  __device__ __forceinline__ void sums(int& si, int& sq) const {
    si = 0; sq = 0;
#pragma unroll
    for (int r = 0; r < R0; r++) {
#pragma unroll
      for (int x = 0; x < WPR; x++) {
        if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
          const int v = int(w[r * WPR + x]);
          if constexpr (RUN == 2) {          // one sample: I in byte 0, Q in byte 1
            si = __dp4a(v, 0x00000001, si);
            sq = __dp4a(v, 0x00000100, sq);
          } else {
            si = __dp4a(v, 0x00010001, si);
            sq = __dp4a(v, 0x01000100, sq);
          }
        } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
          const int v = int(w[r * WPR + x]);
          si = __dp2a_lo(v, 0x00000001, si);
          sq = __dp2a_lo(v, 0x00000100, sq);
        } else if constexpr (KIND == SCN_KIND_SHORT) {
          const int a = int(w[(r * 2) * WPR + x]), b = int(w[(r * 2 + 1) * WPR + x]);
          const int sel = (RUN == 2) ? 0x00000001 : 0x00000101;   // one or two int16 per word
          si = __dp2a_lo(a, sel, si);
          sq = __dp2a_lo(b, sel, sq);
        }
      }
    }
  }

  // Converted + windowed samples into register slots q = m + r*M0.
  //   reference: float(int(x) - dc) * onebymax, then * window  (utility.cpp:52-55, process.cpp:28-34)
  //   here: onebymax is folded into w (exact, a power of two).  Fast path: the integer is placed in
  //   the mantissa of 1.5*2^23 (PRMT), one exact FADD2 removes the magic, the offset and dc, one
  //   FMUL2 applies the window -- bit-identical to int subtract -> I2F -> multiply while
  //   |dc| <= 2^21.  Slow path (the unsigned-division quirk can make dc ~ 2^32/N): integer
  //   subtract and I2F, exactly as written in the reference.
  __device__ __forceinline__ void convert(float2 (&v)[kPts], const float (&win)[kPts], int dci, int dcq) const {
    if constexpr (KIND == SCN_KIND_FLOAT_COMPLEX) {
#pragma unroll
      for (int r = 0; r < R0; r++)
#pragma unroll
        for (int m = 0; m < M0; m++) {
          const int q = m + r * M0;
          const float2 x = make_float2(__uint_as_float(w[r * WPR + 2 * m]), __uint_as_float(w[r * WPR + 2 * m + 1]));
          v[q] = __fmul2_rn(x, make_float2(win[q], win[q]));
        }
    } else {
      constexpr bool k8 = KIND == SCN_KIND_BYTE_COMPLEX;
      constexpr float kOff = k8 ? 128.0f : 32768.0f;
      const bool fast = (dci >= -(1 << 21)) && (dci <= (1 << 21)) && (dcq >= -(1 << 21)) && (dcq <= (1 << 21));
      if (fast) {
        const float2 negc = make_float2(-(kMagic + kOff + float(dci)), -(kMagic + kOff + float(dcq)));
#pragma unroll
        for (int r = 0; r < R0; r++)
#pragma unroll
          for (int m = 0; m < M0; m++) {
            const int q = m + r * M0;
            uint32_t bi, bq;
            if constexpr (k8) {
              const uint32_t x = w[r * WPR + (m >> 1)] ^ 0x80808080u;
              bi = __byte_perm(x, kMagicBits, 0x7650 + 2 * (m & 1));
              bq = __byte_perm(x, kMagicBits, 0x7651 + 2 * (m & 1));
            } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
              const uint32_t x = w[r * WPR + m] ^ 0x80008000u;
              bi = __byte_perm(x, kMagicBits, 0x7610);
              bq = __byte_perm(x, kMagicBits, 0x7632);
            } else {
              const uint32_t xa = w[(r * 2) * WPR + (m >> 1)] ^ 0x80008000u;
              const uint32_t xb = w[(r * 2 + 1) * WPR + (m >> 1)] ^ 0x80008000u;
              bi = __byte_perm(xa, kMagicBits, (m & 1) ? 0x7632 : 0x7610);
              bq = __byte_perm(xb, kMagicBits, (m & 1) ? 0x7632 : 0x7610);
            }
            const float2 d = __fadd2_rn(make_float2(__uint_as_float(bi), __uint_as_float(bq)), negc);
            v[q] = __fmul2_rn(d, make_float2(win[q], win[q]));
          }
      } else {
#pragma unroll
        for (int r = 0; r < R0; r++)
#pragma unroll
          for (int m = 0; m < M0; m++) {
            const int q = m + r * M0;
            int xi, xq;
            if constexpr (k8) {
              const uint32_t x = w[r * WPR + (m >> 1)] >> (16 * (m & 1));
              xi = int(static_cast<signed char>(x & 0xff));
              xq = int(static_cast<signed char>((x >> 8) & 0xff));
            } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
              const uint32_t x = w[r * WPR + m];
              xi = int(static_cast<short>(x & 0xffff));
              xq = int(static_cast<short>(x >> 16));
            } else {
              xi = int(static_cast<short>((w[(r * 2) * WPR + (m >> 1)] >> (16 * (m & 1))) & 0xffff));
              xq = int(static_cast<short>((w[(r * 2 + 1) * WPR + (m >> 1)] >> (16 * (m & 1))) & 0xffff));
            }
            v[q].x = __fmul_rn(float(xi - dci), win[q]);
            v[q].y = __fmul_rn(float(xq - dcq), win[q]);
          }
      }
    }
  }
};

// Resident CTAs per SM asked of ptxas.  Measured on B200 (profiles/README.md): the kernel is bound by
// FP32-pipe / shared-memory-crossbar overlap, and more resident warps beat more registers until
// spilling starts.  Register need ~ 80 (int8, K = 1) + extra raw words + 16 accumulators when K > 1.
template <int LOG2N, int KIND, bool AVG>
constexpr int min_ctas() {
#if defined(SCN_MINCTAS128)
  return (Geometry<LOG2N>::THREADS <= 128) ? SCN_MINCTAS128 : 1;
#elif defined(SCN_MINCTAS)
  return SCN_MINCTAS;
#else
  const int raw_words = RawTile<LOG2N, KIND>::WORDS;
  int regs = 80 + (raw_words > 8 ? raw_words - 8 : 0) + (AVG ? 16 : 0);
  regs = (regs + 7) / 8 * 8;
  if (regs > 128) regs = 128;
  int ctas = 65536 / (Geometry<LOG2N>::THREADS * regs);
  return ctas < 1 ? 1 : ctas;
#endif
}

// ROWS: row mode of the four-step path (scn_large.cu) -- a separate instantiation so that the
// addressing and the power-out epilogue cost the ordinary kernels nothing.
template <int LOG2N, int KIND, bool DC, bool AVG, bool ROWS = false>
__global__ void __launch_bounds__(Geometry<LOG2N>::THREADS, min_ctas<LOG2N, KIND, AVG>())
spectrum_sense_kernel(const KernelParams p) {
  using G = Geometry<LOG2N>;
  using Raw = RawTile<LOG2N, KIND>;
  constexpr int N = G::N, T = G::T, F = G::F;
  constexpr int NP = num_passes(LOG2N);
  constexpr bool kInt = KindTraits<KIND>::kInt;
  constexpr bool kDC = DC && kInt;
  constexpr size_t kBufBytes = size_t(N) * KindTraits<KIND>::kBytes;
  constexpr int R0 = Raw::R0, M0 = Raw::M0;
  static_assert(NP >= 2 && NP <= 4, "supported sizes: 2^5 .. 2^16 with 16 points per thread");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* xch_all = reinterpret_cast<float2*>(smem_raw);
  uint32_t* smask = reinterpret_cast<uint32_t*>(smem_raw + G::kXchBytes);          // [2][F][WORDS]
  int32_t* sred = reinterpret_cast<int32_t*>(smem_raw + G::kXchBytes + G::kMaskBytes);   // [2][RED_SLOTS][2]
  uint32_t* swork = reinterpret_cast<uint32_t*>(smem_raw + G::kXchBytes + G::kMaskBytes + G::kRedBytes);   // [2]

  const int tid = threadIdx.x;
  const int f = tid / T;             // which resident transform
  const int t = tid - f * T;         // thread index inside the transform
  const int lane = tid & 31;
  const int warp = tid >> 5;
  float2* xch0 = xch_all + size_t(f) * xch_elems(N);
  float2* xch1 = (G::XBUFS == 2) ? xch0 + size_t(F) * xch_elems(N) : xch0;

  // Window taps of this thread's 16 sample positions (slot q = m + r*M0) stay in registers.
  constexpr bool kWinReg = LOG2N <= SCN_WINREG_MAXLOG2;
  float win[kPts];
  auto load_window = [&]() {
#pragma unroll
    for (int r = 0; r < R0; r++)
#pragma unroll
      for (int m = 0; m < M0; m++) win[m + r * M0] = __ldg(p.window + M0 * t + m + r * (N / R0));
  };
  if constexpr (kWinReg) load_window();

  const uint32_t K = AVG ? p.averaging : 1u;
  const uint32_t n_groups = (p.n_spectra + F - 1) / F;
  const uint32_t half = N / 2;

  // Candidate bins of this thread (process.cpp:46-53), fixed for the whole launch:
  // bit q set <=> FFT bin j = t + q*T is inside the used band and outside the DC hole.
  uint32_t candbits = 0;
#pragma unroll
  for (int q = 0; q < kPts; q++) {
    const uint32_t j = t + q * T;            // FFT bin (magnitudes[j])
    const uint32_t i = j ^ half;             // shifted index: (i + N/2) % N == j
    bool cand = !(j < p.dc_ignore || (N - j) < p.dc_ignore);
    cand = cand && !(i < (half - p.use_window) || i > (half + p.use_window));
    candbits |= (cand ? 1u : 0u) << q;
  }

  // ---- tile stream of this CTA: (group, k), k = 0..K-1; groups blockIdx.x and blockIdx.x + gridDim.x are
  // static, later ones come from the launch's work counter (WorkQueue above; row mode keeps the static stride) ----
  WorkQueue wq(p.work, gridDim.x, p.always_zero);
  uint32_t g = blockIdx.x, g_after = blockIdx.x + gridDim.x, g_after2 = 0, ticket = 0;
  if (g >= n_groups) {
    if (!ROWS && tid == 0) wq.retire();
    return;
  }
  uint32_t k = 0;
  uint32_t xsel = 0;      // ping-pong selector of the exchange tile
  uint32_t tpar = 0;      // tile parity (DC partial sums ping-pong)
  uint32_t spar = 0;      // spectrum parity (mask ping-pong)

  auto tile_ptr = [&](uint32_t gg, uint32_t kk, bool& live) -> const uint8_t* {
    const uint32_t s = gg * F + f;
    live = s < p.n_spectra;
    const uint32_t ss = live ? s : 0;
    if constexpr (ROWS) {
      // buffer (ss >> rpb_shift) * K + kk, row (ss & mask) inside it
      const size_t buffer = size_t(ss >> p.rpb_shift) * K + kk;
      const size_t row = (buffer << p.rpb_shift) + (ss & ((1u << p.rpb_shift) - 1u));
      return p.raw + row * kBufBytes;
    } else {
      return p.raw + (size_t(ss) * K + kk) * kBufBytes;
    }
  };
  // DC block reduction, warp part: one (si, sq) slot per warp (T >= 32) or per transform (T < 32)
  auto reduce_dc = [&](int si, int sq, int32_t* red) {
    if constexpr (T >= 32) {
      si = __reduce_add_sync(0xffffffffu, si);      // REDUX.SUM: one instruction per sum
      sq = __reduce_add_sync(0xffffffffu, sq);
      if (lane == 0) { red[2 * warp] = si; red[2 * warp + 1] = sq; }
    } else {
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) {
        si += __shfl_xor_sync(0xffffffffu, si, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      if (t == 0) { red[2 * f] = si; red[2 * f + 1] = sq; }
    }
  };
  auto finish_dc = [&](const int32_t* red, int& odci, int& odcq) {
    int si = 0, sq = 0;
    if constexpr (T >= 32) {
      const int w0 = f * G::WARPS_PER_FFT;
#pragma unroll
      for (int i = 0; i < G::WARPS_PER_FFT; i++) { si += red[2 * (w0 + i)]; sq += red[2 * (w0 + i) + 1]; }
    } else {
      si = red[2 * f]; sq = red[2 * f + 1];
    }
    // dc = int32(uint32(sum) / N): the unsigned division of utility.cpp:49-50, N = 2^LOG2N
    odci = int(unsigned(si) >> LOG2N);
    odcq = int(unsigned(sq) >> LOG2N);
  };

  constexpr bool kTwTree = LOG2N >= 14;
// The twiddles of pass P do not depend on the exchange that feeds it, so they are produced BEFORE
// the scatter/barrier (their loads and products overlap the barrier wait) and applied after the gather.
#define SCN_TWIDDLE_PREP(P, TWL)                                                         \
    if constexpr (kTwTree) power_twiddles<LOG2N, P>(TWL, p.twiddles, t);                 \
    else product_twiddles<LOG2N, P>(TWL, p.twiddles, t);

  Raw raw;
  int dci = 0, dcq = 0;
  bool live;
  {
    const uint8_t* buf = tile_ptr(g, 0, live);
    if (live) raw.load(buf, t); else raw.zero();
    if constexpr (kDC) {
      int si, sq;
      raw.sums(si, sq);
      reduce_dc(si, sq, sred);
      __syncthreads();
      finish_dc(sred, dci, dcq);
      tpar = 1;
    }
  }

  float acc[kPts];
  while (true) {
    // ---- convert + window (current tile), then put the next tile's loads in flight ------------------
    float2 v[kPts];
    if constexpr (!kWinReg) load_window();
    raw.convert(v, win, dci, dcq);
    const bool cur_live = live;
    const uint32_t cur_g = g, cur_k = k;
    uint32_t ng = g, nk = k + 1;
    if (nk == K) { nk = 0; ng = g_after; }
    const bool has_next = ng < n_groups;
    if constexpr (!ROWS) {
      if (cur_k == 0 && tid == 0) ticket = wq.take();   // the group after g_after: requested now, parked in a register,
                                                         // handed to the CTA behind the epilogue barrier
    }
    bool next_live = false;
    if (has_next) {
      const uint8_t* nbuf = tile_ptr(ng, nk, next_live);
      if (next_live) raw.load(nbuf, t); else raw.zero();
    }

    // ---- FFT: Stockham passes with shared-memory exchanges ------------------------------------------
    // Ping-pong tiles: a tile is rewritten only two exchanges later and every thread has passed the
    // intervening barrier after its last read of it, so one barrier per exchange suffices.
    // The DC sums of the NEXT tile consume the prefetched loads, so they sit as late as possible:
    // before the epilogue barrier when this tile ends a spectrum (always, when K == 1), else before
    // the last exchange barrier.  Either way the reduction costs no barrier of its own.
    const bool epilogue_tile = (cur_k == K - 1);
    pass_butterflies<pass_log2r(LOG2N, 0)>(v);
    int ndci = 0, ndcq = 0;
#define SCN_EXCHANGE(P, LAST)                                                                   \
    {                                                                                           \
      float2* xb = (xsel & 1u) ? xch1 : xch0;                                                   \
      if constexpr (G::XBUFS == 1) __syncthreads();                                             \
      if constexpr ((P) == 0) pass0_scatter<LOG2N>(v, xb, t); else pass_scatter<LOG2N, (P)>(v, xb, t); \
      if constexpr (kDC && AVG && (LAST)) {                                                     \
        if (has_next && !epilogue_tile) {                                                       \
          int si, sq; raw.sums(si, sq); reduce_dc(si, sq, sred + tpar * (2 * G::RED_SLOTS)); }  \
      }                                                                                         \
      __syncthreads();                                                                          \
      if constexpr (kDC && AVG && (LAST)) {                                                     \
        if (has_next && !epilogue_tile) {                                                       \
          finish_dc(sred + tpar * (2 * G::RED_SLOTS), ndci, ndcq); tpar ^= 1u; }                \
      }                                                                                         \
      pass_gather<LOG2N>(v, xb, t);                                                             \
      xsel ^= 1u;                                                                               \
    }
    {
      float2 twl[15];
      SCN_TWIDDLE_PREP(1, twl)
      SCN_EXCHANGE(0, NP == 2)
      apply_twiddles(v, twl);
      dft16(v);
    }
    if constexpr (NP > 2) {
      float2 twl[15];
      SCN_TWIDDLE_PREP(2, twl)
      SCN_EXCHANGE(1, NP == 3)
      apply_twiddles(v, twl);
      dft16(v);
    }
    if constexpr (NP > 3) {
      float2 twl[15];
      SCN_TWIDDLE_PREP(3, twl)
      SCN_EXCHANGE(2, NP == 4)
      apply_twiddles(v, twl);
      dft16(v);
    }
#undef SCN_EXCHANGE

    // ---- power, K-averaging (fp32, buffer order) ---------------------------------------
    float pw[kPts];
#pragma unroll
    for (int q = 0; q < kPts; q++) {
      const float2 sq2 = __fmul2_rn(v[q], v[q]);            // fl(re*re), fl(im*im): no FMA contraction
      pw[q] = __fadd_rn(sq2.x, sq2.y);
      if constexpr (AVG) pw[q] = acc[q] = (cur_k == 0) ? pw[q] : __fadd_rn(acc[q], pw[q]);
    }

    if (ROWS && epilogue_tile) {
      // ---- row mode: averaged power out, detection happens in the finalize kernel -----------------------
      const uint32_t s = cur_g * F + f;
      if (cur_live) {
        float* out = p.power_out + size_t(s) * N;
#pragma unroll
        for (int q = 0; q < kPts; q++) out[t + q * T] = AVG ? __fmul_rn(pw[q], p.inv_averaging) : pw[q];
      }
    } else if (!ROWS && epilogue_tile) {
      // ---- dB + detection for spectrum s ----------------------------------------------------------------
      const uint32_t s = cur_g * F + f;
      uint32_t* sm = smask + spar * (F * G::WORDS);
      float db[kPts];
      bool anyraw = false;
#pragma unroll
      for (int q = 0; q < kPts; q++) {
        const float pbar = AVG ? __fmul_rn(pw[q], p.inv_averaging) : pw[q];
        db[q] = kDbPerLog2 * __log2f(pbar);
        anyraw = anyraw || (db[q] > p.threshold);          // strict >, NaN never hits (process.cpp:54)
      }
      if (p.spectra != nullptr && cur_live) {
        float* out = p.spectra + size_t(s) * N;
#pragma unroll
        for (int q = 0; q < kPts; q++) out[t + q * T] = db[q];
      }
      // Mask words of this warp (16 words: 32 lanes x 16 bins) are owned by this warp alone.
      // Common case: no lane of the warp has a bin above threshold -> just zero them.
      uint32_t hitbits = 0;
      {
        int zf, zw;                       // (transform, word) zeroed by lane < 16
        if constexpr (T >= 32) { zf = f; zw = int((uint32_t((t & ~31) + lane * T) ^ half) >> 5); }
        else { zf = warp * G::FFTS_PER_WARP + lane / G::WORDS; zw = lane % G::WORDS; }
        if (lane < 16) sm[zf * G::WORDS + zw] = 0u;
      }
      const bool warp_any = __any_sync(0xffffffffu, anyraw);
      uint32_t warp_bits = 0;
      if (warp_any) {
#pragma unroll
        for (int q = 0; q < kPts; q++) hitbits |= (db[q] > p.threshold ? 1u : 0u) << q;
        hitbits = cur_live ? (hitbits & candbits) : 0u;
        warp_bits = __reduce_or_sync(0xffffffffu, hitbits);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kPts; q++) {
          if ((warp_bits >> q) & 1u) {                      // warp-uniform
            const uint32_t b = __ballot_sync(0xffffffffu, (hitbits >> q) & 1u);
            if constexpr (T >= 32) {
              const uint32_t i0 = (uint32_t(t & ~31) + q * T) ^ half;     // shifted index of lane 0's bin
              if (lane == 0) sm[f * G::WORDS + (i0 >> 5)] = b;
            } else {
              const int fl = lane / T;
              const uint32_t bits = (b >> (fl * T)) & ((1u << T) - 1u);
              const uint32_t ib = uint32_t(q * T) ^ half;
              if (t == 0 && bits) atomicOr(&sm[f * G::WORDS + (ib >> 5)], bits << (ib & 31));
            }
          }
        }
      }
      if constexpr (kDC) {
        if (has_next) { int si, sq; raw.sums(si, sq); reduce_dc(si, sq, sred + tpar * (2 * G::RED_SLOTS)); }
      }
      if (tid == 0) swork[spar] = ticket + wq.first_dynamic();
      __syncthreads();
      g_after2 = swork[spar];
      if constexpr (kDC) {
        if (has_next) { finish_dc(sred + tpar * (2 * G::RED_SLOTS), ndci, ndcq); tpar ^= 1u; }
      }

      // per-transform: global mask words + hit count (one warp per transform, round-robin)
      for (int ff = warp; ff < F; ff += G::WARPS) {
        const uint32_t ss = cur_g * F + ff;
        if (ss >= p.n_spectra) continue;
        uint32_t total = 0;
        for (int c = 0; c < G::WORDS; c += 32) {
          const int wi = c + lane;
          const uint32_t mw = (wi < G::WORDS) ? sm[ff * G::WORDS + wi] : 0u;
          if (p.masks != nullptr && wi < G::WORDS) p.masks[size_t(ss) * G::WORDS + wi] = mw;
          total += __popc(mw);
        }
        total = __reduce_add_sync(0xffffffffu, total);
        if (lane == 0 && p.counts != nullptr) p.counts[ss] = total;
      }

      // hit records, ascending in shifted bin: rank = hits in earlier words + hits in lower bits
      if (p.hits != nullptr && warp_bits != 0u) {
#pragma unroll
        for (int q = 0; q < kPts; q++) {
          if ((warp_bits >> q) & 1u) {                      // warp-uniform
            const uint32_t i = (uint32_t(t) + q * T) ^ half;
            const uint32_t word = i >> 5;
            uint32_t before = 0, lower;
            if constexpr (T >= 32) {
              // `word` is warp-uniform and owned by this warp: its lanes count the earlier words
              // cooperatively, and the word itself is this warp's ballot (bit == lane).
              for (uint32_t x = lane; x < word; x += 32) before += __popc(sm[f * G::WORDS + x]);
              before = __reduce_add_sync(0xffffffffu, before);
              const uint32_t b = __ballot_sync(0xffffffffu, (hitbits >> q) & 1u);
              lower = __popc(b & ((1u << lane) - 1u));
            } else {
              for (uint32_t x = 0; x < word; x++) before += __popc(sm[f * G::WORDS + x]);   // < 16 words
              lower = __popc(sm[f * G::WORDS + word] & ((1u << (i & 31)) - 1u));
            }
            if ((hitbits >> q) & 1u) {
              const uint32_t rank = before + lower;
              if (rank < p.hit_cap) {
                scn_hit h;
                h.bin = i;
                h.power_db = db[q];
                p.hits[size_t(s) * p.hit_cap + rank] = h;
              }
            }
          }
        }
      }
      spar ^= 1u;
    }

    if (!has_next) break;
    if (nk == 0) g_after = ROWS ? g_after + gridDim.x : g_after2;
    g = ng; k = nk; live = next_live; dci = ndci; dcq = ndcq;
  }
  if (!ROWS && tid == 0) wq.retire();
#undef SCN_TWIDDLE_PREP
}

}  // namespace scn
