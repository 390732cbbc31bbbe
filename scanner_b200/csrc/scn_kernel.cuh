// scn_kernel.cuh -- the fused spectrum-sense kernel (sm_100a).
//
// One persistent CTA walks "groups" of F spectra (F = transforms resident in one CTA);
// for each spectrum it streams K raw IQ buffers through
//   load -> [DC sum] -> convert+scale+window -> FFT -> |X|^2 -> accumulate
// entirely in registers/shared memory, then emits dB, the detection mask, the hit
// count and the compact hit records.  Each raw sample crosses HBM exactly once.
//
// Reference arithmetic restated (file:line under the reference tree):
//   conversion      utility.cpp:34-56 (int8), :58-84 (int16 interleaved), :9-32 (int16 split)
//   DC quirk        utility.cpp:49-50  (int32 /= uint32  => unsigned division)
//   window          process.cpp:28-34  (one fp32 multiply per component)
//   FFT             fft.cpp:20-25      (forward, unnormalised)
//   dB              utility.cpp:91-97  (10*log2(sqrt(re^2+im^2))/log2(10))
//   detection       process.cpp:46-61  (fftshift index, DC hole, used band, strict >)
//   time domain     process.cpp:203-237
#pragma once
#include "scn_fft.cuh"
#include "../../include/scanner_b200.h"

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
};

// dB = 10*log2(sqrt(p))/log2(10) = (5/log2(10)) * log2(p)
constexpr float kDbPerLog2 = 1.5051499783199060f;

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
  // Two exchange tiles (ping-pong: one barrier per exchange) whenever they leave room for
  // several CTAs per SM; the largest size falls back to one tile and two barriers.
  static constexpr int XBUFS = (LOG2N <= 13) ? 2 : 1;
  static constexpr size_t kXchTile = sizeof(float2) * size_t(xch_elems(N)) * F;
  static constexpr size_t kXchBytes = kXchTile * XBUFS;
  // register budget: 128/thread up to 512-thread CTAs
  static constexpr int MIN_CTAS = (THREADS <= 128) ? 4 : (THREADS <= 256 ? 2 : 1);
  static constexpr size_t kMaskBytes = sizeof(uint32_t) * size_t(WORDS) * F * 2;   // mask + prefix
  static constexpr size_t kRedBytes = sizeof(int32_t) * 2 * WARPS;
  static constexpr size_t kSmemBytes = kXchBytes + kMaskBytes + kRedBytes;
};

// Raw sample fetch: integer I,Q of sample n of the buffer at `buf`.
template <int KIND, int N>
__device__ __forceinline__ void load_raw_int(const uint8_t* __restrict__ buf, int n, int& i, int& q) {
  if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
    const unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(buf) + n);
    i = static_cast<int>(static_cast<signed char>(u & 0xff));
    q = static_cast<int>(static_cast<signed char>(u >> 8));
  } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
    const unsigned int u = __ldg(reinterpret_cast<const unsigned int*>(buf) + n);
    i = static_cast<int>(static_cast<short>(u & 0xffff));
    q = static_cast<int>(static_cast<short>(u >> 16));
  } else {   // SCN_KIND_SHORT: re[N] then im[N]
    const short* p = reinterpret_cast<const short*>(buf);
    i = static_cast<int>(__ldg(p + n));
    q = static_cast<int>(__ldg(p + N + n));
  }
}

template <int LOG2N, int KIND, bool DC>
__global__ void __launch_bounds__(Geometry<LOG2N>::THREADS, Geometry<LOG2N>::MIN_CTAS)
spectrum_sense_kernel(const KernelParams p) {
  using G = Geometry<LOG2N>;
  constexpr int N = G::N, T = G::T, F = G::F;
  constexpr int NP = num_passes(LOG2N);
  constexpr bool kInt = KindTraits<KIND>::kInt;
  constexpr size_t kBufBytes = size_t(N) * KindTraits<KIND>::kBytes;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* xch_all = reinterpret_cast<float2*>(smem_raw);
  uint32_t* smask = reinterpret_cast<uint32_t*>(smem_raw + G::kXchBytes);     // [F][WORDS]
  uint32_t* sprefix = smask + F * G::WORDS;                                     // [F][WORDS]
  int32_t* sred = reinterpret_cast<int32_t*>(smem_raw + G::kXchBytes + G::kMaskBytes);

  const int tid = threadIdx.x;
  const int f = tid / T;             // which resident transform
  const int t = tid - f * T;         // thread index inside the transform
  const int lane = tid & 31;
  const int warp = tid >> 5;
  float2* xch0 = xch_all + size_t(f) * xch_elems(N);
  float2* xch1 = (G::XBUFS == 2) ? xch0 + size_t(F) * xch_elems(N) : xch0;

  // Window taps for this thread's 16 sample positions stay in registers for the whole launch.
  float w[kPts];
#pragma unroll
  for (int q = 0; q < kPts; q++) w[q] = __ldg(p.window + t + q * T);

  const uint32_t K = p.averaging;
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
  uint32_t xsel = 0;   // ping-pong selector of the exchange tile

  for (uint32_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const uint32_t s = g * F + f;                 // this transform's spectrum
    const bool live = s < p.n_spectra;
    float acc[kPts];

    for (uint32_t k = 0; k < K; k++) {
      const uint8_t* buf = p.raw + (size_t(live ? s : 0) * K + k) * kBufBytes;
      float2 v[kPts];

      // ---- load + convert + window -------------------------------------------------
      if constexpr (kInt) {
        int xi[kPts], xq[kPts];
#pragma unroll
        for (int q = 0; q < kPts; q++) load_raw_int<KIND, N>(buf, t + q * T, xi[q], xq[q]);
        int dci = 0, dcq = 0;
        if constexpr (DC) {
          // int32 sums over the whole buffer (utility.cpp:44-48), then the unsigned
          // division of utility.cpp:49-50: dc = int32(uint32(sum) / N), N = 2^LOG2N.
          int si = 0, sq = 0;
#pragma unroll
          for (int q = 0; q < kPts; q++) { si += xi[q]; sq += xq[q]; }
          constexpr int SEG = (T < 32) ? T : 32;
#pragma unroll
          for (int o = SEG / 2; o > 0; o >>= 1) {
            si += __shfl_xor_sync(0xffffffffu, si, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
          }
          if constexpr (T > 32) {
            if (lane == 0) { sred[2 * warp] = si; sred[2 * warp + 1] = sq; }
            __syncthreads();
            si = 0; sq = 0;
            const int w0 = f * G::WARPS_PER_FFT;
#pragma unroll
            for (int i = 0; i < G::WARPS_PER_FFT; i++) { si += sred[2 * (w0 + i)]; sq += sred[2 * (w0 + i) + 1]; }
          }
          dci = static_cast<int>(static_cast<unsigned>(si) >> LOG2N);
          dcq = static_cast<int>(static_cast<unsigned>(sq) >> LOG2N);
        }
        // float(int(x) - dc) * onebymax * window: onebymax is a signed power of two, so
        // folding it into the window table is exact.
#pragma unroll
        for (int q = 0; q < kPts; q++) {
          v[q].x = __fmul_rn(static_cast<float>(xi[q] - dci), w[q]);
          v[q].y = __fmul_rn(static_cast<float>(xq[q] - dcq), w[q]);
        }
      } else {
        const float2* fb = reinterpret_cast<const float2*>(buf);
#pragma unroll
        for (int q = 0; q < kPts; q++) v[q] = __ldg(fb + t + q * T);
#pragma unroll
        for (int q = 0; q < kPts; q++) {
          v[q].x = __fmul_rn(v[q].x, w[q]);
          v[q].y = __fmul_rn(v[q].y, w[q]);
        }
      }

      // ---- FFT: Stockham passes with shared-memory exchanges ---------------------------
      // Ping-pong tiles: a tile is rewritten only two exchanges later, and every thread has
      // passed the intervening barrier after its last read of it, so one barrier per
      // exchange suffices (two when there is a single tile).
      pass_butterflies<pass_log2r(LOG2N, 0)>(v);
#define SCN_EXCHANGE(P)                                                      \
      {                                                                      \
        float2* xb = (xsel & 1u) ? xch1 : xch0;                              \
        if constexpr (G::XBUFS == 1) __syncthreads();                        \
        pass_scatter<LOG2N, P>(v, xb, t);                                    \
        __syncthreads();                                                     \
        pass_gather<LOG2N>(v, xb, t);                                        \
        xsel ^= 1u;                                                          \
      }
      if constexpr (NP > 1) {
        SCN_EXCHANGE(0)
        pass_twiddle<LOG2N, 1>(v, p.twiddles, t);
        pass_butterflies<pass_log2r(LOG2N, 1)>(v);
      }
      if constexpr (NP > 2) {
        SCN_EXCHANGE(1)
        pass_twiddle<LOG2N, 2>(v, p.twiddles, t);
        pass_butterflies<pass_log2r(LOG2N, 2)>(v);
      }
      if constexpr (NP > 3) {
        SCN_EXCHANGE(2)
        pass_twiddle<LOG2N, 3>(v, p.twiddles, t);
        pass_butterflies<pass_log2r(LOG2N, 3)>(v);
      }
#undef SCN_EXCHANGE

      // ---- power, K-averaging (fp32, buffer order) -------------------------
#pragma unroll
      for (int q = 0; q < kPts; q++) {
        const float pw = __fadd_rn(__fmul_rn(v[q].x, v[q].x), __fmul_rn(v[q].y, v[q].y));
        acc[q] = (k == 0) ? pw : __fadd_rn(acc[q], pw);
      }
    }

    // ---- dB + detection ------------------------------------------------------------------
    float db[kPts];
    uint32_t hitbits = 0;
#pragma unroll
    for (int q = 0; q < kPts; q++) {
      const float pbar = (K == 1) ? acc[q] : __fmul_rn(acc[q], p.inv_averaging);
      db[q] = kDbPerLog2 * __log2f(pbar);
      hitbits |= (db[q] > p.threshold ? 1u : 0u) << q;     // strict >, NaN never hits (process.cpp:54)
    }
    hitbits = live ? (hitbits & candbits) : 0u;
    if (p.spectra != nullptr && live) {
      float* out = p.spectra + size_t(s) * N;
#pragma unroll
      for (int q = 0; q < kPts; q++) out[t + q * T] = db[q];
    }

    // mask words into shared memory (bit i of word i>>5)
    if constexpr (T < 32) {
      for (int x = tid; x < F * G::WORDS; x += G::THREADS) smask[x] = 0;
      __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < kPts; q++) {
      const uint32_t b = __ballot_sync(0xffffffffu, (hitbits >> q) & 1u);
      const uint32_t i0 = (uint32_t(t & ~31) + q * T) ^ half;       // shifted index of lane 0's bin (T >= 32)
      if constexpr (T >= 32) {
        if (lane == 0) smask[f * G::WORDS + (i0 >> 5)] = b;
      } else {
        // two or more transforms share a warp: lanes [fl*T, fl*T+T) belong to transform f
        const int fl = lane / T;
        const uint32_t bits = (b >> (fl * T)) & ((1u << T) - 1u);
        const uint32_t ib = (uint32_t(q * T)) ^ half;
        if ((lane % T) == 0 && bits) atomicOr(&smask[f * G::WORDS + (ib >> 5)], bits << (ib & 31));
      }
    }
    __syncthreads();

    // per-transform scan of the mask: global mask words, exclusive prefix, hit count
    for (int ff = warp; ff < F; ff += G::WARPS) {
      const uint32_t ss = g * F + ff;
      if (ss >= p.n_spectra) continue;
      uint32_t base = 0;
      for (int c = 0; c < G::WORDS; c += 32) {
        const int wi = c + lane;
        const uint32_t mw = (wi < G::WORDS) ? smask[ff * G::WORDS + wi] : 0u;
        if (p.masks != nullptr && wi < G::WORDS) p.masks[size_t(ss) * G::WORDS + wi] = mw;
        const uint32_t pc = __popc(mw);
        uint32_t incl = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += y;
        }
        if (wi < G::WORDS) sprefix[ff * G::WORDS + wi] = base + incl - pc;
        base += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0 && p.counts != nullptr) p.counts[ss] = base;
    }

    if (p.hits != nullptr) {
      __syncthreads();
      if (hitbits) {
#pragma unroll
        for (int q = 0; q < kPts; q++) {
          if ((hitbits >> q) & 1u) {
            const uint32_t i = (uint32_t(t) + q * T) ^ half;
            const uint32_t mw = smask[f * G::WORDS + (i >> 5)];
            const uint32_t rank = sprefix[f * G::WORDS + (i >> 5)] + __popc(mw & ((1u << (i & 31)) - 1u));
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
    __syncthreads();   // smask / sprefix / xch are reused by the next group
  }
}

}  // namespace scn
