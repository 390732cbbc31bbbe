// scanner_oracle.cpp -- CPU restatement of wpats/scanner's spectrum-sense hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under scanner_b200/ may link, import or call
// this file; it is used by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.
//
// Parity pinning: the reference ships no tests or golden vectors.
// This restatement is pinned against the reference's OWN sources compiled here
// (oracle/_ref, see oracle/Makefile): utility.cpp, frequencyTable.cpp, fft.cpp,
// process.cpp and messageQueue.h are compiled where they lie under /root/reference
// with shim headers standing in for the absent third-party libraries (FFTW3, VOLK,
// gr-fft, Boost).  The FFT itself (FFTW, absent) is the published forward,
// unnormalised DFT X[k] = sum_n x[n] exp(-2 pi i nk/N); FFTW_MEASURE makes the
// reference's own rounding machine dependent, so FFT parity is tolerance based
// (1e-3 dB) against the double-precision transform below.
//
// Every function cites the reference file:line it follows.
//
// Build: see oracle/Makefile (g++ -O3 -march=native -ffp-contract=off).
//   -ffp-contract=off matters: the reference is built for baseline x86-64
//   (Makefile:23, no -march), so re*re + im*im is two roundings, never an FMA.

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>
#include <mutex>
#include <map>
#include <memory>
#include <limits>
#include <algorithm>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

// ---------------------------------------------------------------------------
// Sample kinds -- numbering of SampleQueue::SampleKind, messageQueue.h:31-37.
// ---------------------------------------------------------------------------
enum SampleKind : uint32_t {
  kIllegal = 0,
  kByteComplex = 1,   // int8_t  [N][2]        (HackRF, RTL)
  kShort = 2,         // int16_t re[N], im[N]  (SDRplay)  -- raw layout here: re block then im block
  kShortComplex = 3,  // int16_t [N][2]        (BladeRF)
  kFloatComplex = 4   // float   [N][2]        (B210, Airspy)
};

inline uint32_t bytes_per_sample(uint32_t kind) {
  switch (kind) {
    case kByteComplex: return 2;
    case kShort: return 4;
    case kShortComplex: return 4;
    case kFloatComplex: return 8;
    default: return 0;
  }
}

// ---------------------------------------------------------------------------
// Conversion -- utility.cpp:9-84.
//   max      = intK_t(1 << (enob-1))          (wraps: enob 8 -> int8 -128, utility.cpp:40)
//   onebymax = float(1.0 / max)
//   dc       = int32 sum /= uint32 count      (unsigned division quirk, utility.cpp:25-26,49-50,77-78)
//   dst      = float(int(src) - dc) * onebymax
// ---------------------------------------------------------------------------
inline int32_t dc_divide(int32_t sum, uint32_t count) {
  // `dc_real /= sampleCount` with int32_t /= uint32_t: usual arithmetic conversions make
  // both operands unsigned; the quotient converts back to int32_t (modular).
  return static_cast<int32_t>(static_cast<uint32_t>(sum) / count);
}

void convert_byte_complex(const int8_t* src, float* dst, uint32_t n, uint32_t enob, bool dc) {
  int8_t max = static_cast<int8_t>(1 << (enob - 1));           // utility.cpp:40
  float onebymax = float(1.0 / max);                            // utility.cpp:41
  int32_t dc_re = 0, dc_im = 0;
  if (dc) {                                                     // utility.cpp:44-51
    for (uint32_t i = 0; i < n; i++) { dc_re += src[2 * i]; dc_im += src[2 * i + 1]; }
    dc_re = dc_divide(dc_re, n);
    dc_im = dc_divide(dc_im, n);
  }
  for (uint32_t i = 0; i < n; i++) {                            // utility.cpp:52-55
    dst[2 * i] = float(src[2 * i] - dc_re) * onebymax;
    dst[2 * i + 1] = float(src[2 * i + 1] - dc_im) * onebymax;
  }
}

void convert_short_complex(const int16_t* src, float* dst, uint32_t n, uint32_t enob, bool dc) {
  int16_t max = static_cast<int16_t>(1 << (enob - 1));         // utility.cpp:64
  float onebymax = float(1.0 / max);                            // utility.cpp:65
  int32_t dc_re = 0, dc_im = 0;
  if (dc) {                                                     // utility.cpp:70-79
    for (uint32_t i = 0; i < n; i++) { dc_re += src[2 * i]; dc_im += src[2 * i + 1]; }
    dc_re = dc_divide(dc_re, n);
    dc_im = dc_divide(dc_im, n);
  }
  for (uint32_t i = 0; i < n; i++) {                            // utility.cpp:80-83
    dst[2 * i] = float(src[2 * i] - dc_re) * onebymax;
    dst[2 * i + 1] = float(src[2 * i + 1] - dc_im) * onebymax;
  }
}

void convert_short_split(const int16_t* re, const int16_t* im, float* dst, uint32_t n,
                         uint32_t enob, bool dc) {
  int16_t max = static_cast<int16_t>(1 << (enob - 1));         // utility.cpp:16
  float onebymax = float(1.0 / max);                            // utility.cpp:17
  int32_t dc_re = 0, dc_im = 0;
  if (dc) {                                                     // utility.cpp:20-27
    for (uint32_t i = 0; i < n; i++) { dc_re += re[i]; dc_im += im[i]; }
    dc_re = dc_divide(dc_re, n);
    dc_im = dc_divide(dc_im, n);
  }
  for (uint32_t i = 0; i < n; i++) {                            // utility.cpp:28-31
    dst[2 * i] = float(re[i] - dc_re) * onebymax;
    dst[2 * i + 1] = float(im[i] - dc_im) * onebymax;
  }
}

// Kind dispatch of MessageQueue::AppendSamples, messageQueue.h:190-237.
void convert_any(uint32_t kind, const void* raw, float* dst, uint32_t n, uint32_t enob, bool dc) {
  switch (kind) {
    case kByteComplex: convert_byte_complex(static_cast<const int8_t*>(raw), dst, n, enob, dc); break;
    case kShort: {
      const int16_t* p = static_cast<const int16_t*>(raw);
      convert_short_split(p, p + n, dst, n, enob, dc);
      break;
    }
    case kShortComplex: convert_short_complex(static_cast<const int16_t*>(raw), dst, n, enob, dc); break;
    case kFloatComplex: std::memcpy(dst, raw, sizeof(float) * 2 * n); break;   // messageQueue.h:231-237
    default: break;
  }
}

// ---------------------------------------------------------------------------
// Window tables -- process.cpp:14-21 calls gr::fft::window::build(type, N, 0.0).
// gr-fft is not under /root/reference (linked by Makefile:10, version unpinned);
// this restates its published definitions: symmetric windows, M = N-1, coefficients
// evaluated in double and stored as float.  Enum numbering follows gr::fft::window::win_type.
// The table crosses the C ABI as data, so oracle and GPU always see identical taps.
// ---------------------------------------------------------------------------
enum WinType : int {
  WIN_HAMMING = 0, WIN_HANN = 1, WIN_BLACKMAN = 2, WIN_RECTANGULAR = 3,
  WIN_KAISER = 4, WIN_BLACKMAN_hARRIS = 5
};

void window_build(int type, uint32_t n, float* out) {
  const double M = double(n) - 1.0;
  for (uint32_t i = 0; i < n; i++) {
    double x = (n > 1) ? double(i) / M : 0.0;
    double w = 1.0;
    switch (type) {
      case WIN_HAMMING: w = 0.54 - 0.46 * std::cos(2 * kPi * x); break;
      case WIN_HANN: w = 0.5 - 0.5 * std::cos(2 * kPi * x); break;
      case WIN_BLACKMAN: w = 0.42 - 0.5 * std::cos(2 * kPi * x) + 0.08 * std::cos(4 * kPi * x); break;
      case WIN_RECTANGULAR: w = 1.0; break;
      case WIN_BLACKMAN_hARRIS:   // 92 dB, 4-term (gr-fft default attenuation)
        w = 0.35875 - 0.48829 * std::cos(2 * kPi * x) + 0.14128 * std::cos(4 * kPi * x) -
            0.01168 * std::cos(6 * kPi * x);
        break;
      default: w = 1.0; break;
    }
    out[i] = float(w);
  }
}

// FFTWindow::apply, process.cpp:28-34: volk_32fc_32f_multiply_32fc_a == one IEEE fp32
// multiply per component.
void window_apply(float* iq, const float* w, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) {
    iq[2 * i] = iq[2 * i] * w[i];
    iq[2 * i + 1] = iq[2 * i + 1] * w[i];
  }
}

// ---------------------------------------------------------------------------
// FFT -- fft.cpp:4-25: forward (sign -1), unnormalised, out of place, size N.
// Templated Stockham radix-4 (+ one radix-2 pass when log2 N is odd) with a
// precomputed double-derived twiddle table.  T = float models the reference's fp32
// transform and is the timed CPU baseline; T = double is the accuracy arbiter.
// ---------------------------------------------------------------------------
template <typename T>
struct FftPlan {
  uint32_t n = 0;
  std::vector<T> tw;        // tw[2k], tw[2k+1] = cos, -sin of 2 pi k / n
  std::vector<T> scratch;   // ping-pong buffer, 2n
  explicit FftPlan(uint32_t size) : n(size), tw(2 * size), scratch(2 * size) {
    for (uint32_t k = 0; k < size; k++) {
      double a = -2.0 * kPi * double(k) / double(size);
      tw[2 * k] = T(std::cos(a));
      tw[2 * k + 1] = T(std::sin(a));
    }
  }
  // out may alias neither in nor scratch.
  void execute(const T* in, T* out) {
    const uint32_t N = n;
    if (N == 1) { out[0] = in[0]; out[1] = in[1]; return; }
    uint32_t log2n = 0;
    while ((1u << log2n) < N) log2n++;
    uint32_t passes = log2n / 2 + (log2n & 1);
    // choose ping-pong start so the final pass lands in `out`
    T* bufs[2] = {scratch.data(), out};
    const T* src = in;
    uint32_t which = (passes & 1) ? 1 : 0;   // destination of first pass
    uint32_t Ns = 1;
    uint32_t remaining = log2n;
    while (remaining > 0) {
      T* dst = bufs[which];
      if (remaining >= 2) {
        const uint32_t R = 4, Tn = N / R;
        for (uint32_t j = 0; j < Tn; j++) {
          uint32_t k = j & (Ns - 1);
          uint32_t tstep = (N / (Ns * R)) * k;   // W_{Ns*R}^{k r} = W_N^{tstep r}
          T v[8];
          for (uint32_t r = 0; r < 4; r++) {
            T xr = src[2 * (j + r * Tn)], xi = src[2 * (j + r * Tn) + 1];
            T c = tw[2 * ((tstep * r) & (N - 1))], s = tw[2 * ((tstep * r) & (N - 1)) + 1];
            v[2 * r] = xr * c - xi * s;
            v[2 * r + 1] = xr * s + xi * c;
          }
          T a0r = v[0] + v[4], a0i = v[1] + v[5];
          T a1r = v[0] - v[4], a1i = v[1] - v[5];
          T a2r = v[2] + v[6], a2i = v[3] + v[7];
          T a3r = v[3] - v[7], a3i = v[6] - v[2];   // (v1 - v3) * (-i)
          uint32_t j0 = ((j - k) * R) + k;
          dst[2 * (j0)] = a0r + a2r;            dst[2 * (j0) + 1] = a0i + a2i;
          dst[2 * (j0 + Ns)] = a1r + a3r;       dst[2 * (j0 + Ns) + 1] = a1i + a3i;
          dst[2 * (j0 + 2 * Ns)] = a0r - a2r;   dst[2 * (j0 + 2 * Ns) + 1] = a0i - a2i;
          dst[2 * (j0 + 3 * Ns)] = a1r - a3r;   dst[2 * (j0 + 3 * Ns) + 1] = a1i - a3i;
        }
        Ns *= 4;
        remaining -= 2;
      } else {
        const uint32_t R = 2, Tn = N / R;
        for (uint32_t j = 0; j < Tn; j++) {
          uint32_t k = j & (Ns - 1);
          uint32_t tstep = (N / (Ns * R)) * k;
          T x0r = src[2 * j], x0i = src[2 * j + 1];
          T xr = src[2 * (j + Tn)], xi = src[2 * (j + Tn) + 1];
          T c = tw[2 * tstep], s = tw[2 * tstep + 1];
          T x1r = xr * c - xi * s, x1i = xr * s + xi * c;
          uint32_t j0 = ((j - k) * R) + k;
          dst[2 * j0] = x0r + x1r;          dst[2 * j0 + 1] = x0i + x1i;
          dst[2 * (j0 + Ns)] = x0r - x1r;   dst[2 * (j0 + Ns) + 1] = x0i - x1i;
        }
        Ns *= 2;
        remaining -= 1;
      }
      src = dst;
      which ^= 1;
    }
  }
};

// Eight transforms at once, one per SIMD lane (GCC vector extension; AVX on the hosts this runs on): the same
// Stockham passes, the same operation order per lane, so every lane is BIT-IDENTICAL to FftPlan<float>::execute
// (tests/test_oracle_golden.py checks that).  This is what lets the timed CPU baseline run its FFT at the speed of
// a tuned SIMD library (FFTW vectorises inside one transform; batching across buffers is the simple way to the same
// lane utilisation) without changing a single rounding.
typedef float vf8 __attribute__((vector_size(32), aligned(32)));
struct FftPlanBatch8 {
  uint32_t n = 0;
  std::vector<float> tw;
  std::vector<vf8> a, b;                       // ping-pong, 2n vectors each (re, im interleaved per sample)
  explicit FftPlanBatch8(uint32_t size) : n(size), tw(2 * size), a(2 * size), b(2 * size) {
    for (uint32_t k = 0; k < size; k++) {
      double ang = -2.0 * kPi * double(k) / double(size);
      tw[2 * k] = float(std::cos(ang));
      tw[2 * k + 1] = float(std::sin(ang));
    }
  }
  // in[l], out[l]: 8 interleaved complex buffers of n samples
  void execute(const float* const in[8], float* const out[8]) {
    const uint32_t N = n;
    for (uint32_t i = 0; i < 2 * N; i++) {
      vf8 v;
      for (int l = 0; l < 8; l++) v[l] = in[l][i];
      a[i] = v;
    }
    uint32_t log2n = 0;
    while ((1u << log2n) < N) log2n++;
    const vf8* src = a.data();
    vf8* bufs[2] = {b.data(), a.data()};
    uint32_t which = 0, Ns = 1, remaining = log2n;
    while (remaining > 0) {
      vf8* dst = bufs[which];
      if (remaining >= 2) {
        const uint32_t R = 4, Tn = N / R;
        for (uint32_t j = 0; j < Tn; j++) {
          uint32_t k = j & (Ns - 1);
          uint32_t tstep = (N / (Ns * R)) * k;
          vf8 v[8];
          for (uint32_t r = 0; r < 4; r++) {
            vf8 xr = src[2 * (j + r * Tn)], xi = src[2 * (j + r * Tn) + 1];
            float c = tw[2 * ((tstep * r) & (N - 1))], sn = tw[2 * ((tstep * r) & (N - 1)) + 1];
            v[2 * r] = xr * c - xi * sn;
            v[2 * r + 1] = xr * sn + xi * c;
          }
          vf8 a0r = v[0] + v[4], a0i = v[1] + v[5];
          vf8 a1r = v[0] - v[4], a1i = v[1] - v[5];
          vf8 a2r = v[2] + v[6], a2i = v[3] + v[7];
          vf8 a3r = v[3] - v[7], a3i = v[6] - v[2];
          uint32_t j0 = ((j - k) * R) + k;
          dst[2 * (j0)] = a0r + a2r;            dst[2 * (j0) + 1] = a0i + a2i;
          dst[2 * (j0 + Ns)] = a1r + a3r;       dst[2 * (j0 + Ns) + 1] = a1i + a3i;
          dst[2 * (j0 + 2 * Ns)] = a0r - a2r;   dst[2 * (j0 + 2 * Ns) + 1] = a0i - a2i;
          dst[2 * (j0 + 3 * Ns)] = a1r - a3r;   dst[2 * (j0 + 3 * Ns) + 1] = a1i - a3i;
        }
        Ns *= 4;
        remaining -= 2;
      } else {
        const uint32_t R = 2, Tn = N / R;
        for (uint32_t j = 0; j < Tn; j++) {
          uint32_t k = j & (Ns - 1);
          uint32_t tstep = (N / (Ns * R)) * k;
          vf8 x0r = src[2 * j], x0i = src[2 * j + 1];
          vf8 xr = src[2 * (j + Tn)], xi = src[2 * (j + Tn) + 1];
          float c = tw[2 * tstep], sn = tw[2 * tstep + 1];
          vf8 x1r = xr * c - xi * sn, x1i = xr * sn + xi * c;
          uint32_t j0 = ((j - k) * R) + k;
          dst[2 * j0] = x0r + x1r;          dst[2 * j0 + 1] = x0i + x1i;
          dst[2 * (j0 + Ns)] = x0r - x1r;   dst[2 * (j0 + Ns) + 1] = x0i - x1i;
        }
        Ns *= 2;
        remaining -= 1;
      }
      src = dst;
      which ^= 1;
    }
    for (uint32_t i = 0; i < 2 * N; i++) {
      const vf8 v = src[i];
      for (int l = 0; l < 8; l++) out[l][i] = v[l];
    }
  }
};

// ---------------------------------------------------------------------------
// dB -- Utility::complex_to_magnitude, utility.cpp:86-98:
//   double log10 = log2(10);
//   float mag = sqrt(re*re + im*im);
//   magnitudes[i] = 10 * log2(mag) / log10;
// variant 0 ("cmath"): log2/sqrt bind to the float overloads (GCC >= 6 <math.h>, and what
//   oracle/_ref builds here): L = log2f(mag); 10*L in fp32; divide in double; store fp32.
// variant 1 ("c-math"): log2/sqrt bind to double (pre-GCC-6 <math.h>):
//   mag = float(sqrt(double(p))); L = log2(double(mag)); 10*L/log10 in double; store fp32.
// SURVEY.md section 0 defect 3: the two differ by <= 2 ulp.
// ---------------------------------------------------------------------------
inline float power_of(float re, float im) { return re * re + im * im; }   // two roundings + add

inline float db_from_power(float p, int variant) {
  const double log10v = std::log2(10.0);
  if (variant == 1) {
    float mag = float(std::sqrt(double(p)));
    return float(10 * std::log2(double(mag)) / log10v);
  }
  float mag = std::sqrt(p);
  return float(10 * std::log2(mag) / log10v);
}

inline double db_from_power_f64(double p) {
  return 10.0 * std::log2(std::sqrt(p)) / std::log2(10.0);
}

// ---------------------------------------------------------------------------
// Detection -- ProcessSamples::process_fft, process.cpp:36-64; derived parameters
// process.cpp:85-88 (useWindow = uint32(useBandWidth * N / 2.0), dcIgnoreWindow = 4).
// Emits hits in ascending shifted index i; mask bit i (word i>>5, bit i&31).
// ---------------------------------------------------------------------------
template <typename DB>
uint32_t detect(const DB* db, uint32_t n, uint32_t use_window, uint32_t dc_ignore, float threshold,
                uint32_t* mask /* n/32 words, zeroed here */, uint32_t* hit_bins, uint32_t hit_cap) {
  uint32_t words = (n + 31) / 32;
  if (mask) std::memset(mask, 0, sizeof(uint32_t) * words);
  uint32_t half = n / 2, count = 0;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t j = (i + half) % n;                                          // process.cpp:47
    if (j < dc_ignore || (n - j) < dc_ignore) continue;                   // process.cpp:48-50
    if (i < (half - use_window) || i > (half + use_window)) continue;     // process.cpp:51-53 (uint32 arithmetic)
    if (db[j] > threshold) {                                              // process.cpp:54
      if (mask) mask[i >> 5] |= 1u << (i & 31);
      if (hit_bins && count < hit_cap) hit_bins[count] = i;
      count++;
    }
  }
  return count;
}

// Hz of shifted bin i -- process.cpp:38-39,55,57.
inline uint64_t hit_frequency(double center, uint32_t sample_rate, uint32_t n, uint32_t i) {
  double start_frequency = center - sample_rate / 2;   // uint32 / int -> uint32, then to double
  uint32_t bin_step = sample_rate / n;
  double frequency = start_frequency + i * bin_step;   // uint32 * uint32 (wraps), then to double
  return uint64_t(frequency);
}

// ---------------------------------------------------------------------------
// Frequency table -- FrequencyTable::FrequencyTable, frequencyTable.cpp:9-37.
// ---------------------------------------------------------------------------
uint32_t frequency_table(uint32_t sample_rate, double start, double stop, double use_bw,
                         double dc_ignore, double* out, uint32_t cap) {
  double f1 = start + use_bw / 2 * sample_rate;
  double step = use_bw;
  if (dc_ignore > 0) step = (use_bw - dc_ignore) / 2;
  uint32_t count = 0;
  if (stop == 0.0) {
    count = 1;
  } else {
    for (; (f1 + count * step * double(sample_rate)) < stop; count++) {}
  }
  for (uint32_t i = 0; i < count && i < cap; i++) out[i] = f1 + i * step * double(sample_rate);
  return count;
}

// ---------------------------------------------------------------------------
// Whole pipeline for a batch of spectra.
// Order of operations: process.cpp:131-144 (Run) / 272-310 (ThreadWorker):
//   convert -> window -> FFT -> dB -> detect, one buffer at a time.
// Extension (absent from the reference): K-buffer linear power
// averaging, acc = sum_t p_t accumulated in fp32 in buffer order, Pbar = acc * (1/K)
// (K = 1 => Pbar == p bit for bit), then dB from Pbar.
// precision 0: fp32 FFT, fp32 power  (reference-like);
// precision 1: fp64 FFT of the SAME fp32 windowed samples, fp64 power and dB (ground truth).
// ---------------------------------------------------------------------------
struct PipelineConfig {
  uint32_t n, sample_rate, enob, kind, correct_dc, averaging;
  float threshold;
  uint32_t use_window, dc_ignore_window;
  int db_variant;
};

template <typename T>
struct PipelineWorker {
  PipelineConfig c;
  const float* window;
  FftPlan<T> plan;
  std::vector<float> iq, dbf;
  std::vector<T> fin, fout, acc;
  std::vector<double> dbd;
  PipelineWorker(const PipelineConfig& cfg, const float* w)
      : c(cfg), window(w), plan(cfg.n), iq(2 * size_t(cfg.n)), dbf(cfg.n), fin(2 * size_t(cfg.n)),
        fout(2 * size_t(cfg.n)), acc(cfg.n), dbd(cfg.n) {}

  // Processes spectrum s (buffers s*K .. s*K+K-1 of raw); returns its hit count.
  uint32_t run(const uint8_t* raw, uint32_t s, float* spectra_db, double* spectra_db64,
               uint32_t* masks) {
    const uint32_t N = c.n, K = c.averaging ? c.averaging : 1;
    const size_t buf_bytes = size_t(N) * bytes_per_sample(c.kind);
    const uint32_t words = (N + 31) / 32;
    for (uint32_t k = 0; k < K; k++) {
      const uint8_t* buf = raw + (size_t(s) * K + k) * buf_bytes;
      convert_any(c.kind, buf, iq.data(), N, c.enob, c.correct_dc != 0);
      if (window) window_apply(iq.data(), window, N);
      for (size_t i = 0; i < 2 * size_t(N); i++) fin[i] = T(iq[i]);
      plan.execute(fin.data(), fout.data());
      for (uint32_t i = 0; i < N; i++) {
        T re = fout[2 * i], im = fout[2 * i + 1];
        T p = re * re + im * im;
        acc[i] = (k == 0) ? p : acc[i] + p;
      }
    }
    if (sizeof(T) == sizeof(float)) {
      const float invk = 1.0f / float(K);
      for (uint32_t i = 0; i < N; i++) {
        float p = (K == 1) ? float(acc[i]) : float(acc[i]) * invk;
        dbf[i] = db_from_power(p, c.db_variant);
      }
      if (spectra_db) std::memcpy(spectra_db + size_t(s) * N, dbf.data(), sizeof(float) * N);
      return detect(dbf.data(), N, c.use_window, c.dc_ignore_window, c.threshold,
                    masks ? masks + size_t(s) * words : nullptr, nullptr, 0);
    }
    for (uint32_t i = 0; i < N; i++) dbd[i] = db_from_power_f64(double(acc[i]) / double(K));
    if (spectra_db64) std::memcpy(spectra_db64 + size_t(s) * N, dbd.data(), sizeof(double) * N);
    if (spectra_db) for (uint32_t i = 0; i < N; i++) spectra_db[size_t(s) * N + i] = float(dbd[i]);
    return detect(dbd.data(), N, c.use_window, c.dc_ignore_window, c.threshold,
                  masks ? masks + size_t(s) * words : nullptr, nullptr, 0);
  }
};

template <typename T>
void pipeline_range(const PipelineConfig& c, const float* window, const uint8_t* raw, uint32_t s_begin,
                    uint32_t s_end, float* spectra_db, double* spectra_db64, uint32_t* masks,
                    uint32_t* counts) {
  PipelineWorker<T> w(c, window);
  for (uint32_t s = s_begin; s < s_end; s++) {
    uint32_t cnt = w.run(raw, s, spectra_db, spectra_db64, masks);
    if (counts) counts[s] = cnt;
  }
}

// "faithful" single-buffer path for the timed CPU baseline: keeps the reference's extra
// copies (messageQueue.h:74-75 memset+memcpy into the message; process.cpp:293-295 memcpy
// into the thread buffer; fft.cpp:22,24 memcpy in/out of the plan buffers).
struct FaithfulWorker {
  PipelineConfig c;
  const float* window;
  FftPlan<float> plan;
  std::vector<float> conv, msg, thr_in, fftw_in, fftw_out, thr_out, mags;
  FaithfulWorker(const PipelineConfig& cfg, const float* w)
      : c(cfg), window(w), plan(cfg.n), conv(2 * size_t(cfg.n)), msg(2 * size_t(cfg.n)),
        thr_in(2 * size_t(cfg.n)), fftw_in(2 * size_t(cfg.n)), fftw_out(2 * size_t(cfg.n)),
        thr_out(2 * size_t(cfg.n)), mags(cfg.n) {}
  uint32_t run(const uint8_t* buf) {
    const uint32_t N = c.n;
    const size_t bytes = sizeof(float) * 2 * size_t(N);
    convert_any(c.kind, buf, conv.data(), N, c.enob, c.correct_dc != 0);   // producer thread
    std::memset(msg.data(), 0, bytes);                                      // messageQueue.h:74
    std::memcpy(msg.data(), conv.data(), bytes);                            // messageQueue.h:75
    std::memcpy(thr_in.data(), msg.data(), bytes);                          // process.cpp:293-295
    if (window) window_apply(thr_in.data(), window, N);                     // process.cpp:296
    std::memcpy(fftw_in.data(), thr_in.data(), bytes);                      // fft.cpp:22
    plan.execute(fftw_in.data(), fftw_out.data());                          // fft.cpp:23
    std::memcpy(thr_out.data(), fftw_out.data(), bytes);                    // fft.cpp:24
    for (uint32_t i = 0; i < N; i++)                                        // utility.cpp:92-97
      mags[i] = db_from_power(power_of(thr_out[2 * i], thr_out[2 * i + 1]), c.db_variant);
    return detect(mags.data(), N, c.use_window, c.dc_ignore_window, c.threshold,
                  static_cast<uint32_t*>(nullptr), nullptr, 0);             // process.cpp:46-61
  }
};

// The same faithful path for eight buffers of one worker, FFTs through FftPlanBatch8 (bit-identical results).
struct FaithfulWorker8 {
  PipelineConfig c;
  const float* window;
  FftPlanBatch8 plan;
  std::vector<float> conv, msg, thr_in, fftw_in, fftw_out, thr_out, mags;
  FaithfulWorker8(const PipelineConfig& cfg, const float* w)
      : c(cfg), window(w), plan(cfg.n), conv(2 * size_t(cfg.n)), msg(16 * size_t(cfg.n)), thr_in(16 * size_t(cfg.n)),
        fftw_in(16 * size_t(cfg.n)), fftw_out(16 * size_t(cfg.n)), thr_out(2 * size_t(cfg.n)), mags(cfg.n) {}
  uint32_t run8(const uint8_t* const bufs[8]) {
    const uint32_t N = c.n;
    const size_t len = 2 * size_t(N), bytes = sizeof(float) * len;
    const float* ins[8];
    float* outs[8];
    for (int l = 0; l < 8; l++) {
      convert_any(c.kind, bufs[l], conv.data(), N, c.enob, c.correct_dc != 0);   // producer thread
      std::memset(msg.data() + l * len, 0, bytes);                                // messageQueue.h:74
      std::memcpy(msg.data() + l * len, conv.data(), bytes);                      // messageQueue.h:75
      std::memcpy(thr_in.data() + l * len, msg.data() + l * len, bytes);          // process.cpp:293-295
      if (window) window_apply(thr_in.data() + l * len, window, N);               // process.cpp:296
      std::memcpy(fftw_in.data() + l * len, thr_in.data() + l * len, bytes);      // fft.cpp:22
      ins[l] = fftw_in.data() + l * len;
      outs[l] = fftw_out.data() + l * len;
    }
    plan.execute(ins, outs);                                                       // fft.cpp:23, eight plans' worth
    uint32_t hits = 0;
    for (int l = 0; l < 8; l++) {
      std::memcpy(thr_out.data(), outs[l], bytes);                                 // fft.cpp:24
      for (uint32_t i = 0; i < N; i++)                                             // utility.cpp:92-97
        mags[i] = db_from_power(power_of(thr_out[2 * i], thr_out[2 * i + 1]), c.db_variant);
      hits += detect(mags.data(), N, c.use_window, c.dc_ignore_window, c.threshold,
                     static_cast<uint32_t*>(nullptr), nullptr, 0);                 // process.cpp:46-61
    }
    return hits;
  }
};

}  // namespace

// ===========================================================================
// C interface (ctypes-friendly)
// ===========================================================================

ORC_API uint32_t orc_bytes_per_sample(uint32_t kind) { return bytes_per_sample(kind); }

ORC_API void orc_convert(uint32_t kind, const void* raw, float* dst, uint32_t n, uint32_t enob,
                         uint32_t correct_dc) {
  convert_any(kind, raw, dst, n, enob, correct_dc != 0);
}

ORC_API void orc_window_build(int type, uint32_t n, float* out) { window_build(type, n, out); }

ORC_API void orc_window_apply(float* iq, const float* w, uint32_t n) { window_apply(iq, w, n); }

ORC_API void orc_fft_f32(const float* in, float* out, uint32_t n) {
  FftPlan<float> plan(n);
  plan.execute(in, out);
}

// eight transforms at once (in / out: [8][n] interleaved complex) -- must equal orc_fft_f32 bit for bit
ORC_API void orc_fft_f32_batch8(const float* in, float* out, uint32_t n) {
  FftPlanBatch8 plan(n);
  const float* ins[8];
  float* outs[8];
  for (int l = 0; l < 8; l++) { ins[l] = in + size_t(l) * 2 * n; outs[l] = out + size_t(l) * 2 * n; }
  plan.execute(ins, outs);
}

ORC_API void orc_fft_f64(const double* in, double* out, uint32_t n) {
  FftPlan<double> plan(n);
  plan.execute(in, out);
}

ORC_API void orc_magnitude_db(const float* fft, float* db, uint32_t n, int variant) {
  for (uint32_t i = 0; i < n; i++) db[i] = db_from_power(power_of(fft[2 * i], fft[2 * i + 1]), variant);
}

ORC_API uint32_t orc_use_window(double use_bandwidth, uint32_t n) {
  return uint32_t(use_bandwidth * n / 2.0);   // process.cpp:85
}

ORC_API uint32_t orc_detect(const float* db, uint32_t n, uint32_t use_window, uint32_t dc_ignore,
                            float threshold, uint32_t* mask, uint32_t* hit_bins, uint32_t hit_cap) {
  return detect(db, n, use_window, dc_ignore, threshold, mask, hit_bins, hit_cap);
}

ORC_API uint64_t orc_hit_frequency(double center, uint32_t sample_rate, uint32_t n, uint32_t i) {
  return hit_frequency(center, sample_rate, n, i);
}

ORC_API uint32_t orc_frequency_table(uint32_t sample_rate, double start, double stop, double use_bw,
                                     double dc_ignore, double* out, uint32_t cap) {
  return frequency_table(sample_rate, start, stop, use_bw, dc_ignore, out, cap);
}

// Batch pipeline.  raw: n_spectra * averaging * N samples of `kind`, contiguous.
// precision 0 = fp32 FFT (reference-like), 1 = fp64 FFT (ground truth; spectra_db64 optional).
ORC_API void orc_pipeline(uint32_t n, uint32_t sample_rate, uint32_t enob, uint32_t kind,
                          uint32_t correct_dc, uint32_t averaging, float threshold,
                          uint32_t use_window, uint32_t dc_ignore_window, int db_variant,
                          const float* window, const void* raw, uint32_t n_spectra, int precision,
                          float* spectra_db, double* spectra_db64, uint32_t* masks, uint32_t* counts,
                          uint32_t threads) {
  PipelineConfig c{n, sample_rate, enob, kind, correct_dc, averaging ? averaging : 1, threshold,
                   use_window, dc_ignore_window, db_variant};
  if (threads == 0) threads = 1;
  if (threads > n_spectra) threads = n_spectra ? n_spectra : 1;
  std::vector<std::thread> pool;
  for (uint32_t t = 0; t < threads; t++) {
    uint32_t b = uint32_t(uint64_t(n_spectra) * t / threads);
    uint32_t e = uint32_t(uint64_t(n_spectra) * (t + 1) / threads);
    auto fn = [=]() {
      if (precision == 0)
        pipeline_range<float>(c, window, static_cast<const uint8_t*>(raw), b, e, spectra_db, nullptr,
                              masks, counts);
      else
        pipeline_range<double>(c, window, static_cast<const uint8_t*>(raw), b, e, spectra_db,
                               spectra_db64, masks, counts);
    };
    if (threads == 1) fn(); else pool.emplace_back(fn);
  }
  for (auto& th : pool) th.join();
}

// Timed CPU baseline.  Processes `n_buffers` raw buffers `repeats` times on `threads`
// workers (one FFT plan per worker -- NOT the reference's shared, racy one, fft.cpp:20-25)
// and returns elapsed seconds.  faithful != 0 keeps the reference's per-buffer copies,
// faithful == 0 is a fused loop over the same arithmetic (K averaging honoured).
// total_hits (optional) receives the hit total so the work cannot be optimised away.
ORC_API double orc_bench(uint32_t n, uint32_t sample_rate, uint32_t enob, uint32_t kind,
                         uint32_t correct_dc, uint32_t averaging, float threshold, uint32_t use_window,
                         uint32_t dc_ignore_window, const float* window, const void* raw,
                         uint32_t n_buffers, uint32_t repeats, uint32_t threads, int faithful,
                         uint64_t* total_hits) {
  PipelineConfig c{n, sample_rate, enob, kind, correct_dc, averaging ? averaging : 1, threshold,
                   use_window, dc_ignore_window, 0};
  if (threads == 0) threads = std::thread::hardware_concurrency();
  if (threads == 0) threads = 1;
  const size_t buf_bytes = size_t(n) * bytes_per_sample(kind);
  std::atomic<uint64_t> hits{0};
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (uint32_t t = 0; t < threads; t++) {
    pool.emplace_back([&, t]() {
      uint64_t local = 0;
      if (faithful == 2) {
        // faithful path, FFTs eight at a time per worker (same results, SIMD-library-class FFT speed)
        FaithfulWorker8 w8(c, window);
        FaithfulWorker w1(c, window);
        const uint8_t* base = static_cast<const uint8_t*>(raw);
        for (uint32_t rep = 0; rep < repeats; rep++) {
          uint32_t b = t * 8;
          for (; b + 8 <= n_buffers; b += threads * 8) {
            const uint8_t* bufs[8];
            for (int l = 0; l < 8; l++) bufs[l] = base + size_t(b + l) * buf_bytes;
            local += w8.run8(bufs);
          }
          if (t == 0)
            for (uint32_t r = n_buffers - n_buffers % 8; r < n_buffers; r++) local += w1.run(base + size_t(r) * buf_bytes);
        }
      } else if (faithful) {
        FaithfulWorker w(c, window);
        for (uint32_t rep = 0; rep < repeats; rep++)
          for (uint32_t b = t; b < n_buffers; b += threads)
            local += w.run(static_cast<const uint8_t*>(raw) + size_t(b) * buf_bytes);
      } else {
        const uint32_t n_spectra = n_buffers / c.averaging;
        PipelineWorker<float> w(c, window);
        for (uint32_t rep = 0; rep < repeats; rep++)
          for (uint32_t s = t; s < n_spectra; s += threads)
            local += w.run(static_cast<const uint8_t*>(raw), s, nullptr, nullptr, nullptr);
      }
      hits += local;
    });
  }
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  if (total_hits) *total_hits = hits.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Time-domain mode -- ProcessSamples::DoTimeDomainThresholding, process.cpp:203-237, applied to
// the converted (not windowed) samples of each buffer.  max is seeded with
// numeric_limits<float>::min() (smallest positive normal, process.cpp:207), min with max()
// (process.cpp:208); trigger iff maxMagnitude >= threshold (process.cpp:226).
ORC_API void orc_time_domain(uint32_t n, uint32_t enob, uint32_t kind, uint32_t correct_dc, float threshold,
                             const void* raw, uint32_t n_buffers, uint32_t* trigger, float* max_min) {
  const double log10v = std::log2(10.0);
  std::vector<float> iq(2 * size_t(n));
  const size_t buf_bytes = size_t(n) * bytes_per_sample(kind);
  for (uint32_t b = 0; b < n_buffers; b++) {
    convert_any(kind, static_cast<const uint8_t*>(raw) + size_t(b) * buf_bytes, iq.data(), n, enob,
                correct_dc != 0);
    float maxm = std::numeric_limits<float>::min();
    float minm = std::numeric_limits<float>::max();
    for (uint32_t i = 0; i < n; i++) {
      float re = iq[2 * i], im = iq[2 * i + 1];
      float mag = std::sqrt(re * re + im * im);
      float magnitude = float(10 * std::log2(mag) / log10v);
      maxm = std::max(maxm, magnitude);
      minm = std::min(minm, magnitude);
    }
    if (trigger) trigger[b] = (maxm >= threshold) ? 1u : 0u;
    if (max_min) { max_min[2 * size_t(b)] = maxm; max_min[2 * size_t(b) + 1] = minm; }
  }
}

// HackRF sweep-frame pre-pass -- HackRFSource::interpolateSamples, hackRFSource.cpp:186-222, restated
// with its quirks: the block loop runs valid_length/2/8192 times but always inspects the FIRST block
// (`ubuf` is never advanced, :192), so iterations after the first only act when the patched bytes
// themselves read 0x7F 0x7F (sample 5 saturated); the patch value is sample 5, averaged for i > 0 with
// sample i-1 in int arithmetic (truncating toward zero, :209-210) and narrowed to int8; samples 0..4
// are overwritten in place (:213-216).  frequency_hz[t] = the last header value (0 if none; the
// source adds m_scanOffset, :221); status[t] bit 0 = a header was seen, bits 8.. = number of
// "frequencyHz != thisFrequencyHz" prints (:202-206).
ORC_API void orc_hackrf_prepass(uint8_t* transfers, uint32_t n_transfers, uint32_t valid_length,
                                uint64_t* frequency_hz, uint32_t* status) {
  for (uint32_t t = 0; t < n_transfers; t++) {
    uint8_t* ubuf = transfers + size_t(t) * valid_length;
    const uint32_t count = valid_length / 2;
    uint64_t freq = 0;
    uint32_t st = 0;
    for (uint32_t i = 0; i < count; i += 8192) {
      if (ubuf[0] == 0x7F && ubuf[1] == 0x7F) {
        uint64_t cur = 0;
        for (int k = 7; k >= 0; k--) cur = (cur << 8) | ubuf[2 + k];
        if (freq != 0 && freq != cur) st += 1u << 8;
        freq = cur;
        st |= 1u;
        int8_t post[2] = {int8_t(ubuf[10]), int8_t(ubuf[11])};
        if (i > 0) {
          post[0] = int8_t((int(post[0]) + int(int8_t(ubuf[2 * (i - 1)]))) / 2);
          post[1] = int8_t((int(post[1]) + int(int8_t(ubuf[2 * (i - 1) + 1]))) / 2);
        }
        for (uint32_t j = 0; j < 5; j++) { ubuf[2 * j] = uint8_t(post[0]); ubuf[2 * j + 1] = uint8_t(post[1]); }
      }
    }
    if (frequency_hz) frequency_hz[t] = freq;
    if (status) status[t] = st;
  }
}

ORC_API uint32_t orc_hardware_threads() { return std::thread::hardware_concurrency(); }
