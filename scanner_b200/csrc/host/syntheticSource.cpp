#include "syntheticSource.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct Rng {   // counter-based stream: key fixed per buffer, counter per draw
  uint64_t key, ctr = 0;
  explicit Rng(uint64_t k) : key(k) {}
  uint64_t next() { return splitmix64(key ^ splitmix64(ctr++)); }
  double uniform() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)
  void gauss2(double& a, double& b) {                                              // Box-Muller pair
    const double u1 = 1.0 - uniform(), u2 = uniform();
    const double r = std::sqrt(-2.0 * std::log(u1)), t = 6.283185307179586 * u2;
    a = r * std::cos(t);
    b = r * std::sin(t);
  }
};

constexpr double kTwoPi = 6.283185307179586476925286766559;

}  // namespace

size_t SyntheticSource::BufferBytes(SampleQueue::SampleKind kind, uint32_t n) {
  switch (kind) {
    case SampleQueue::ByteComplex: return size_t(n) * 2;
    case SampleQueue::Short:
    case SampleQueue::ShortComplex: return size_t(n) * 4;
    case SampleQueue::FloatComplex: return size_t(n) * 8;
    default: return 0;
  }
}

SyntheticSource::SyntheticSource(SampleQueue::SampleKind kind, uint32_t enob, uint64_t seed,
                                 uint32_t buffersPerStep, uint32_t sampleRate, uint32_t sampleCount,
                                 double startFrequency, double stopFrequency, double useBandWidth,
                                 double dcIgnoreWidth)
    : SignalSource(sampleRate, sampleCount, startFrequency, stopFrequency, useBandWidth, dcIgnoreWidth),
      m_kind(kind), m_enob(enob), m_seed(seed), m_buffersPerStep(buffersPerStep ? buffersPerStep : 1),
      m_useBandWidth(useBandWidth) {}

void SyntheticSource::Generate(uint32_t sweep, uint32_t step, uint32_t buffer, void* raw) const {
  const uint32_t N = m_sampleCount;
  Rng rng(splitmix64(m_seed) ^ splitmix64((uint64_t(sweep) << 40) ^ (uint64_t(step) << 20) ^ buffer));
  const double fullScale = (m_kind == SampleQueue::FloatComplex) ? 1.0 : double((1u << (m_enob - 1)) - 1);
  const uint32_t useW = uint32_t(m_useBandWidth * N / 2.0);
  const uint32_t nTones = uint32_t(rng.next() % 5);
  double amp[4], bin[4], phase[4];
  for (uint32_t k = 0; k < 4; k++) {
    const uint32_t i = N / 2 - useW + uint32_t(rng.next() % (2 * useW + 1));   // shifted bin
    bin[k] = double((i + N / 2) % N);                                           // FFT bin
    amp[k] = (k < nTones) ? std::pow(10.0, (-40.0 + 37.0 * rng.uniform()) / 20.0) : 0.0;
    phase[k] = kTwoPi * rng.uniform();
  }
  const double lo = (m_kind == SampleQueue::ByteComplex) ? -128.0 : -32768.0;
  const double hi = (m_kind == SampleQueue::ByteComplex) ? 127.0 : 32767.0;
  for (uint32_t n = 0; n < N; n++) {
    double re, im;
    rng.gauss2(re, im);
    re = 0.05 * re + 0.01;
    im = 0.05 * im + 0.01;
    for (uint32_t k = 0; k < nTones; k++) {
      const double a = kTwoPi * std::fmod(bin[k] * n, double(N)) / N + phase[k];
      re += amp[k] * std::cos(a);
      im += amp[k] * std::sin(a);
    }
    if (m_kind == SampleQueue::FloatComplex) {
      static_cast<float*>(raw)[2 * n] = float(re);
      static_cast<float*>(raw)[2 * n + 1] = float(im);
      continue;
    }
    const double qr = std::min(std::max(std::nearbyint(re * fullScale), std::max(lo, -fullScale - 1)), std::min(hi, fullScale));
    const double qi = std::min(std::max(std::nearbyint(im * fullScale), std::max(lo, -fullScale - 1)), std::min(hi, fullScale));
    switch (m_kind) {
      case SampleQueue::ByteComplex:
        static_cast<int8_t*>(raw)[2 * n] = int8_t(qr);
        static_cast<int8_t*>(raw)[2 * n + 1] = int8_t(qi);
        break;
      case SampleQueue::ShortComplex:
        static_cast<int16_t*>(raw)[2 * n] = int16_t(qr);
        static_cast<int16_t*>(raw)[2 * n + 1] = int16_t(qi);
        break;
      default:   // Short: re block then im block
        static_cast<int16_t*>(raw)[n] = int16_t(qr);
        static_cast<int16_t*>(raw)[N + n] = int16_t(qi);
        break;
    }
  }
}

bool SyntheticSource::StartStreaming(uint32_t numIterations, SampleQueue& sampleQueue) {
  return StartThread(numIterations, sampleQueue);
}

double SyntheticSource::Retune(double frequency) {
  m_currentFrequency = frequency;   // nothing to settle: no LO
  return frequency;
}

static void AppendRaw(SampleQueue* q, SampleQueue::SampleKind kind, void* raw, uint32_t n, double f, time_t t) {
  switch (kind) {
    case SampleQueue::ByteComplex: q->AppendSamples(static_cast<int8_t(*)[2]>(raw), f, t); break;
    case SampleQueue::ShortComplex: q->AppendSamples(static_cast<int16_t(*)[2]>(raw), f, t); break;
    case SampleQueue::Short: q->AppendSamples(static_cast<int16_t*>(raw), static_cast<int16_t*>(raw) + n, f, t); break;
    default: q->AppendSamples(static_cast<fftwf_complex*>(raw), f, t); break;
  }
}

bool SyntheticSource::GetNextSamples(SampleQueue* sampleQueue, double_t& centerFrequency) {
  if (GetIsDone()) return false;
  std::vector<char> raw(BufferBytes(m_kind, m_sampleCount));
  centerFrequency = GetCurrentFrequency();
  const bool isScanStart = GetIsScanStart();
  for (uint32_t b = 0; b < m_buffersPerStep; b++) {
    Generate(GetIterationCount(), m_frequencyTable.GetCurrentIndex(), b, raw.data());
    AppendRaw(sampleQueue, m_kind, raw.data(), m_sampleCount, centerFrequency,
              (isScanStart && b == 0) ? time(nullptr) : 0);
  }
  Retune(GetNextFrequency());
  return true;
}

void SyntheticSource::ThreadWorker() {
  double_t f;
  while (!GetIsDone()) GetNextSamples(m_sampleQueue, f);
}

// ---- ReplaySource ---------------------------------------------------------------------------------------

ReplaySource::ReplaySource(SampleQueue::SampleKind kind, const void* raw, const double* frequencies,
                           size_t nBuffers, uint32_t buffersPerSweep, uint32_t sampleRate, uint32_t sampleCount)
    : SignalSource(sampleRate, sampleCount, nBuffers ? frequencies[0] : 0.0, 0.0),
      m_kind(kind), m_raw(static_cast<const char*>(raw)), m_frequencies(frequencies), m_nBuffers(nBuffers),
      m_buffersPerSweep(buffersPerSweep), m_bufferBytes(SyntheticSource::BufferBytes(kind, sampleCount)) {}

void ReplaySource::Append(SampleQueue* q, size_t b) {
  const time_t t = (m_buffersPerSweep && (b % m_buffersPerSweep) == 0) ? time_t(1000000000 + b) : 0;
  AppendRaw(q, m_kind, const_cast<char*>(m_raw) + b * m_bufferBytes, m_sampleCount, m_frequencies[b], t);
}

bool ReplaySource::GetNextSamples(SampleQueue* sampleQueue, double_t& centerFrequency) {
  if (m_repeatTotal) {                       // capture ring: wrap until the requested total has been streamed
    if (m_streamed >= m_repeatTotal || m_isDone || m_nBuffers == 0) return false;
    if (m_next >= m_nBuffers) m_next = 0;
  }
  if (m_next >= m_nBuffers || m_isDone) return false;
  centerFrequency = m_frequencies[m_next];
  if (m_appendBatch > 1 && m_kind != SampleQueue::Short) {
    size_t n = m_nBuffers - m_next < m_appendBatch ? m_nBuffers - m_next : m_appendBatch;
    if (m_repeatTotal && m_repeatTotal - m_streamed < n) n = size_t(m_repeatTotal - m_streamed);
    m_streamed += n;
    const time_t* times = nullptr;
    if (m_buffersPerSweep) {
      m_times.assign(n, 0);
      for (size_t i = 0; i < n; i++)
        if ((m_next + i) % m_buffersPerSweep == 0) m_times[i] = time_t(1000000000 + m_next + i);
      times = m_times.data();
    }
    sampleQueue->AppendSamplesBatch(m_raw + m_next * m_bufferBytes, uint32_t(n), m_frequencies + m_next, times);
    m_next += n;
    return true;
  }
  m_streamed++;
  Append(sampleQueue, m_next++);
  return true;
}

bool ReplaySource::StartStreaming(uint32_t numIterations, SampleQueue& sampleQueue) {
  return StartThread(numIterations, sampleQueue);
}

void ReplaySource::ThreadWorker() {
  if (m_localCopy) {
    m_local.assign(m_raw, m_raw + m_nBuffers * m_bufferBytes);
    m_localFrequencies.assign(m_frequencies, m_frequencies + m_nBuffers);
    m_raw = m_local.data();
    m_frequencies = m_localFrequencies.data();
  }
  double_t f;
  while (GetNextSamples(m_sampleQueue, f)) {}
}

double ReplaySource::Retune(double frequency) { return frequency; }
