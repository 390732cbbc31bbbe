// SyntheticSource / ReplaySource -- SignalSource plugins that stand in for SDR hardware.
//
// SyntheticSource generates seeded IQ per (retune step, buffer): complex white Gaussian noise
// (sigma = 0.05 FS per rail) + up to 4 complex tones at random shifted-bin centres inside the used
// band, amplitudes log-uniform in [-40, -3] dBFS, +0.01 FS DC, quantised round-to-nearest with
// saturation.  Counter-based generator (splitmix64 of seed/step/buffer/sample),
// so any buffer can be regenerated independently -- which is what lets ranks own disjoint retune steps.
// Its ThreadWorker follows the pattern of the reference's sync sources (bladerfSource.cpp:256-300):
// while (!GetIsDone()) { f = GetCurrentFrequency(); fill; GetNextFrequency(); AppendSamples(..., isScanStart ? time : 0); }
//
// ReplaySource feeds recorded raw buffers + their centre frequencies from memory (tests, captures).
#pragma once
#include <cstdint>
#include <vector>

#include "signalSource.h"

class SyntheticSource : public SignalSource {
 public:
  SyntheticSource(SampleQueue::SampleKind kind, uint32_t enob, uint64_t seed, uint32_t buffersPerStep,
                  uint32_t sampleRate, uint32_t sampleCount, double startFrequency, double stopFrequency,
                  double useBandWidth = 0.75, double dcIgnoreWidth = 0.0);
  bool GetNextSamples(SampleQueue* sampleQueue, double_t& centerFrequency) override;
  bool StartStreaming(uint32_t numIterations, SampleQueue& sampleQueue) override;
  void ThreadWorker() override;
  double Retune(double frequency) override;

  // Fills `raw` (one buffer of the source's kind) for (sweep, step, buffer); deterministic.
  void Generate(uint32_t sweep, uint32_t step, uint32_t buffer, void* raw) const;
  static size_t BufferBytes(SampleQueue::SampleKind kind, uint32_t sampleCount);

 private:
  SampleQueue::SampleKind m_kind;
  uint32_t m_enob;
  uint64_t m_seed;
  uint32_t m_buffersPerStep;
  double m_useBandWidth;
  double m_currentFrequency = 0.0;
};

class ReplaySource : public SignalSource {
 public:
  // raw: nBuffers contiguous buffers; frequencies: nBuffers centre frequencies; buffersPerSweep marks
  // scan starts (time != 0 on the first buffer of each sweep), 0 == never.
  ReplaySource(SampleQueue::SampleKind kind, const void* raw, const double* frequencies, size_t nBuffers,
               uint32_t buffersPerSweep, uint32_t sampleRate, uint32_t sampleCount);
  bool GetNextSamples(SampleQueue* sampleQueue, double_t& centerFrequency) override;
  bool StartStreaming(uint32_t numIterations, SampleQueue& sampleQueue) override;
  void ThreadWorker() override;
  double Retune(double frequency) override;
  // Hand the queue `n` consecutive buffers per call (SampleQueue::AppendSamplesBatch) the way an SDR driver delivers
  // a whole USB transfer at once (hackRFSource.cpp:251-264: 64 buffers per 262 144-byte transfer); 1 == one
  // AppendSamples call per buffer.  Interleaved kinds only.
  void SetAppendBatch(uint32_t n) { m_appendBatch = n ? n : 1; }
  // Stream `totalBuffers` buffers by cycling over the recorded ones (a capture ring replayed for as long as asked).
  void SetRepeat(uint64_t totalBuffers) { m_repeatTotal = totalBuffers; }
  // Copy the recording into memory first touched by the producer thread itself before streaming from it (keeps the
  // source side of the hand-off's memcpy local to the core that runs it on multi-socket hosts).
  void SetLocalCopy(bool on) { m_localCopy = on; }

 private:
  void Append(SampleQueue* q, size_t b);
  uint32_t m_appendBatch = 1;
  std::vector<time_t> m_times;
  uint64_t m_repeatTotal = 0, m_streamed = 0;
  bool m_localCopy = false;
  std::vector<char> m_local;
  std::vector<double> m_localFrequencies;
  SampleQueue::SampleKind m_kind;
  const char* m_raw;
  const double* m_frequencies;
  size_t m_nBuffers;
  uint32_t m_buffersPerSweep;
  size_t m_next = 0;
  size_t m_bufferBytes;
};
