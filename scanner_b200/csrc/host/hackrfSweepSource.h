// HackRFSweepSource -- the reference's HackRF source (hackRFSource.h:5-47, hackRFSource.cpp) with the
// libhackrf device replaced by a captured sweep stream.  Same constructor arguments and the same
// SignalSource overrides; what libhackrf's rx thread would deliver to `hackRF_rx_callback`
// (hackRFSource.cpp:224-264) is replayed from memory by ThreadWorker instead.
//
// Kept from the reference, line for line in behaviour:
//   * sweep parameters (hackRFSource.cpp:108-112): step width uint32(0.75 fs), offset uint32(step / 2.0);
//   * InterpolateSamples (:186-222): frame header 0x7F 0x7F + LE64 frequency parsed from the FIRST block only
//     (the loop never advances its block pointer), samples 0..4 overwritten with sample 5, the later-iteration
//     quirk when sample 5 is saturated, the "frequencyHz != thisFrequencyHz" print; returns
//     double(frequency + offset);
//   * RxCallback (:224-264): a change of centre frequency advances the FrequencyTable (which is what counts
//     sweeps), the transfer that wraps the table is stamped as scan start, and EVERY sampleCount chunk of
//     that transfer is appended with the stamp -- so the queue's "drop until the second scan start"
//     (messageQueue.h:67-72) drops exactly one chunk.
// The 10-byte patch runs on the producer thread as in the reference; for sweep streams that are already
// device resident the same pre-pass exists as a kernel (scn_hackrf_prepass_device).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "signalSource.h"

class HackRFSweepSource : public SignalSource {
 public:
  HackRFSweepSource(std::string args, uint32_t sampleRate, uint32_t sampleCount, double startFrequency,
                    double stopFrequency);
  bool GetNextSamples(SampleQueue* sampleQueue, double_t& centerFrequency) override;
  bool StartStreaming(uint32_t numIterations, SampleQueue& sampleQueue) override;
  bool Start() override;
  void ThreadWorker() override;
  double Retune(double frequency) override;

  // The transfers ThreadWorker replays (patched in place, like libhackrf's buffers are).
  void SetCapture(uint8_t* stream, size_t bytes, uint32_t validLength);
  // Scan-start stamps: the k-th stamp is base + k * step instead of time(NULL) (reproducible replays).
  void SetReplayClock(time_t base, time_t step);

  double InterpolateSamples(uint8_t* buffer, uint32_t validLength);
  int RxCallback(uint8_t* buffer, uint32_t validLength);
  uint32_t GetScanStepWidth() const { return m_scanStepWidth; }
  uint32_t GetScanOffset() const { return m_scanOffset; }
  bool GetStreamDone() const { return m_streamingState == Done; }

 private:
  enum StreamingState { Illegal = 0, Streaming, DoRetune, Done };
  std::atomic<int> m_streamingState{Illegal};
  double m_centerFrequency = 1e12;             // hackRFSource.cpp:44
  uint16_t m_scanStartFrequency, m_scanStopFrequency;
  uint32_t m_scanNumBytes, m_scanStepWidth, m_scanOffset;
  uint8_t* m_capture = nullptr;
  size_t m_captureBytes = 0;
  uint32_t m_validLength = 0;
  bool m_replayClock = false;
  time_t m_clockBase = 0, m_clockStep = 0;
  uint64_t m_clockCalls = 0;
  time_t Now();
};
