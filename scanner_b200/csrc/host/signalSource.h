// SignalSource -- abstract device source plugin, the boundary SDR back ends are written against.
// Interface parity with the reference (signalSource.h:9-68): a source owns the FrequencyTable of its sweep,
// runs one producer thread that hands raw buffers to a SampleQueue, and stops after `numIterations` sweeps or
// on StopStreaming().  A vendor back end (HackRF, B210, BladeRF, Airspy, SDRplay, RTL) overrides
// GetNextSamples / StartStreaming / ThreadWorker / Retune exactly as it does in the reference; this repo ships
// SyntheticSource, ReplaySource and HackRFSweepSource (no SDR hardware in the loop).
#pragma once
#include <atomic>
#include <cmath>
#include <cstdint>
#include <ctime>
#include <memory>
#include <thread>
#include <vector>

#include "frequencyTable.h"
#include "sampleQueue.h"

class SignalSource {
 public:
  SignalSource(uint32_t sampleRate, uint32_t sampleCount, double startFrequency, double stopFrequency,
               double useBandWidth = 0.75, double dcIgnoreWidth = 0.0, bool doTiming = false);
  virtual ~SignalSource();

  // ---- what a back end implements (signalSource.h:53-58)
  virtual bool GetNextSamples(SampleQueue* sampleQueue, double_t& centerFrequency) = 0;
  virtual bool StartStreaming(uint32_t numIterations, SampleQueue& sampleQueue) = 0;
  virtual void ThreadWorker() = 0;
  virtual double Retune(double frequency) = 0;
  virtual bool Start();
  virtual bool Stop();

  // ---- what the application calls
  void StopStreaming();
  void Join();                                   // wait for the producer thread to finish its sweeps
  bool DoRetune();                               // synchronous mode: retune only once the consumer has acked
  bool GetIsScanStart();
  uint32_t GetFrequencyCount();
  FrequencyTable& GetFrequencyTable() { return m_frequencyTable; }

  // ---- optional retune / acquisition timing (signalSource.cpp:132-190)
  void StartTimer();
  void StopTimer();
  void AddRetuneTime();
  void AddGetSamplesTime();
  void WriteTimingData();

 protected:
  // helpers for back ends (signalSource.h:33-42)
  bool StartThread(uint32_t numIterations, SampleQueue& sampleQueue);
  bool StopThread();
  void ThreadWorkerHelper();
  void SetIsDone();
  bool GetIsDone();
  uint32_t GetIterationCount();
  double GetCurrentFrequency(void** pinfo = nullptr);
  double GetNextFrequency(void** pinfo = nullptr);
  double GetStartFrequency();
  double GetStopFrequency();

  // sweep definition
  uint32_t m_sampleRate, m_sampleCount;
  double m_startFrequency, m_stopFrequency;
  FrequencyTable m_frequencyTable;
  uint32_t m_iterationLimit = 0;
  SampleQueue* m_sampleQueue = nullptr;

  // producer thread state (the reference leaves m_isDone uninitialised, signalSource.cpp:8-30)
  std::atomic<bool> m_isDone{false};
  std::atomic<bool> m_finished{false};
  bool m_synchronousMode = false;
  std::unique_ptr<std::thread> m_thread;

  // timing
  static const uint32_t s_maxIndex = 10000;
  bool m_doTiming;
  struct timespec m_start, m_stop;
  double m_elapsedTime = 0.0;
  uint32_t m_retuneTimeIndex = 0, m_getSamplesTimeIndex = 0;
  std::vector<double> m_retuneTime, m_getSamplesTime;
};
