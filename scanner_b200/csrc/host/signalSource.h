// SignalSource -- abstract device source plugin.  Same shape as the reference's SignalSource
// (signalSource.h:9-68): a source owns the FrequencyTable, runs a producer thread that hands raw
// buffers to a SampleQueue, and stops after `numIterations` sweeps or StopStreaming().
// Vendor SDK sources (HackRF, B210, BladeRF, Airspy, SDRplay, RTL) plug in by overriding the four
// pure virtuals exactly as they do in the reference; SyntheticSource / ReplaySource are the ones
// this repo ships (no SDR hardware in the loop).
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
 protected:
  bool m_doTiming;
  struct timespec m_start, m_stop;
  double m_elapsedTime = 0.0;
  uint32_t m_retuneTimeIndex = 0;
  uint32_t m_getSamplesTimeIndex = 0;
  std::atomic<bool> m_isDone{false};      // the reference leaves this uninitialised (signalSource.cpp:8-30)
  std::atomic<bool> m_finished{false};
  bool m_synchronousMode = false;
  std::unique_ptr<std::thread> m_thread;
  std::vector<double> m_retuneTime;
  std::vector<double> m_getSamplesTime;
  static const uint32_t s_maxIndex = 10000;

  uint32_t m_sampleRate;
  uint32_t m_sampleCount;
  double m_startFrequency;
  double m_stopFrequency;
  FrequencyTable m_frequencyTable;
  uint32_t m_iterationLimit = 0;
  SampleQueue* m_sampleQueue = nullptr;

  void SetIsDone();
  bool StopThread();
  bool StartThread(uint32_t numIterations, SampleQueue& sampleQueue);
  void ThreadWorkerHelper();
  uint32_t GetIterationCount();
  double GetCurrentFrequency(void** pinfo = nullptr);
  double GetNextFrequency(void** pinfo = nullptr);
  double GetStartFrequency();
  double GetStopFrequency();
  bool GetIsDone();

 public:
  SignalSource(uint32_t sampleRate, uint32_t sampleCount, double startFrequency, double stopFrequency,
               double useBandWidth = 0.75, double dcIgnoreWidth = 0.0, bool doTiming = false);
  virtual ~SignalSource();
  virtual bool Start();
  virtual bool GetNextSamples(SampleQueue* sampleQueue, double_t& centerFrequency) = 0;
  virtual bool StartStreaming(uint32_t numIterations, SampleQueue& sampleQueue) = 0;
  virtual void ThreadWorker() = 0;
  virtual bool Stop();
  virtual double Retune(double frequency) = 0;
  bool DoRetune();
  uint32_t GetFrequencyCount();
  bool GetIsScanStart();
  void StopStreaming();
  void Join();                       // wait for the producer thread to finish its sweeps
  void StartTimer();
  void StopTimer();
  void AddRetuneTime();
  void AddGetSamplesTime();
  void WriteTimingData();
  FrequencyTable& GetFrequencyTable() { return m_frequencyTable; }
};
