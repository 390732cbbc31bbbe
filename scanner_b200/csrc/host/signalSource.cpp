#include "signalSource.h"

#include <cstdio>

SignalSource::SignalSource(uint32_t sampleRate, uint32_t sampleCount, double startFrequency,
                           double stopFrequency, double useBandWidth, double dcIgnoreWidth, bool doTiming)
    : m_sampleRate(sampleRate), m_sampleCount(sampleCount), m_startFrequency(startFrequency),
      m_stopFrequency(stopFrequency),
      m_frequencyTable(sampleRate, startFrequency, stopFrequency, useBandWidth, dcIgnoreWidth),
      m_doTiming(doTiming), m_retuneTime(doTiming ? s_maxIndex : 0), m_getSamplesTime(doTiming ? s_maxIndex : 0) {}

SignalSource::~SignalSource() {
  if (m_thread && m_thread->joinable()) {
    m_isDone = true;
    m_thread->join();
  }
}

bool SignalSource::Start() { return true; }
bool SignalSource::Stop() { return true; }

double SignalSource::GetNextFrequency(void** pinfo) { return m_frequencyTable.GetNextFrequency(pinfo); }
double SignalSource::GetCurrentFrequency(void** pinfo) { return m_frequencyTable.GetCurrentFrequency(pinfo); }
double SignalSource::GetStartFrequency() { return m_frequencyTable.GetStartFrequency(); }
double SignalSource::GetStopFrequency() { return m_frequencyTable.GetStopFrequency(); }
uint32_t SignalSource::GetFrequencyCount() { return m_frequencyTable.GetFrequencyCount(); }
bool SignalSource::GetIsScanStart() { return m_frequencyTable.GetIsScanStart(); }
uint32_t SignalSource::GetIterationCount() { return m_frequencyTable.GetIterationCount(); }

bool SignalSource::DoRetune() {
  // synchronous mode gates the retune on the consumer's ack (signalSource.cpp:75-81)
  if (m_synchronousMode && m_sampleQueue != nullptr) return m_sampleQueue->ReceivedAck();
  return true;
}

bool SignalSource::StartThread(uint32_t numIterations, SampleQueue& sampleQueue) {
  printf("Starting source thread...\n");
  m_iterationLimit = numIterations;
  m_sampleQueue = &sampleQueue;
  m_finished = false;
  m_thread.reset(new std::thread(&SignalSource::ThreadWorkerHelper, this));
  return true;
}

bool SignalSource::StopThread() {
  if (m_thread) {
    printf("Stopping source thread...\n");
    m_finished = true;
    if (m_thread->joinable()) m_thread->join();
  }
  return true;
}

void SignalSource::Join() {
  if (m_thread && m_thread->joinable()) m_thread->join();
}

bool SignalSource::GetIsDone() { return GetIterationCount() >= m_iterationLimit || m_isDone; }
void SignalSource::SetIsDone() { m_isDone = true; }

void SignalSource::ThreadWorkerHelper() {
  ThreadWorker();
  m_sampleQueue->SetIsDone();     // lets the consumers drain and exit (signalSource.cpp:123)
}

void SignalSource::StopStreaming() {
  SetIsDone();
  StopThread();
}

void SignalSource::StartTimer() {
  if (m_doTiming) clock_gettime(CLOCK_REALTIME, &m_start);
}

void SignalSource::StopTimer() {
  if (!m_doTiming) return;
  clock_gettime(CLOCK_REALTIME, &m_stop);
  m_elapsedTime = (m_stop.tv_sec * 1000.0 + m_stop.tv_nsec / 1e6) - (m_start.tv_sec * 1000.0 + m_start.tv_nsec / 1e6);
}

void SignalSource::AddRetuneTime() {
  if (m_doTiming && m_retuneTimeIndex < s_maxIndex) m_retuneTime[m_retuneTimeIndex++] = m_elapsedTime;
}

void SignalSource::AddGetSamplesTime() {
  if (m_doTiming && m_getSamplesTimeIndex < s_maxIndex) m_getSamplesTime[m_getSamplesTimeIndex++] = m_elapsedTime;
}

void SignalSource::WriteTimingData() {
  if (!m_doTiming) return;
  if (FILE* f = fopen("timings.txt", "w")) {
    for (uint32_t i = 0; i < m_retuneTimeIndex; i++)
      fprintf(f, "%f, %f\n", m_retuneTime[i], i < m_getSamplesTimeIndex ? m_getSamplesTime[i] : 0.0);
    fclose(f);
  }
  m_doTiming = false;
}
