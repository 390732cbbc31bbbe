#include "hackrfSweepSource.h"

#include <cassert>
#include <cstdio>

HackRFSweepSource::HackRFSweepSource(std::string args, uint32_t sampleRate, uint32_t sampleCount,
                                     double startFrequency, double stopFrequency)
    : SignalSource(sampleRate, sampleCount, startFrequency, stopFrequency, 0.75, 0.0) {
  (void)args;                                        // gain / bias options address the hardware only
  m_scanStartFrequency = uint16_t(startFrequency / 1e6);
  m_scanStopFrequency = uint16_t(stopFrequency / 1e6);
  m_scanNumBytes = sampleCount * 2;
  m_scanStepWidth = 0.75 * sampleRate;
  m_scanOffset = m_scanStepWidth / 2.0;
}

void HackRFSweepSource::SetCapture(uint8_t* stream, size_t bytes, uint32_t validLength) {
  m_capture = stream;
  m_captureBytes = bytes;
  m_validLength = validLength;
}

void HackRFSweepSource::SetReplayClock(time_t base, time_t step) {
  m_replayClock = true;
  m_clockBase = base;
  m_clockStep = step;
}

time_t HackRFSweepSource::Now() {
  if (!m_replayClock) return time(nullptr);
  return m_clockBase + time_t(m_clockCalls++) * m_clockStep;
}

double HackRFSweepSource::InterpolateSamples(uint8_t* ubuf, uint32_t validLength) {
  const uint32_t count = validLength / 2;
  uint64_t frequencyHz = 0;
  for (uint32_t i = 0; i < count; i += 8192) {
    if (ubuf[0] == 0x7F && ubuf[1] == 0x7F) {
      uint64_t thisFrequencyHz = 0;
      for (int k = 9; k >= 2; k--) thisFrequencyHz = (thisFrequencyHz << 8) | ubuf[k];
      if (frequencyHz != 0 && frequencyHz != thisFrequencyHz)
        printf("interpolateSamples: frequencyHz[%f] != thisFrequencyHz[%f]\n", double(frequencyHz),
               double(thisFrequencyHz));
      frequencyHz = thisFrequencyHz;
      int8_t post[2] = {int8_t(ubuf[10]), int8_t(ubuf[11])};
      if (i > 0) {
        post[0] = int8_t((post[0] + int8_t(ubuf[2 * (i - 1)])) / 2);
        post[1] = int8_t((post[1] + int8_t(ubuf[2 * (i - 1) + 1])) / 2);
      }
      for (uint32_t j = 0; j < 5; j++) {
        ubuf[2 * j] = uint8_t(post[0]);
        ubuf[2 * j + 1] = uint8_t(post[1]);
      }
    }
  }
  return double(frequencyHz + m_scanOffset);
}

int HackRFSweepSource::RxCallback(uint8_t* buffer, uint32_t validLength) {
  if (m_streamingState != Streaming) return 0;
  if (!GetIsDone()) {
    const double centerFrequency = InterpolateSamples(buffer, validLength);
    bool isScanStart = false;
    if (centerFrequency != m_centerFrequency) {
      GetNextFrequency();                          // "solely to decrement iteration count" (hackRFSource.cpp:234-236)
      isScanStart = GetIsScanStart();
      m_centerFrequency = centerFrequency;
    }
    const uint32_t count = validLength / 2;
    const time_t startTime = isScanStart ? Now() : 0;
    assert(count >= m_sampleCount);
    for (uint32_t i = 0; i < count; i += m_sampleCount)
      m_sampleQueue->AppendSamples(reinterpret_cast<int8_t(*)[2]>(&buffer[2 * i]), centerFrequency, startTime);
  } else {
    m_streamingState = Done;
  }
  return 0;
}

bool HackRFSweepSource::GetNextSamples(SampleQueue*, double_t& centerFrequency) {
  // hackRFSource.cpp:266-283: the sweep runs in firmware; this only waits for the stream to finish.
  centerFrequency = GetNextFrequency();
  m_streamingState = Streaming;
  while (m_streamingState != Done) std::this_thread::yield();
  return true;
}

bool HackRFSweepSource::Start() {
  if (m_streamingState != Streaming) m_streamingState = Streaming;   // hackRFSource.cpp:126-129
  return true;
}

bool HackRFSweepSource::StartStreaming(uint32_t numIterations, SampleQueue& sampleQueue) {
  Start();
  return StartThread(numIterations, sampleQueue);
}

void HackRFSweepSource::ThreadWorker() {
  // Stands in for libhackrf's rx thread + the reference's state-polling worker (hackRFSource.cpp:301-331):
  // deliver captured transfers until the callback reports Done or the capture runs out.
  for (size_t off = 0; m_validLength && off + m_validLength <= m_captureBytes && m_streamingState != Done;
       off += m_validLength)
    RxCallback(m_capture + off, m_validLength);
  m_streamingState = Done;
}

double HackRFSweepSource::Retune(double frequency) { return frequency; }
