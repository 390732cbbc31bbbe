#include "sampleBuffer.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scanner_b200.h"

SampleBuffer::SampleBuffer(SampleKind kind, uint32_t enob, uint32_t count, uint32_t capacityBuffers)
    : m_kind(kind), m_sampleCount(count), m_enob(enob), m_capacity(capacityBuffers ? capacityBuffers : 1),
      m_bufferBytes(size_t(count) * (kind == FloatComplex ? 8 : 4)) {
  assert(kind > Illegal && kind <= FloatComplex);
}

uint32_t SampleBuffer::GetScnKind() const {
  switch (m_kind) {
    case Short: return SCN_KIND_SHORT;
    case ShortComplex: return SCN_KIND_SHORT_COMPLEX;
    default: return SCN_KIND_FLOAT_COMPLEX;
  }
}

void SampleBuffer::Push(std::vector<uint8_t>&& raw, double centerFrequency) {
  std::unique_lock<std::mutex> lock(m_mutex);
  m_conditionFull.wait(lock, [this] { return m_queue.size() < m_capacity; });
  const bool wake = m_queue.empty();
  m_queue.push_back(Item{m_nextSequenceId, centerFrequency, std::move(raw)});
  m_nextSequenceId += m_sampleCount;
  if (wake) m_conditionEmpty.notify_one();
}

void SampleBuffer::AppendSamples(int16_t* realSamples, int16_t* imagSamples, double centerFrequency) {
  assert(m_kind == Short);
  std::vector<uint8_t> raw(m_bufferBytes);
  memcpy(raw.data(), realSamples, m_bufferBytes / 2);
  memcpy(raw.data() + m_bufferBytes / 2, imagSamples, m_bufferBytes / 2);
  Push(std::move(raw), centerFrequency);
}

void SampleBuffer::AppendSamples(int16_t shortComplexSamples[][2], double centerFrequency) {
  assert(m_kind == ShortComplex);
  std::vector<uint8_t> raw(m_bufferBytes);
  memcpy(raw.data(), shortComplexSamples, m_bufferBytes);
  Push(std::move(raw), centerFrequency);
}

void SampleBuffer::AppendSamples(fftwf_complex* floatComplexSamples, double centerFrequency) {
  assert(m_kind == FloatComplex);
  std::vector<uint8_t> raw(m_bufferBytes);
  memcpy(raw.data(), floatComplexSamples, m_bufferBytes);
  Push(std::move(raw), centerFrequency);
}

uint32_t SampleBuffer::GetNextSamples(ProcessInterface<uint8_t>* process, std::vector<double>& centerFrequencies,
                                      uint32_t maxBuffers) {
  centerFrequencies.clear();
  std::vector<Item> taken;
  {
    std::unique_lock<std::mutex> lock(m_mutex);
    m_conditionEmpty.wait(lock, [this] { return m_done || !m_queue.empty(); });
    if (m_queue.empty()) return 0;
    const bool wake = m_queue.size() >= m_capacity;
    while (!m_queue.empty() && taken.size() < maxBuffers) {
      taken.push_back(std::move(m_queue.front()));
      m_queue.pop_front();
    }
    if (wake) m_conditionFull.notify_all();
  }
  process->Begin(taken.front().sequenceId, uint32_t(taken.size() * m_bufferBytes));
  for (auto& it : taken) {
    process->Process(it.raw.data(), uint32_t(it.raw.size()));
    centerFrequencies.push_back(it.frequency);
  }
  process->End();
  return uint32_t(taken.size());
}

bool SampleBuffer::GetNextSamples(fftwf_complex* outputBuffer, double& centerFrequency) {
  Item item;
  {
    std::unique_lock<std::mutex> lock(m_mutex);
    m_conditionEmpty.wait(lock, [this] { return m_done || !m_queue.empty(); });
    if (m_queue.empty()) return false;
    const bool wake = m_queue.size() >= m_capacity;
    item = std::move(m_queue.front());
    m_queue.pop_front();
    if (wake) m_conditionFull.notify_all();
  }
  centerFrequency = item.frequency;
  if (m_kind == FloatComplex) {
    memcpy(outputBuffer, item.raw.data(), m_bufferBytes);
  } else {
    if (!m_convert) { fprintf(stderr, "SampleBuffer: int16 samples need a converter (SetConverter)\n"); exit(1); }
    if (!m_convert(item.raw.data(), 1, &outputBuffer[0][0])) {
      fprintf(stderr, "SampleBuffer: sample conversion failed: %s\n", scn_last_error());
      exit(1);
    }
  }
  return true;
}

void SampleBuffer::SetIsDone() {
  std::unique_lock<std::mutex> lock(m_mutex);
  m_done = true;
  m_conditionEmpty.notify_all();
}

bool SampleBuffer::GetIsDone() {
  std::unique_lock<std::mutex> lock(m_mutex);
  return m_done;
}
