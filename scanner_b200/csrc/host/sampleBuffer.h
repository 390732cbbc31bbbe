// SampleBuffer -- the reference's legacy (pre-SampleQueue) hand-off, sampleBuffer.h:8-46: a FIFO of
// sample buffers with their centre frequencies, appended by a source and drained by a consumer.
// Same AppendSamples overloads and SampleKind; as everywhere in this repo the buffer holds RAW
// samples and the consumer is a ProcessInterface visitor (the reference's GetNextSamples runs a
// CopyBufferProcessInterface over its CircularBuffer, sampleBuffer.cpp:127-152; pass a
// GpuProcessInterface here and the drained run goes straight to the fused kernel).
#pragma once
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <vector>

#include "buffer.h"
#include "sampleQueue.h"   // fftwf_complex

class SampleBuffer {
 public:
  enum SampleKind { Illegal = 0, Short, ShortComplex, FloatComplex } m_kind;   // sampleBuffer.h:34-39
  SampleBuffer(SampleKind kind, uint32_t enob, uint32_t count, uint32_t capacityBuffers = 16);
  void AppendSamples(int16_t* realSamples, int16_t* imagSamples, double centerFrequency);
  void AppendSamples(int16_t shortComplexSamples[][2], double centerFrequency);
  void AppendSamples(fftwf_complex* floatComplexSamples, double centerFrequency);
  // Visits up to maxBuffers queued buffers (blocks for the first): Begin(sequence id of the first
  // sample, total bytes), one Process per buffer, End.  Returns the number visited, 0 == done and
  // empty.  centerFrequencies receives one entry per visited buffer.
  uint32_t GetNextSamples(ProcessInterface<uint8_t>* process, std::vector<double>& centerFrequencies,
                          uint32_t maxBuffers = 1);
  // The reference's own drain (sampleBuffer.h:44, sampleBuffer.cpp:127-152): the next buffer as fftwf_complex.
  // fc32 buffers are copied; int16 kinds go through the converter installed with SetConverter (scn_convert_host
  // of a context of this kind: the reference's converter arithmetic on the GPU) -- without one the call fails
  // loudly, there is no CPU converter in this library.  false == done and empty.
  typedef SampleQueue::Converter Converter;
  void SetConverter(Converter convert) { m_convert = convert; }
  bool GetNextSamples(fftwf_complex* outputBuffer, double& centerFrequency);
  void SetIsDone();
  bool GetIsDone();
  // SCN_KIND_* of this buffer's samples, for scn_config.sample_kind
  uint32_t GetScnKind() const;
  size_t GetBufferBytes() const { return m_bufferBytes; }

 private:
  void Push(std::vector<uint8_t>&& raw, double centerFrequency);
  struct Item { uint64_t sequenceId; double frequency; std::vector<uint8_t> raw; };
  uint32_t m_sampleCount, m_enob, m_capacity;
  size_t m_bufferBytes;
  uint64_t m_nextSequenceId = 0;      // counts samples, like CircularBuffer's sequence ids
  std::deque<Item> m_queue;
  std::mutex m_mutex;
  std::condition_variable m_conditionEmpty, m_conditionFull;
  bool m_done = false;
  Converter m_convert;
};
