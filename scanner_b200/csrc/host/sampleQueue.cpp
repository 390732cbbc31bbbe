#include "sampleQueue.h"

#include <cassert>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scanner_b200.h"

// Copy of a run of raw buffers into the pinned slab.  The slab is written once by the producer and then read by the GPU's
// DMA engine, never by this CPU again, so large copies use non-temporal stores: no read-for-ownership of the destination
// lines and no pollution of the caches the producer's source data lives in (a third less DRAM traffic per sample on the
// host, which is what bounds the plugin surface once several producers run).
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void StreamCopyAvx2(char* dst, const char* src, size_t bytes) {
  size_t i = 0;
  for (; i + 128 <= bytes; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();                               // the DMA that follows must see the data
  if (i < bytes) memcpy(dst + i, src + i, bytes - i);
}
#endif
static void SlabCopy(void* dst, const void* src, size_t bytes) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2 && bytes >= (32u << 10) && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
    StreamCopyAvx2(static_cast<char*>(dst), static_cast<const char*>(src), bytes);
    return;
  }
#endif
  memcpy(dst, src, bytes);
}

static size_t BytesPerSample(SampleQueue::SampleKind kind) {
  switch (kind) {
    case SampleQueue::ByteComplex: return 2;
    case SampleQueue::Short: return 4;
    case SampleQueue::ShortComplex: return 4;
    case SampleQueue::FloatComplex: return 8;
    default: return 0;
  }
}

SampleQueue::SampleQueue(SampleKind kind, uint32_t enob, uint32_t sampleCount, uint32_t bufferCount,
                         bool correctDCOffset, bool doWrite)
    : m_kind(kind), m_enob(enob), m_sampleCount(sampleCount), m_bufferCount(bufferCount),
      m_correctDCOffset(correctDCOffset), m_doWrite(doWrite),
      m_bufferBytes(size_t(sampleCount) * BytesPerSample(kind)),
      m_writeCapacity(bufferCount / 10 ? bufferCount / 10 : 1) {      // messageQueue.h:149
  assert(kind > Illegal && kind <= FloatComplex);
  // like the reference's pool: 10 % more messages than queue slots (messageQueue.h:150)
  const uint32_t poolCount = uint32_t(bufferCount * 1.1) + 1;
  if (scn_alloc_pinned(m_bufferBytes * poolCount, &m_slab) != SCN_OK) {
    // no CUDA runtime / device: the queue itself still works from pageable memory; the GPU consumer
    // will fail loudly when it is created.
    m_slab = nullptr;
  }
  m_messages.resize(poolCount);
  for (uint32_t i = 0; i < poolCount; i++) {
    MessageType& m = m_messages[i];
    m.m_bytes = m_bufferBytes;
    m.m_data = m_slab ? static_cast<char*>(m_slab) + size_t(i) * m_bufferBytes : malloc(m_bufferBytes);
    m.m_header = MessageHeader{MessageHeader::Free, 0, 0.0, 0, 0};
    m_free.push_back(&m);
  }
  // Ring mode needs every message to come back in bounded time; the recording history parks messages for as long
  // as a window is open, so a recording queue keeps the free list.
  if (!doWrite) {
    m_ring = true;
    m_ringFree.assign(poolCount, 1);
    m_free.clear();
  }
  if (doWrite) {
    printf("Starting write thread...\n");                            // messageQueue.h:164-168
    m_writeThread.reset(new std::thread(&SampleQueue::WriteThreadWorker, this));
  }
}

SampleQueue::~SampleQueue() {
  if (m_writeThread) {
    printf("Stopping write thread...\n");                            // messageQueue.h:175-178
    {
      std::unique_lock<std::mutex> lock(m_writeMutex);
      m_writeShutdown = true;
      m_conditionWrite.notify_all();
    }
    m_writeThread->join();
  }
  if (m_slab) {
    scn_free_pinned(m_slab);
  } else {
    for (auto& m : m_messages) free(m.m_data);
  }
}

// messageQueue.h:67-72: every buffer before the SECOND scan-start marker is dropped.  Caller holds m_allocMutex
// (which also serialises several producers' view of the iteration count).
bool SampleQueue::AcceptOrDrop(time_t time) {
  if (time) m_iterationCount++;
  if (m_dropFirstSweep && m_iterationCount < 2) {
    m_dropped++;
    return false;
  }
  return true;
}

// Ring mode (SetFifoPool without recording): the slab is handed out strictly in address order, wrapping, and a
// message becomes available again when the ring head reaches it after it was freed.  Whatever order the consumers
// finish their batches in, consecutive appends therefore ALWAYS land in consecutive slab slots: a drained batch is
// one address run (two at the wrap) for ever -- a FIFO free list fragments a little more with every batch that two
// workers return out of order.  The price is head-of-line blocking on the oldest unfreed message, which the
// consumers' in-order batches make a non-issue.  Caller holds m_poolMutex; returns how many were taken (>= 1).
size_t SampleQueue::TakeFromRing(std::unique_lock<std::mutex>& lock, size_t want, std::vector<MessageType*>* out,
                                 std::deque<MessageType*>* cache) {
  m_poolAvailable.wait(lock, [this] { return m_ringFree[m_ringHead] != 0; });
  size_t got = 0;
  while (got < want && m_ringFree[m_ringHead]) {
    m_ringFree[m_ringHead] = 0;
    MessageType* m = &m_messages[m_ringHead];
    if (out) out->push_back(m); else cache->push_back(m);
    m_ringHead = (m_ringHead + 1 == m_messages.size()) ? 0 : m_ringHead + 1;
    got++;
  }
  return got;
}

SampleQueue::MessageType* SampleQueue::Allocate() {
  if (m_allocCache.empty() && m_ring) {                          // caller holds m_allocMutex
    std::unique_lock<std::mutex> lock(m_poolMutex);
    TakeFromRing(lock, kAllocChunk, nullptr, &m_allocCache);
  }
  if (m_allocCache.empty()) {
    std::unique_lock<std::mutex> lock(m_poolMutex);
    m_poolAvailable.wait(lock, [this] { return !m_free.empty(); });
    for (size_t i = 0; i < kAllocChunk && !m_free.empty(); i++) {
      if (m_fifoPool) {                      // oldest first: ascending slab addresses
        m_allocCache.push_back(m_free.front());
        m_free.pop_front();
      } else {                               // most recently freed first (cache-warm)
        m_allocCache.push_front(m_free.back());
        m_free.pop_back();
      }
    }
  }
  MessageType* m = m_allocCache.front();
  m_allocCache.pop_front();
  return m;
}

void SampleQueue::Free(MessageType* m) {
  std::unique_lock<std::mutex> lock(m_poolMutex);
  m->m_header.m_kind = MessageHeader::Free;
  if (m_ring) m_ringFree[size_t(m - m_messages.data())] = 1;
  else m_free.push_back(m);
  m_poolAvailable.notify_all();
}

// `count` free messages, oldest first in FIFO mode so that they form as few address runs as possible.
// Caller holds m_allocMutex.  Blocks while the pool is empty (the consumers return messages in batches).
void SampleQueue::AllocateMany(uint32_t count, std::vector<MessageType*>& out) {
  out.clear();
  while (out.size() < count) {
    while (!m_allocCache.empty() && out.size() < count) {
      out.push_back(m_allocCache.front());
      m_allocCache.pop_front();
    }
    if (out.size() == count) break;
    std::unique_lock<std::mutex> lock(m_poolMutex);
    if (m_ring) {
      TakeFromRing(lock, count - out.size(), &out, nullptr);
      continue;
    }
    m_poolAvailable.wait(lock, [this] { return !m_free.empty(); });
    while (!m_free.empty() && out.size() < count) {
      if (m_fifoPool) { out.push_back(m_free.front()); m_free.pop_front(); }
      else { out.push_back(m_free.back()); m_free.pop_back(); }
    }
  }
}

void SampleQueue::AppendSamplesBatch(const void* interleavedSamples, uint32_t count, const double* centerFrequencies,
                                     const time_t* times) {
  assert(m_kind != Short);                    // the split layout has no batched form (two pointers per buffer)
  const char* src = static_cast<const char*>(interleavedSamples);
  // pieces small enough that neither the pool nor the queue bound can deadlock on one batch
  const uint32_t piece = m_bufferCount / 4 ? m_bufferCount / 4 : 1;
  std::vector<MessageType*> msgs;
  std::vector<uint32_t> accepted;
  for (uint32_t first = 0; first < count; first += piece) {
    const uint32_t n = count - first < piece ? count - first : piece;
    accepted.clear();
    {
      std::unique_lock<std::mutex> cacheLock(m_allocMutex);
      for (uint32_t i = 0; i < n; i++)
        if (AcceptOrDrop(times ? times[first + i] : 0)) accepted.push_back(first + i);
      if (accepted.empty()) continue;
      AllocateMany(uint32_t(accepted.size()), msgs);
    }
    // copy outside every lock: one memcpy per run that is contiguous in BOTH the source and the slab
    for (size_t i = 0; i < accepted.size();) {
      size_t j = i + 1;
      while (j < accepted.size() && accepted[j] == accepted[j - 1] + 1 &&
             static_cast<char*>(msgs[j]->m_data) == static_cast<char*>(msgs[j - 1]->m_data) + m_bufferBytes) j++;
      SlabCopy(msgs[i]->m_data, src + size_t(accepted[i]) * m_bufferBytes, (j - i) * m_bufferBytes);
      i = j;
    }
    for (size_t i = 0; i < accepted.size(); i++) {
      MessageHeader& header = msgs[i]->m_header;
      header.m_time = times ? times[accepted[i]] : 0;
      header.m_frequency = centerFrequencies[accepted[i]];
      header.m_kind = MessageHeader::ProcessData;
      header.m_referenceCount = 0;
    }
    std::unique_lock<std::mutex> lock(m_mutex);
    m_conditionFull.wait(lock, [&] { return m_buffer.size() + accepted.size() <= m_bufferCount || m_buffer.empty(); });
    for (size_t i = 0; i < accepted.size(); i++) {
      msgs[i]->m_header.m_sequenceId = m_nextBufferSequenceId++;
      m_buffer.push_back(msgs[i]);
    }
    if (m_waiters && m_buffer.size() >= m_waitNeed) m_conditionEmpty.notify_all();
    ClearAck();
  }
}

void SampleQueue::SetProducerCount(uint32_t producers) {
  std::unique_lock<std::mutex> lock(m_mutex);
  m_producersLeft = producers ? producers : 1;
}

void SampleQueue::SynchronizedAppend(const void* a, size_t aBytes, const void* b, size_t bBytes,
                                     double centerFrequency, time_t time) {
  MessageType* message;
  {
    std::unique_lock<std::mutex> cacheLock(m_allocMutex);          // uncontended with a single producer
    if (!AcceptOrDrop(time)) return;
    message = Allocate();
  }
  memcpy(message->m_data, a, aBytes);
  if (bBytes) memcpy(static_cast<char*>(message->m_data) + aBytes, b, bBytes);
  MessageHeader& header = message->m_header;
  header.m_time = time;
  header.m_frequency = centerFrequency;
  header.m_kind = MessageHeader::ProcessData;
  header.m_referenceCount = 0;
  std::unique_lock<std::mutex> lock(m_mutex);
  header.m_sequenceId = m_nextBufferSequenceId++;
  m_conditionFull.wait(lock, [this] { return m_buffer.size() < m_bufferCount; });
  m_buffer.push_back(message);
  // wake consumers once what the least demanding waiter asked for is queued.  (Waking only on the empty ->
  // non-empty edge, as the reference's single-message consumers do, would strand a consumer that waits for a
  // group of K > 1 buffers: it wakes at 1, goes back to sleep, and nobody notifies again.)
  if (m_waiters && m_buffer.size() >= m_waitNeed) m_conditionEmpty.notify_all();
  ClearAck();
}

void SampleQueue::AppendSamples(int16_t* realSamples, int16_t* imagSamples, double centerFrequency, time_t time) {
  assert(m_kind == Short);
  const size_t half = size_t(m_sampleCount) * sizeof(int16_t);
  SynchronizedAppend(realSamples, half, imagSamples, half, centerFrequency, time);   // re block, then im block
}

void SampleQueue::AppendSamples(int16_t shortComplexSamples[][2], double centerFrequency, time_t time) {
  assert(m_kind == ShortComplex);
  SynchronizedAppend(shortComplexSamples, m_bufferBytes, nullptr, 0, centerFrequency, time);
}

void SampleQueue::AppendSamples(int8_t (*byteComplexSamples)[2], double centerFrequency, time_t time) {
  assert(m_kind == ByteComplex);
  SynchronizedAppend(byteComplexSamples, m_bufferBytes, nullptr, 0, centerFrequency, time);
}

void SampleQueue::AppendSamples(fftwf_complex* floatComplexSamples, double centerFrequency, time_t time) {
  assert(m_kind == FloatComplex);
  SynchronizedAppend(floatComplexSamples, m_bufferBytes, nullptr, 0, centerFrequency, time);
}

void SampleQueue::WaitForQueued(std::unique_lock<std::mutex>& lock, uint32_t need) {
  if (m_done || m_buffer.size() >= need) return;
  m_waiters++;
  if (m_waitNeed == 0 || need < m_waitNeed) m_waitNeed = need;
  m_conditionEmpty.wait(lock, [&] { return m_done || m_buffer.size() >= need; });
  if (--m_waiters == 0) m_waitNeed = 0;
}

SampleQueue::MessageType* SampleQueue::GetNextSamples() {
  std::unique_lock<std::mutex> lock(m_mutex);
  WaitForQueued(lock, 1);
  if (m_buffer.empty()) return nullptr;
  const bool wake = m_buffer.size() >= m_bufferCount;
  MessageType* message = m_buffer.front();
  m_buffer.pop_front();
  if (wake) m_conditionFull.notify_one();
  return message;
}

void SampleQueue::SetFifoPool(bool fifo) {
  // only meaningful for a recording queue (free-list mode); a ring-mode queue is always in address order
  std::unique_lock<std::mutex> cacheLock(m_allocMutex);
  std::unique_lock<std::mutex> lock(m_poolMutex);
  m_fifoPool = fifo;
}

uint32_t SampleQueue::GetNextBatch(std::vector<MessageType*>& out, uint32_t maxCount, uint32_t multiple,
                                   bool wait, bool contiguous, uint32_t minCount, uint32_t maxWaitMicros) {
  out.clear();
  if (multiple == 0) multiple = 1;
  std::unique_lock<std::mutex> lock(m_mutex);
  // a group of `multiple` buffers forms one averaged spectrum: wait for a whole group (or the end)
  if (wait) {
    WaitForQueued(lock, multiple);
    if (minCount > maxCount) minCount = maxCount;
    if (maxWaitMicros && !m_done && m_buffer.size() < minCount) {
      // linger: a deeper batch is worth a bounded wait (the producer notifies when `minCount` are queued)
      m_waiters++;
      if (m_waitNeed == 0 || minCount < m_waitNeed) m_waitNeed = minCount;
      m_conditionEmpty.wait_for(lock, std::chrono::microseconds(maxWaitMicros),
                                [&] { return m_done || m_buffer.size() >= minCount; });
      if (--m_waiters == 0) m_waitNeed = 0;
    }
  } else if (!(m_done || m_buffer.size() >= multiple)) return 0;
  size_t take = m_buffer.size() < maxCount ? m_buffer.size() : maxCount;
  if (contiguous) {
    size_t run = take ? 1 : 0;
    while (run < take && static_cast<char*>(m_buffer[run]->m_data) ==
                             static_cast<char*>(m_buffer[run - 1]->m_data) + m_bufferBytes) run++;
    // at least one whole group even across a break (the caller copies such a batch): never starve on fragmentation
    if (run >= multiple || run == take) take = run; else take = take < multiple ? take : multiple;
  }
  if (!(m_done && m_buffer.size() <= maxCount && take == m_buffer.size())) take -= take % multiple;
  if (take == 0) return 0;
  const bool wake = m_buffer.size() >= m_bufferCount;
  for (size_t i = 0; i < take; i++) {
    out.push_back(m_buffer.front());
    m_buffer.pop_front();
  }
  if (wake) m_conditionFull.notify_all();
  return uint32_t(take);
}

void SampleQueue::MessageProcessed(MessageType* message) {
  assert(message->m_header.m_kind != MessageHeader::Illegal);
  if (!m_doWrite) {            // no recording: nothing will ever read the history, the message returns to the pool
    Free(message);
    return;
  }
  // messageQueue.h:259-273: park the message in the write history, evicting the oldest
  std::unique_lock<std::mutex> lock(m_writeMutex);
  m_writeBuffer[message->m_header.m_sequenceId] = message;
  TrimWriteHistory();
  m_conditionWrite.notify_all();
}

void SampleQueue::TrimWriteHistory() {
  // the history holds bufferCount/10 messages (messageQueue.h:149,264-269) -- plus whatever an open window has
  // not written yet: the pool (and with it the producer) absorbs a writer that falls behind
  while (m_writeBuffer.size() > m_writeCapacity) {
    auto oldest = m_writeBuffer.begin();
    if (!m_writeJobs.empty() && oldest->first >= m_writeJobs.front().cursor) break;
    m_evictedBelow = oldest->first + 1;
    Free(oldest->second);
    m_writeBuffer.erase(oldest);
  }
}

void SampleQueue::MessageProcessed(const std::vector<MessageType*>& messages) {
  if (messages.empty()) return;
  if (!m_doWrite) {
    std::unique_lock<std::mutex> lock(m_poolMutex);
    for (MessageType* m : messages) {
      assert(m->m_header.m_kind != MessageHeader::Illegal);
      m->m_header.m_kind = MessageHeader::Free;
      if (m_ring) m_ringFree[size_t(m - m_messages.data())] = 1;
      else m_free.push_back(m);
    }
    m_poolAvailable.notify_all();
    return;
  }
  std::unique_lock<std::mutex> lock(m_writeMutex);
  for (MessageType* m : messages) m_writeBuffer[m->m_header.m_sequenceId] = m;
  TrimWriteHistory();
  m_conditionWrite.notify_all();
}

void SampleQueue::SetWriteConverter(Converter convert) {
  std::unique_lock<std::mutex> lock(m_writeMutex);
  m_convert = convert;
}

void SampleQueue::BeginWrite(uint64_t startSequenceId, std::string fileName, uint64_t limit) {
  printf("BeginWrite %s: %lu\n", fileName.c_str(), (unsigned long)startSequenceId);   // messageQueue.h:275-282
  std::unique_lock<std::mutex> lock(m_writeMutex);
  m_writeStartSequenceId = startSequenceId;
  m_writeEndSequenceId = UINT64_MAX;
  if (!m_doWrite) return;
  FILE* f = fopen(fileName.c_str(), "w");
  if (!f) { perror(fileName.c_str()); exit(1); }
  m_writeJobs.push_back(WriteJob{startSequenceId, UINT64_MAX, limit, f});
  m_conditionWrite.notify_all();
}

void SampleQueue::LimitWrite(uint64_t limit) {
  std::unique_lock<std::mutex> lock(m_writeMutex);
  if (m_writeJobs.empty() || m_writeJobs.back().end != UINT64_MAX) return;      // no open window
  if (limit > m_writeJobs.back().limit) m_writeJobs.back().limit = limit;
  m_conditionWrite.notify_all();
}

void SampleQueue::EndWrite(uint64_t sequenceId) {
  printf("EndWrite %lu\n", (unsigned long)sequenceId);                                  // messageQueue.h:284-288
  std::unique_lock<std::mutex> lock(m_writeMutex);
  m_writeEndSequenceId = sequenceId;
  if (!m_writeJobs.empty()) m_writeJobs.back().end = sequenceId;
  m_conditionWrite.notify_all();
}

void SampleQueue::WriteThreadWorker() {
  std::unique_lock<std::mutex> lock(m_writeMutex);
  while (true) {
    m_conditionWrite.wait(lock, [this] { return !m_writeJobs.empty() || m_writeShutdown; });
    if (m_writeJobs.empty()) break;
    WriteJob& job = m_writeJobs.front();                  // (push_back does not invalidate a deque reference)
    uint64_t& seq = job.cursor;
    while (true) {
      // next message of the window: the exact id once it has been processed; an id the history no longer holds
      // is skipped; at shutdown whatever is still parked is flushed
      // while the window is open nothing at or past the caller's limit is written (it may turn out to lie outside)
      auto bound = [&] { return job.end != UINT64_MAX ? job.end : job.limit; };
      m_conditionWrite.wait(lock, [&] {
        if (seq < m_evictedBelow) seq = m_evictedBelow;
        return seq >= job.end || (seq < bound() && m_writeBuffer.count(seq) != 0) || m_writeShutdown;
      });
      if (seq >= bound()) break;                          // closed and complete (or shut down at the limit)
      auto it = m_writeBuffer.find(seq);
      if (it == m_writeBuffer.end()) {                    // shutting down: jump to the next parked id, if any
        it = m_writeBuffer.lower_bound(seq);
        if (it == m_writeBuffer.end() || it->first >= bound()) break;
        seq = it->first;
      }
      MessageType* message = it->second;
      printf("Writing %lu\n", (unsigned long)seq);                                       // messageQueue.h:125
      const void* data = message->m_data;
      if (m_kind != FloatComplex) {
        if (!m_convert) { fprintf(stderr, "SampleQueue: recording needs a converter (SetWriteConverter)\n"); exit(1); }
        m_writeScratch.resize(size_t(m_sampleCount) * 2);
        if (!m_convert(message->m_data, 1, m_writeScratch.data())) {
          fprintf(stderr, "SampleQueue: sample conversion failed: %s\n", scn_last_error());
          exit(1);
        }
        data = m_writeScratch.data();
      }
      fwrite(data, sizeof(fftwf_complex), m_sampleCount, job.file);                       // messageQueue.h:127-130
      m_written++;
      seq++;
      TrimWriteHistory();
    }
    fclose(job.file);
    m_writeJobs.pop_front();
    TrimWriteHistory();
  }
}

void SampleQueue::SetIsDone() {
  std::unique_lock<std::mutex> lock(m_mutex);
  if (m_producersLeft > 1) {       // other producer threads are still appending
    m_producersLeft--;
    return;
  }
  m_producersLeft = 0;
  m_done = true;
  m_conditionEmpty.notify_all();
}

bool SampleQueue::GetIsDone() {
  std::unique_lock<std::mutex> lock(m_mutex);
  return m_done;
}

bool SampleQueue::ReceivedAck() { return m_acknowledged; }

void SampleQueue::SendAck() {
  bool expected = false;
  m_acknowledged.compare_exchange_strong(expected, true);
}

void SampleQueue::ClearAck() {
  bool expected = true;
  m_acknowledged.compare_exchange_strong(expected, false);
}
