// SampleQueue -- the live hand-off between a SignalSource (producer thread) and ProcessSamples
// (consumer threads).  Same public surface as the reference's SampleQueue = MessageQueue<fftwf_complex>
// (messageQueue.h:11-328): the four AppendSamples overloads, GetNextSamples / MessageProcessed,
// SetIsDone / GetIsDone, SendAck / ReceivedAck / ClearAck, MessageHeader and SampleKind.
//
// What is different, by design: AppendSamples stores the RAW device samples (2/4/8 bytes per
// sample) in pinned host memory and does NOT convert them -- the reference converts to fp32 on the
// producer thread (messageQueue.h:196,210,223); here conversion is fused into the GPU kernel, so
// a message is 4x (int8) / 2x (int16) smaller and is DMA-able as it lies.  Semantics kept:
//   * caller keeps ownership of its buffer, the queue copies (messageQueue.h:74-75);
//   * everything before the SECOND scan-start marker (time != 0) is dropped (messageQueue.h:67-72)
//     unless SetDropFirstSweep(false);
//   * sequence ids count accepted buffers; FIFO delivery; bounded: Append blocks when full.
// GetNextBatch() is the batched form of GetNextSamples() the GPU consumer uses.
// Triggered recording (messageQueue.h:98-139, 259-288): with doWrite the queue parks processed messages in a
// history of bufferCount/10 and a writer thread writes the window [BeginWrite start, EndWrite id) to the named
// file as fftwf_complex, exactly the bytes the reference writes.  The messages hold RAW samples here, so the
// writer converts through the callback ProcessSamples installs (scn_convert_host: the reference's converters
// on the GPU); there is no CPU converter in this library.  Unlike the reference's writer -- which can miss a
// BeginWrite wake-up, stops at SetIsDone with messages unwritten and waits forever for an evicted start --
// this one writes strictly in sequence order whatever order workers finish in, keeps what an open window
// still needs out of the history's eviction, skips what had already left it, and flushes at shutdown; where the reference's races do not bite, the files are identical.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <ctime>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <thread>
#include <mutex>
#include <string>
#include <vector>

typedef float fftwf_complex[2];   // the only thing the plugin surface used from <fftw3.h>

class SampleQueue {
 public:
  struct MessageHeader {
    enum MessageKind { Illegal = 0, ProcessData, WriteData, WriteDataAndStop, Free } m_kind;
    uint32_t m_referenceCount;
    double m_frequency;
    uint64_t m_sequenceId;
    time_t m_time;
  };
  enum SampleKind { Illegal = 0, ByteComplex, Short, ShortComplex, FloatComplex };

  // One pooled message: header + raw bytes (pinned).
  class MessageType {
   public:
    MessageHeader m_header;
    MessageHeader& GetHeader() { return m_header; }
    void* GetData() { return m_data; }           // raw samples of the queue's SampleKind
    size_t GetDataBytes() const { return m_bytes; }
   private:
    friend class SampleQueue;
    void* m_data = nullptr;
    size_t m_bytes = 0;
  };

  SampleKind m_kind;

  SampleQueue(SampleKind kind, uint32_t enob, uint32_t sampleCount, uint32_t bufferCount,
              bool correctDCOffset, bool doWrite);
  ~SampleQueue();
  SampleQueue(const SampleQueue&) = delete;
  SampleQueue& operator=(const SampleQueue&) = delete;

  void AppendSamples(int16_t* realSamples, int16_t* imagSamples, double centerFrequency, time_t time);
  void AppendSamples(int16_t shortComplexSamples[][2], double centerFrequency, time_t time);
  void AppendSamples(int8_t (*byteComplexSamples)[2], double centerFrequency, time_t time);
  void AppendSamples(fftwf_complex* floatComplexSamples, double centerFrequency, time_t time);
  // Batched form of the three interleaved overloads above -- `count` consecutive buffers of the queue's kind, e.g.
  // one 262 144-byte HackRF transfer = 64 buffers of 2048 int8 IQ samples (hackRFSource.cpp:251-264).  Exactly
  // `count` AppendSamples calls in order (drop rule, sequence ids, FIFO), but ONE pool transaction, one memcpy per
  // contiguous slab run and one queue lock.  times == nullptr: all zero.  Safe to call from several producer
  // threads (SetProducerCount); a batch's buffers get consecutive sequence ids.
  void AppendSamplesBatch(const void* interleavedSamples, uint32_t count, const double* centerFrequencies,
                          const time_t* times);
  // Number of producer threads that will each call SetIsDone() when they finish (default 1, the reference's
  // single source thread): the queue is done when the last one has.
  void SetProducerCount(uint32_t producers);

  MessageType* GetNextSamples();                                   // nullptr == done and drained
  // Blocks for the first message, then takes what is queued: up to maxCount, a multiple of `multiple`
  // unless the queue is done.  Returns the number taken (0 == done and drained).
  // wait == false: never blocks -- returns 0 when no whole group is queued yet (used by the consumer to
  // finish an in-flight batch instead of sleeping on the queue).
  // contiguous == true: the batch stops where the next message does not follow the previous one in memory, so the
  // whole batch is ONE address run of the pinned slab and can be handed to scn_submit without a staging copy.
  // minCount / maxWaitMicros (wait == true only): once a first group is queued, linger up to maxWaitMicros for at
  // least minCount messages -- a GPU consumer wants launches of thousands of buffers, not of whatever happened to
  // be queued when it looked; the linger bounds the latency this adds when the source is slow.
  uint32_t GetNextBatch(std::vector<MessageType*>& out, uint32_t maxCount, uint32_t multiple = 1,
                        bool wait = true, bool contiguous = false, uint32_t minCount = 0,
                        uint32_t maxWaitMicros = 0);
  // A queue that does not record hands its slab out in address order (a ring, see TakeFromRing in the .cpp), so
  // consecutive appends land in consecutive slab slots and a drained batch is one address run.  A recording queue
  // (doWrite) parks messages in its history, so it keeps a free list; SetFifoPool(true) makes that list first-in
  // first-out (ascending addresses while nothing is parked) instead of most-recently-freed first.
  void SetFifoPool(bool fifo);
  bool IsPinnedSlab() const { return m_slab != nullptr; }
  void MessageProcessed(MessageType* message);
  // the same for a whole drained batch under ONE pool / history lock (a consumer that returns 1024 messages one by
  // one fights the producer's Allocate for the pool mutex 1024 times)
  void MessageProcessed(const std::vector<MessageType*>& messages);

  // `limit`: no message with an id >= limit is written while the window is still open (the caller's CURRENT idea of
  // where it will end; LimitWrite raises it when a later trigger extends the window, EndWrite fixes the end).  Without
  // it the writer thread would race the caller: a message past the end that is already in the history when EndWrite
  // arrives might or might not have been written.  Default: no limit (the reference's interface, messageQueue.h:275).
  void BeginWrite(uint64_t startSequenceId, std::string fileName, uint64_t limit = UINT64_MAX);
  void LimitWrite(uint64_t limit);
  void EndWrite(uint64_t sequenceId);
  // raw buffers of this queue's kind -> interleaved float re/im; false == failure (the writer then exits(1))
  typedef std::function<bool(const void* raw, uint32_t nBuffers, float* out)> Converter;
  void SetWriteConverter(Converter convert);
  uint64_t GetWrittenCount() const { return m_written; }
  void SetIsDone();
  bool GetIsDone();
  bool ReceivedAck();
  void SendAck();
  void ClearAck();

  void SetDropFirstSweep(bool drop) { m_dropFirstSweep = drop; }
  uint32_t GetEnob() const { return m_enob; }
  uint32_t GetSampleCount() const { return m_sampleCount; }
  bool GetCorrectDCOffset() const { return m_correctDCOffset; }
  size_t GetBufferBytes() const { return m_bufferBytes; }
  uint64_t GetAcceptedCount() const { return m_nextBufferSequenceId; }
  uint64_t GetDroppedCount() const { return m_dropped; }

 private:
  void SynchronizedAppend(const void* a, size_t aBytes, const void* b, size_t bBytes,
                          double centerFrequency, time_t time);
  MessageType* Allocate();
  void AllocateMany(uint32_t count, std::vector<MessageType*>& out);
  bool AcceptOrDrop(time_t time);              // the drop rule of messageQueue.h:67-72; m_allocMutex held
  void Free(MessageType* m);
  void WaitForQueued(std::unique_lock<std::mutex>& lock, uint32_t need);   // m_mutex held
  uint32_t m_waiters = 0, m_waitNeed = 0;      // consumers asleep on m_conditionEmpty / the smallest count one waits for

  uint32_t m_enob;
  uint32_t m_sampleCount;
  uint32_t m_bufferCount;
  bool m_correctDCOffset;
  bool m_doWrite;
  bool m_dropFirstSweep = true;
  size_t m_bufferBytes;
  uint32_t m_iterationCount = 0;
  uint64_t m_nextBufferSequenceId = 0;
  uint64_t m_dropped = 0;
  bool m_done = false;
  uint32_t m_producersLeft = 1;
  std::atomic<bool> m_acknowledged{true};
  uint64_t m_writeStartSequenceId = 0, m_writeEndSequenceId = 0;

  // recording
  void WriteThreadWorker();
  std::mutex m_writeMutex;
  std::condition_variable m_conditionWrite;
  std::map<uint64_t, MessageType*> m_writeBuffer;      // processed messages by sequence id (history)
  uint32_t m_writeCapacity;
  uint64_t m_evictedBelow = 0;                          // every id below this has left the history
  std::unique_ptr<std::thread> m_writeThread;
  struct WriteJob { uint64_t cursor, end, limit; FILE* file; };  // one per BeginWrite, written in order, each to completion
  void TrimWriteHistory();                               // caller holds m_writeMutex
  std::deque<WriteJob> m_writeJobs;
  bool m_writeShutdown = false;
  Converter m_convert;
  std::vector<float> m_writeScratch;
  uint64_t m_written = 0;

  std::mutex m_mutex;
  std::condition_variable m_conditionEmpty, m_conditionFull;
  std::deque<MessageType*> m_buffer;          // FIFO of filled messages
  // pool
  std::mutex m_poolMutex;
  std::condition_variable m_poolAvailable;
  std::vector<MessageType> m_messages;
  std::deque<MessageType*> m_free;
  bool m_fifoPool = false;
  // ring mode (see TakeFromRing): which messages are free, and the next slab slot to hand out
  bool m_ring = false;
  std::vector<uint8_t> m_ringFree;
  size_t m_ringHead = 0;
  size_t TakeFromRing(std::unique_lock<std::mutex>& lock, size_t want, std::vector<MessageType*>* out,
                      std::deque<MessageType*>* cache);
  // free messages the appending side has already taken out of the pool (refilled kAllocChunk at a time, so the
  // producer touches the contended pool mutex once per chunk instead of once per buffer)
  static const size_t kAllocChunk = 32;
  std::mutex m_allocMutex;
  std::deque<MessageType*> m_allocCache;
  void* m_slab = nullptr;                     // one pinned allocation backing every message
};
