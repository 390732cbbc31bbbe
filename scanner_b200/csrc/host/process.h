// ProcessSamples -- the consumer side of the plugin surface.  Same constructor arguments, Mode enum
// and entry points as the reference's ProcessSamples (process.h:23-91): StartProcessing(SampleQueue&)
// spawns `threadCount` worker threads and joins them; Run() is the synchronous single-buffer call.
//
// Each worker owns one scn_ctx (the C ABI is thread-compatible) and replaces the reference's
// per-message body (process.cpp:292-299: memcpy -> window -> FFT -> process_fft) with: drain up to
// `maxBatch` raw messages from the queue, ONE fused GPU launch for the whole batch
// (scn_submit/scn_collect), then, per message in sequence order, the reference's own bookkeeping
// (process.cpp:280-287 "Start scan at", :57 "freq %lu power_db %f", :303-309 ack / ProcessWrite /
// MessageProcessed).  With threadCount == 1 the output is line for line what the reference prints.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "sampleQueue.h"
#include "scanner_b200.h"

class SignalSource;

class ProcessSamples {
 public:
  enum Mode { Illegal, TimeDomain, FrequencyDomain };

  // One detection handed to the optional sink, in print order.
  struct Detection {
    uint64_t sequenceId;
    double centerFrequency;
    uint64_t frequencyHz;     // process.cpp:55-57
    float powerDb;
    uint32_t bin;
  };

  ProcessSamples(uint32_t numSamples, uint32_t sampleRate, uint32_t enob, float threshold,
                 int windowType /* SCN_WIN_*, == gr::fft::window::win_type */, Mode mode,
                 uint32_t threadCount = 1, std::string fileNameBase = "", double useBandWidth = 0.75,
                 double dcIgnoreWidth = 0.0, uint32_t preTrigger = 2, uint32_t postTrigger = 4);
  ~ProcessSamples();

  void Run(int16_t sample_buffer[][2], uint32_t centerFrequency);
  bool StartProcessing(SampleQueue& sampleQueue);

  // GPU-path options (defaults reproduce the reference's behaviour)
  void SetDevice(int device) { m_device = device; }
  void SetAveraging(uint32_t k) { m_averaging = k ? k : 1; }
  void SetMaxBatch(uint32_t maxBatch) { m_maxBatch = maxBatch ? maxBatch : 1; }
  // Hand batches to the GPU straight from the queue's pinned slab (scn_submit_gather: one H2D copy per address run, no
  // host copy) whenever a drained batch is made of long runs; the queue then recycles its messages FIFO so that this is
  // the common case.  On by default; off == always pack the batch into a staging buffer first.
  void SetZeroCopy(bool on) { m_zeroCopy = on; }
  // Batching policy.  Default (0, 0): take whatever is queued and launch at once -- results are never held back.
  // (minBatch, micros): once something is queued, linger up to `micros` for `minBatch` buffers, so a fast source gets
  // launches of thousands of buffers instead of dozens (per-launch CPU cost is what bounds a GPU consumer).
  void SetBatchLinger(uint32_t minBatch, uint32_t micros) { m_minBatch = minBatch; m_lingerMicros = micros; }
  uint64_t GetZeroCopyBatches() const { return m_zeroCopyBatches; }
  void SetOutput(FILE* out) { m_out = out; }                               // nullptr silences printing
  void SetDetectionSink(std::function<void(const Detection&)> sink) { m_sink = std::move(sink); }
  uint64_t GetBuffersProcessed() const { return m_buffersProcessed; }
  uint64_t GetHitCount() const { return m_hitCount; }
  uint64_t GetLaunchCount() const { return m_launches; }
  bool m_writeData = false;

 private:
  static const uint32_t MAX_THREADS = 8;
  void ThreadWorker(uint32_t threadId);
  scn_ctx* CreateContext(SampleQueue::SampleKind kind, uint32_t enob, bool correctDC, uint32_t maxSpectra,
                         uint32_t hitCap);
  void TimeToString(time_t t, char* buffer, uint32_t length);
  void ProcessWrite(bool doWrite, double centerFrequency, uint64_t sequenceId);

  uint32_t m_sampleCount, m_sampleRate, m_enob;
  float m_threshold;
  int m_windowType;
  Mode m_mode;
  uint32_t m_threadCount;
  std::string m_fileNameBase;
  uint32_t m_useWindow, m_dcIgnoreWindow;
  uint32_t m_preTrigger, m_postTrigger;
  SampleQueue* m_sampleQueue = nullptr;
  std::thread* m_threads[MAX_THREADS] = {};
  std::vector<float> m_window;
  int m_device = 0;
  uint32_t m_averaging = 1;
  uint32_t m_maxBatch = 1024;
  FILE* m_out = stdout;
  std::function<void(const Detection&)> m_sink;
  std::atomic<uint64_t> m_buffersProcessed{0}, m_hitCount{0}, m_launches{0};
  std::atomic<bool> m_writing{false};
  std::atomic<uint64_t> m_endSequenceId{0};
  scn_ctx* m_runCtx = nullptr;                     // Run()'s context, created on first use
  bool m_zeroCopy = true;
  uint32_t m_minBatch = 0, m_lingerMicros = 0;
  std::atomic<uint64_t> m_zeroCopyBatches{0};
  uint32_t m_fileCounter = 0;                      // process.cpp:169
  std::shared_ptr<scn_ctx> m_writeCtx;             // converts recorded messages (scn_convert_host), writer thread only;
                                                   // shared with the queue's converter so either may be destroyed first
  std::function<time_t()> m_clock;                 // file-name stamps; default time(NULL)
 public:
  void SetClock(std::function<time_t()> clock) { m_clock = std::move(clock); }   // reproducible replays
};
