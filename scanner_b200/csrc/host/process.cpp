#include "process.h"

#include <cassert>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {
std::mutex g_printMutex;   // whole-buffer output stays contiguous when several workers print

[[noreturn]] void Die(const char* what) {
  // the reference's error convention in this layer: message to stderr, exit(1)
  fprintf(stderr, "%s: %s\n", what, scn_last_error());
  exit(1);
}
}  // namespace

ProcessSamples::ProcessSamples(uint32_t numSamples, uint32_t sampleRate, uint32_t enob, float threshold,
                               int windowType, Mode mode, uint32_t threadCount, std::string fileNameBase,
                               double useBandWidth, double dcIgnoreWidth, uint32_t preTrigger,
                               uint32_t postTrigger)
    : m_sampleCount(numSamples), m_sampleRate(sampleRate), m_enob(enob), m_threshold(threshold),
      m_windowType(windowType), m_mode(mode), m_threadCount(threadCount), m_fileNameBase(fileNameBase),
      m_useWindow(scn_use_window(useBandWidth, numSamples)),   // process.cpp:85
      m_dcIgnoreWindow(4),                                     // process.cpp:87: hard-coded, dcIgnoreWidth unused
      m_preTrigger(preTrigger), m_postTrigger(postTrigger), m_window(numSamples) {
  (void)dcIgnoreWidth;
  assert(mode > Illegal && mode <= FrequencyDomain);
  assert(threadCount >= 1 && threadCount <= MAX_THREADS);
  if (scn_window_build(windowType, numSamples, m_window.data()) != SCN_OK) Die("FFTWindow");
}

ProcessSamples::~ProcessSamples() {
  if (m_runCtx) scn_destroy(m_runCtx);
}

// Hit records per spectrum copied back with every batch; a spectrum with more hits than this is re-run
// alone through a full-capacity context (rare: the reference itself treats > 1047 hits as an event,
// process.cpp:62), so every hit is still reported, in order.
static const uint32_t kHitCap = 64;

scn_ctx* ProcessSamples::CreateContext(SampleQueue::SampleKind kind, uint32_t enob, bool correctDC,
                                       uint32_t maxSpectra, uint32_t hitCap) {
  scn_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.device = m_device;
  cfg.sample_count = m_sampleCount;
  cfg.sample_rate = m_sampleRate;
  cfg.enob = enob;
  cfg.sample_kind = uint32_t(kind);               // same numbering (messageQueue.h:31-37)
  cfg.correct_dc_offset = correctDC ? 1 : 0;
  cfg.averaging = (m_mode == FrequencyDomain) ? m_averaging : 1;
  cfg.mode = (m_mode == TimeDomain) ? SCN_MODE_TIME_DOMAIN : SCN_MODE_FREQUENCY_DOMAIN;
  cfg.threshold = m_threshold;
  cfg.use_window = m_useWindow;
  cfg.dc_ignore_window = m_dcIgnoreWindow;
  cfg.window = m_window.data();
  cfg.max_spectra = maxSpectra;
  cfg.max_hits_per_spectrum = hitCap;             // 0 == N
  cfg.flags = SCN_OUT_HITS;
  cfg.ticket_slots = 2;
  scn_ctx* ctx = nullptr;
  if (scn_create(&cfg, &ctx) != SCN_OK) Die("scn_create");
  return ctx;
}

void ProcessSamples::TimeToString(time_t t, char* buffer, uint32_t length) {
  struct tm tmv;
  if (localtime_r(&t, &tmv) == nullptr) { perror("localtime"); exit(1); }
  if (strftime(buffer, length, "%Y%m%d-%T", &tmv) == 0) { fprintf(stderr, "strftime returned 0"); exit(1); }
}

void ProcessSamples::ProcessWrite(bool doWrite, double centerFrequency, uint64_t sequenceId) {
  // Trigger/record bookkeeping of process.cpp:250-270; the queue's writer thread records the window.
  if (m_writing) {
    if (doWrite) {
      uint64_t end = sequenceId + m_postTrigger + 1, cur = m_endSequenceId;
      while (cur < end && !m_endSequenceId.compare_exchange_weak(cur, end)) {}
      m_sampleQueue->LimitWrite(m_endSequenceId);    // the writer thread may go on up to the extended end
    } else if (sequenceId == m_endSequenceId || (m_averaging > 1 && sequenceId > m_endSequenceId)) {
      // K == 1: the reference's strict equality (process.cpp:262).  With K-FFT averaging only the first id of every
      // K-group arrives here (0, K, 2K, ...), so the end of the window may never be hit exactly: the first group
      // at or past it closes the window AT the end id.
      m_sampleQueue->EndWrite(m_endSequenceId);
      m_writing = false;
    }
  } else if (doWrite && !m_fileNameBase.empty()) {
    char name[320], tbuf[64];                       // process.cpp:160-181: base + time + "-<frequency>-<counter>"
    TimeToString(m_clock ? m_clock() : time(nullptr), tbuf, sizeof(tbuf));
    snprintf(name, sizeof(name), "%s%s-%.0f-%u", m_fileNameBase.c_str(), tbuf, centerFrequency, ++m_fileCounter);
    const uint64_t dec = sequenceId < m_preTrigger ? sequenceId : m_preTrigger;
    m_endSequenceId = sequenceId + m_postTrigger + 1;
    m_sampleQueue->BeginWrite(sequenceId - dec, name, m_endSequenceId);   // ... and never past the planned end
    m_writing = true;
  }
}

void ProcessSamples::ThreadWorker(uint32_t threadId) {
  (void)threadId;
  SampleQueue* q = m_sampleQueue;
  const uint32_t K = (m_mode == FrequencyDomain) ? m_averaging : 1;
  const uint32_t maxSpectra = (m_maxBatch + K - 1) / K;
  const uint32_t cap = (m_mode == FrequencyDomain && kHitCap < m_sampleCount) ? kHitCap : m_sampleCount;
  scn_ctx* ctx = CreateContext(q->m_kind, q->GetEnob(), q->GetCorrectDCOffset(), maxSpectra, cap);
  scn_ctx* fullCtx = nullptr;                            // lazily created, one spectrum, cap == N
  std::vector<scn_hit> fullHits;
  const size_t bufBytes = q->GetBufferBytes();
  const uint32_t N = m_sampleCount;

  // Two batches in flight: while batch i is on the GPU (H2D -> fused kernel -> D2H on its ticket's
  // stream), batch i+1 is drained from the queue and staged.  Results are never held back waiting for
  // more input: with nothing queued the in-flight batch is collected at once.
  struct InFlight {
    void* staging = nullptr;                       // contiguous pinned batch (only used when the slab runs are too short)
    std::vector<const void*> runs;                 // what is submitted: address runs of the queue's pinned slab ...
    std::vector<uint32_t> runBuffers;              // ... and their lengths in buffers (or the staging copy as one run)
    std::vector<SampleQueue::MessageType*> batch;
    uint32_t ticket = 0, nSpectra = 0;
    bool active = false;
  } slot[2];
  for (auto& s : slot)
    if (scn_alloc_pinned(bufBytes * size_t(maxSpectra) * K, &s.staging) != SCN_OK) Die("scn_alloc_pinned");
  // results are read in place from the ticket's pinned buffers (scn_collect_view): no copy of counts / hit records
  const uint32_t* counts = nullptr;
  const scn_hit* hits = nullptr;
  const float* tdmm = nullptr;

  const bool zeroCopy = m_zeroCopy && q->IsPinnedSlab();
  // What a batch is submitted from.  The messages live in ONE pinned slab and the pool recycles first-in first-out,
  // so a drained batch is a handful of address runs (one per producer append, or one in all with a single
  // producer): those go to scn_submit_gather as they lie -- one H2D copy per run, no host copy of any sample.
  // Only when the runs are short (single-buffer appends racing a LIFO pool, or no pinned slab) is the batch packed
  // into `staging` first, as the reference's consumer copies every message (process.cpp:293-295).
  std::vector<char> overflowRaw;
  auto stage = [&](InFlight& f) {
    const uint32_t count = f.nSpectra * K;
    f.runs.clear();
    f.runBuffers.clear();
    if (count == 0) return;
    if (zeroCopy) {
      for (uint32_t i = 0; i < count; i++) {
        char* p = static_cast<char*>(f.batch[i]->GetData());
        if (!f.runs.empty() && p == static_cast<const char*>(f.runs.back()) + size_t(f.runBuffers.back()) * bufBytes) {
          f.runBuffers.back()++;
        } else {
          f.runs.push_back(p);
          f.runBuffers.push_back(1);
        }
      }
      if (f.runs.size() == 1 || size_t(count) >= 16 * f.runs.size()) {   // runs average >= 16 buffers: DMA them in place
        m_zeroCopyBatches++;
        return;
      }
      f.runs.clear();
      f.runBuffers.clear();
    }
    for (uint32_t i = 0; i < count; i++)
      memcpy(static_cast<char*>(f.staging) + size_t(i) * bufBytes, f.batch[i]->GetData(), bufBytes);
    f.runs.push_back(f.staging);
    f.runBuffers.push_back(count);
  };
  auto submit = [&](InFlight& f) {
    if (f.nSpectra == 0) return;
    if (scn_submit_gather(ctx, f.runs.data(), f.runBuffers.data(), uint32_t(f.runs.size()), f.nSpectra, &f.ticket) != SCN_OK)
      Die("scn_submit_gather");
    m_launches++;
  };
  bool processedAny = false;
  uint64_t lastSequenceId = 0;
  double lastFrequency = 0.0;
  auto finish = [&](InFlight& f) {
    const uint32_t nSpectra = f.nSpectra;
    if (nSpectra) {
      if (scn_collect_view(ctx, f.ticket, nullptr, &counts, &hits, &tdmm) != SCN_OK) Die("scn_collect_view");
      std::unique_lock<std::mutex> lock(g_printMutex, std::defer_lock);
      if (m_out || m_sink || !m_fileNameBase.empty()) lock.lock();   // output order and the trigger/record bookkeeping
      for (uint32_t s = 0; s < nSpectra; s++) {
        // the first message of the group carries the spectrum's identity
        SampleQueue::MessageHeader& header = f.batch[size_t(s) * K]->GetHeader();
        for (uint32_t k = 0; k < K; k++) {
          SampleQueue::MessageHeader& h = f.batch[size_t(s) * K + k]->GetHeader();
          if (h.m_time != 0 && m_out) {                   // process.cpp:280-287
            char tbuf[64];
            TimeToString(h.m_time, tbuf, sizeof(tbuf));
            fprintf(m_out, "Start scan at %s\n", tbuf);
            fflush(m_out);
          }
        }
        bool doWrite = false;
        if (m_mode == TimeDomain) {
          doWrite = counts[s] != 0;                       // process.cpp:226-235
          if (doWrite && m_out) {
            fprintf(m_out, "Sequence[%llu]: ", (unsigned long long)header.m_sequenceId);
            fprintf(m_out, "Max signal %f above threshold %f frequency %.0f, min %f\n", tdmm[2 * s],
                    m_threshold, header.m_frequency, tdmm[2 * s + 1]);
          }
        } else {
          const uint32_t c = counts[s];
          const scn_hit* list = &hits[size_t(s) * cap];
          if (c > cap) {                                  // overflow: this spectrum alone, full capacity
            if (!fullCtx) {
              fullCtx = CreateContext(q->m_kind, q->GetEnob(), q->GetCorrectDCOffset(), 1, 0);
              fullHits.resize(N);
            }
            uint32_t c2 = 0;
            overflowRaw.resize(size_t(K) * bufBytes);
            for (uint32_t k = 0; k < K; k++)
              memcpy(overflowRaw.data() + size_t(k) * bufBytes, f.batch[size_t(s) * K + k]->GetData(), bufBytes);
            if (scn_process_host(fullCtx, overflowRaw.data(), 1, nullptr, nullptr,
                                 &c2, fullHits.data(), nullptr) != SCN_OK)
              Die("scn_process_host");
            list = fullHits.data();
          }
          for (uint32_t r = 0; r < c && r < N; r++) {
            const scn_hit& h = list[r];
            const uint64_t hz = scn_hit_frequency(header.m_frequency, m_sampleRate, N, h.bin);
            if (m_out) fprintf(m_out, "freq %lu power_db %f\n", (unsigned long)hz, h.power_db);   // process.cpp:57
            if (m_sink) m_sink(Detection{header.m_sequenceId, header.m_frequency, hz, h.power_db, h.bin});
          }
          m_hitCount += c;
          doWrite = c > 1047;                             // process.cpp:62
        }
        if (doWrite) {
          if (m_out) fflush(m_out);
        } else {
          q->SendAck();                                   // process.cpp:303-307
        }
        ProcessWrite(doWrite, header.m_frequency, header.m_sequenceId);
        processedAny = true;
        lastSequenceId = header.m_sequenceId;
        lastFrequency = header.m_frequency;
      }
    }
    q->MessageProcessed(f.batch);                         // process.cpp:309, one lock for the batch
    m_buffersProcessed += f.batch.size();
    f.batch.clear();
    f.active = false;
  };

  uint32_t cur = 0;
  while (true) {
    InFlight& next = slot[cur];
    InFlight& prev = slot[cur ^ 1];
    const uint32_t n = q->GetNextBatch(next.batch, maxSpectra * K, K, /*wait=*/!prev.active, false,
                                       m_minBatch, m_lingerMicros);
    if (n) {
      next.nSpectra = n / K;                              // a trailing partial group at end of stream is dropped
      stage(next);
      submit(next);
      next.active = true;
    }
    if (prev.active) finish(prev);
    if (n) cur ^= 1;
    else if (!slot[0].active && !slot[1].active) {
      // nothing in flight and the non-blocking poll found nothing: block, or stop when drained
      const uint32_t m = q->GetNextBatch(slot[cur].batch, maxSpectra * K, K, /*wait=*/true, false,
                                         m_minBatch, m_lingerMicros);
      if (m == 0) break;
      InFlight& f = slot[cur];
      f.nSpectra = m / K;
      stage(f);
      submit(f);
      f.active = true;
      cur ^= 1;
    }
  }
  if (processedAny) {
    // "Shutdown writing gracefully" (process.cpp:311-313): a window that ends at or before the last message closes
    uint64_t cur = m_endSequenceId;
    while (cur < lastSequenceId && !m_endSequenceId.compare_exchange_weak(cur, lastSequenceId)) {}
    ProcessWrite(false, lastFrequency, lastSequenceId);
  }
  for (auto& s : slot) scn_free_pinned(s.staging);
  if (fullCtx) scn_destroy(fullCtx);
  scn_destroy(ctx);
}

bool ProcessSamples::StartProcessing(SampleQueue& sampleQueue) {
  m_sampleQueue = &sampleQueue;
  if (m_zeroCopy) sampleQueue.SetFifoPool(true);
  if (!m_fileNameBase.empty() && sampleQueue.m_kind != SampleQueue::FloatComplex && !m_writeCtx) {
    // recording writes fftwf_complex (messageQueue.h:127-130); the queue holds raw samples, so its writer thread
    // converts each recorded message with the reference's converter arithmetic on the GPU
    m_writeCtx.reset(CreateContext(sampleQueue.m_kind, sampleQueue.GetEnob(), sampleQueue.GetCorrectDCOffset(), 1, 1),
                     [](scn_ctx* c) { scn_destroy(c); });
    std::shared_ptr<scn_ctx> ctx = m_writeCtx;
    sampleQueue.SetWriteConverter([ctx](const void* raw, uint32_t nBuffers, float* out) {
      return scn_convert_host(ctx.get(), raw, nBuffers, out) == SCN_OK;
    });
  }
  for (uint32_t t = 0; t < m_threadCount; t++) {
    if (m_out) fprintf(m_out, "Starting process thread %u\n", t);
    m_threads[t] = new std::thread(&ProcessSamples::ThreadWorker, this, t);
  }
  for (uint32_t t = 0; t < m_threadCount; t++) {
    m_threads[t]->join();
    delete m_threads[t];
    m_threads[t] = nullptr;
    if (m_out) fprintf(m_out, "Stopped process thread %u\n", t);
  }
  return true;
}

void ProcessSamples::Run(int16_t sample_buffer[][2], uint32_t centerFrequency) {
  // Synchronous single-buffer path (process.cpp:131-144): convert -> window -> FFT -> detect on raw
  // int16 IQ.  (The reference's version dereferences a null header in process_fft and crashes;
  // here the centre frequency argument is used.)
  // the context (device tables, ticket slots) is created on the first call and kept: Run() is a per-buffer call
  if (!m_runCtx) m_runCtx = CreateContext(SampleQueue::ShortComplex, m_enob, false, 1, 0);
  scn_ctx* ctx = m_runCtx;
  std::vector<uint32_t> count(1);
  std::vector<scn_hit> hits(m_sampleCount);
  if (m_mode == FrequencyDomain) {
    if (scn_process_host(ctx, sample_buffer, 1, nullptr, nullptr, count.data(), hits.data(), nullptr) != SCN_OK)
      Die("scn_process_host");
    for (uint32_t r = 0; r < count[0]; r++) {
      const uint64_t hz = scn_hit_frequency(double(centerFrequency), m_sampleRate, m_sampleCount, hits[r].bin);
      if (m_out) fprintf(m_out, "freq %lu power_db %f\n", (unsigned long)hz, hits[r].power_db);
      if (m_sink) m_sink(Detection{0, double(centerFrequency), hz, hits[r].power_db, hits[r].bin});
    }
    m_hitCount += count[0];
  }
  m_launches++;
}
