#include "sweepProcessor.h"

#include <cassert>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>

namespace {
[[noreturn]] void Die(const char* what) {
  // the reference's error convention in this layer: message to stderr, exit(1)
  fprintf(stderr, "%s: %s\n", what, scn_last_error());
  exit(1);
}
// Hit records copied back per spectrum with every batch; a spectrum with more is re-run alone at full capacity
// (same rule as ProcessSamples::ThreadWorker).
const uint32_t kHitCap = 64;

void AppendF(std::string& s, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void AppendF(std::string& s, const char* fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  const int n = vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (n > 0) s.append(buf, size_t(n) < sizeof(buf) ? size_t(n) : sizeof(buf) - 1);
}
}  // namespace

SweepProcessor::SweepProcessor(uint32_t numSamples, uint32_t sampleRate, uint32_t enob, float threshold,
                               int windowType, const std::vector<double>& stepFrequencies,
                               const std::vector<int>& devices, Exchange exchange, double useBandWidth)
    : m_sampleCount(numSamples), m_sampleRate(sampleRate), m_enob(enob), m_threshold(threshold),
      m_stepFrequencies(stepFrequencies), m_devices(devices), m_exchange(exchange),
      m_useWindow(scn_use_window(useBandWidth, numSamples)), m_words(numSamples / 32), m_recWords(numSamples / 32 + 2),
      m_window(numSamples), m_perDevice(devices.size()) {
  assert(!devices.empty() && !stepFrequencies.empty());
  if (scn_window_build(windowType, numSamples, m_window.data()) != SCN_OK) Die("FFTWindow");
  for (uint32_t i = 0; i < m_stepFrequencies.size(); i++) m_stepIndex.emplace(m_stepFrequencies[i], i);
  for (auto& c : m_perDevice) c = 0;
}

SweepProcessor::~SweepProcessor() {
  if (m_gather) scn_nccl_gather_destroy(m_gather);
  for (scn_exchange* x : m_windows) scn_exchange_destroy(x);
}

uint32_t SweepProcessor::GetStepOwner(uint32_t step) const {
  const uint32_t G = uint32_t(m_devices.size()), S = uint32_t(m_stepFrequencies.size());
  for (uint32_t d = 0; d < G; d++) {
    uint32_t b = 0, e = 0;
    scn_shard_steps(S, d, G, &b, &e);          // contiguous, balanced ranges: the same split bench.py's ranks use
    if (step >= b && step < e) return d;
  }
  return G - 1;
}

uint32_t SweepProcessor::StepOf(double frequency) const {
  auto it = m_stepIndex.find(frequency);
  if (it != m_stepIndex.end()) return it->second;
  // not a table entry (a source that reports the tuned, not the requested, frequency): nearest step
  uint32_t best = 0;
  for (uint32_t i = 1; i < m_stepFrequencies.size(); i++)
    if (std::fabs(m_stepFrequencies[i] - frequency) < std::fabs(m_stepFrequencies[best] - frequency)) best = i;
  return best;
}

scn_ctx* SweepProcessor::CreateContext(int device, uint32_t maxSpectra, uint32_t hitCap) {
  scn_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.device = device;
  cfg.sample_count = m_sampleCount;
  cfg.sample_rate = m_sampleRate;
  cfg.enob = m_queue->GetEnob();
  cfg.sample_kind = uint32_t(m_queue->m_kind);
  cfg.correct_dc_offset = m_queue->GetCorrectDCOffset() ? 1 : 0;
  cfg.averaging = m_averaging;
  cfg.mode = SCN_MODE_FREQUENCY_DOMAIN;
  cfg.threshold = m_threshold;
  cfg.use_window = m_useWindow;
  cfg.dc_ignore_window = 4;                        // process.cpp:87
  cfg.window = m_window.data();
  cfg.max_spectra = maxSpectra;
  cfg.max_hits_per_spectrum = hitCap;
  cfg.flags = SCN_OUT_HITS;
  cfg.ticket_slots = 2;
  scn_ctx* ctx = nullptr;
  if (scn_create(&cfg, &ctx) != SCN_OK) Die("scn_create");
  return ctx;
}

void SweepProcessor::Push(uint32_t d, Item&& item) {
  Inbox& in = *m_inbox[d];
  std::unique_lock<std::mutex> lock(in.mutex);
  in.items.push_back(std::move(item));
  in.ready.notify_one();
}

// Output leaves in queue order whatever GPU produced it: every spectrum (and every sweep report) was given its
// position by the router; text that arrives early waits for its predecessors.
void SweepProcessor::Emit(uint64_t index, std::string&& text) {
  std::unique_lock<std::mutex> lock(m_emitMutex);
  m_pending.emplace(index, std::move(text));
  while (!m_pending.empty() && m_pending.begin()->first == m_nextEmit) {
    if (m_out && !m_pending.begin()->second.empty()) fputs(m_pending.begin()->second.c_str(), m_out);
    m_pending.erase(m_pending.begin());
    m_nextEmit++;
  }
  if (m_out) fflush(m_out);
}

// End of a sweep on device d: wait for the other GPUs; the last one to arrive exchanges the partial tables.
void SweepProcessor::SweepArrive(uint32_t d, uint64_t reportEmit) {
  (void)d;
  std::unique_lock<std::mutex> lock(m_sweepMutex);
  const uint64_t generation = m_sweepGeneration;
  if (++m_sweepArrived == m_devices.size()) {
    ExchangeAndReport(reportEmit);
    m_sweepArrived = 0;
    m_sweepGeneration++;
    m_sweepCv.notify_all();
  } else {
    m_sweepCv.wait(lock, [&] { return m_sweepGeneration != generation; });
  }
}

void SweepProcessor::ExchangeAndReport(uint64_t reportEmit) {
  const uint32_t G = uint32_t(m_devices.size()), S = uint32_t(m_stepFrequencies.size());
  std::vector<uint32_t> merged(size_t(S) * m_recWords, 0u);
  if (m_exchange == NcclAllGather) {
    // ncclCommInitAll once, then one grouped all-gather of the G partial tables + the merge kernel on every GPU
    if (!m_gather && scn_nccl_gather_create(m_devices.data(), G, S, m_recWords, &m_gather) != SCN_OK)
      Die("scn_nccl_gather_create");
    std::vector<const uint32_t*> parts(G);
    for (uint32_t d = 0; d < G; d++) parts[d] = m_partial[d].data();
    if (scn_nccl_gather_merge_host(m_gather, parts.data(), merged.data()) != SCN_OK) Die("scn_nccl_gather_merge_host");
  } else {
    // NVLink peer-memory windows: every GPU stores its table into every peer's window, then merges its own window
    if (m_windows.empty()) {
      m_windows.resize(G, nullptr);
      for (uint32_t d = 0; d < G; d++)
        if (scn_exchange_create(m_devices[d], d, G, S, m_recWords, &m_windows[d]) != SCN_OK) Die("scn_exchange_create");
      if (G > 1 && scn_exchange_connect_local(m_windows.data(), G) != SCN_OK) Die("scn_exchange_connect_local");
    }
    uint64_t seq = 0;
    for (uint32_t d = 0; d < G; d++)
      if (scn_exchange_publish_host(m_windows[d], m_partial[d].data(), &seq) != SCN_OK) Die("scn_exchange_publish_host");
    std::vector<uint32_t> other(merged.size());
    for (uint32_t d = 0; d < G; d++) {
      if (scn_exchange_merge_host(m_windows[d], seq, d == 0 ? merged.data() : other.data()) != SCN_OK)
        Die("scn_exchange_merge_host");
      if (d > 0 && other != merged) { fprintf(stderr, "SweepProcessor: GPU %u disagrees after the exchange\n", d); exit(1); }
    }
  }
  std::string text;
  if (m_sweepReport) {
    const uint32_t sweep = m_sweepsDone;
    for (uint32_t s = 0; s < S; s++) {
      const uint32_t* r = &merged[size_t(s) * m_recWords];
      AppendF(text, "sweep %u step %u freq %.0f spectra %u hits %u\n", sweep, s, m_stepFrequencies[s], r[1], r[0]);
    }
  }
  m_lastRecords.swap(merged);
  for (auto& p : m_partial) std::fill(p.begin(), p.end(), 0u);
  m_sweepsDone++;
  Emit(reportEmit, std::move(text));
}

void SweepProcessor::Worker(uint32_t d) {
  SampleQueue* q = m_queue;
  const uint32_t K = m_averaging, N = m_sampleCount;
  const uint32_t maxSpectra = (m_maxBatch + K - 1) / K;
  const uint32_t cap = kHitCap < N ? kHitCap : N;
  scn_ctx* ctx = CreateContext(m_devices[d], maxSpectra, cap);
  scn_ctx* fullCtx = nullptr;
  std::vector<scn_hit> fullHits;
  std::vector<char> overflowRaw;
  const size_t bufBytes = q->GetBufferBytes();
  std::vector<uint32_t>& partial = m_partial[d];

  struct InFlight {
    std::vector<Item> items;
    std::vector<const void*> runs;
    std::vector<uint32_t> runBuffers;
    void* staging = nullptr;
    uint32_t ticket = 0;
    bool active = false;
  } slot[2];
  for (auto& s : slot)
    if (scn_alloc_pinned(bufBytes * size_t(maxSpectra) * K, &s.staging) != SCN_OK) Die("scn_alloc_pinned");

  auto submit = [&](InFlight& f) {
    // address runs of the queue's pinned slab go to the GPU as they lie; short runs are packed first
    f.runs.clear();
    f.runBuffers.clear();
    uint32_t count = 0;
    for (Item& it : f.items)
      for (SampleQueue::MessageType* m : it.msgs) {
        char* p = static_cast<char*>(m->GetData());
        if (q->IsPinnedSlab() && !f.runs.empty() &&
            p == static_cast<const char*>(f.runs.back()) + size_t(f.runBuffers.back()) * bufBytes) {
          f.runBuffers.back()++;
        } else {
          f.runs.push_back(p);
          f.runBuffers.push_back(1);
        }
        count++;
      }
    if (!q->IsPinnedSlab() || (f.runs.size() > 1 && size_t(count) < 16 * f.runs.size())) {
      size_t off = 0;
      for (Item& it : f.items)
        for (SampleQueue::MessageType* m : it.msgs) {
          memcpy(static_cast<char*>(f.staging) + off, m->GetData(), bufBytes);
          off += bufBytes;
        }
      f.runs.assign(1, f.staging);
      f.runBuffers.assign(1, count);
    }
    if (scn_submit_gather(ctx, f.runs.data(), f.runBuffers.data(), uint32_t(f.runs.size()), uint32_t(f.items.size()),
                          &f.ticket) != SCN_OK)
      Die("scn_submit_gather");
    m_launches++;
    f.active = true;
  };

  std::vector<SampleQueue::MessageType*> done;
  auto finish = [&](InFlight& f) {
    const uint32_t* counts = nullptr;
    const uint32_t* masks = nullptr;
    const scn_hit* hits = nullptr;
    if (scn_collect_view(ctx, f.ticket, &masks, &counts, &hits, nullptr) != SCN_OK) Die("scn_collect_view");
    done.clear();
    for (size_t s = 0; s < f.items.size(); s++) {
      Item& it = f.items[s];
      SampleQueue::MessageHeader& header = it.msgs[0]->GetHeader();     // the group's first message is its identity
      std::string text;
      for (SampleQueue::MessageType* m : it.msgs) {
        if (m->GetHeader().m_time != 0) {                               // process.cpp:280-287
          char tbuf[64];
          struct tm tmv;
          time_t t = m->GetHeader().m_time;
          if (localtime_r(&t, &tmv) == nullptr || strftime(tbuf, sizeof(tbuf), "%Y%m%d-%T", &tmv) == 0) exit(1);
          AppendF(text, "Start scan at %s\n", tbuf);
        }
      }
      const uint32_t c = counts[s];
      const scn_hit* list = hits + size_t(s) * cap;
      if (c > cap) {                                                     // overflow: this spectrum alone, full capacity
        if (!fullCtx) {
          fullCtx = CreateContext(m_devices[d], 1, 0);
          fullHits.resize(N);
        }
        overflowRaw.resize(size_t(K) * bufBytes);
        for (uint32_t k = 0; k < K; k++) memcpy(overflowRaw.data() + size_t(k) * bufBytes, it.msgs[k]->GetData(), bufBytes);
        uint32_t c2 = 0;
        if (scn_process_host(fullCtx, overflowRaw.data(), 1, nullptr, nullptr, &c2, fullHits.data(), nullptr) != SCN_OK)
          Die("scn_process_host");
        list = fullHits.data();
      }
      for (uint32_t r = 0; r < c && r < N; r++) {
        const uint64_t hz = scn_hit_frequency(header.m_frequency, m_sampleRate, N, list[r].bin);
        AppendF(text, "freq %lu power_db %f\n", (unsigned long)hz, list[r].power_db);     // process.cpp:57
      }
      // this GPU's record of the step (scn_records.cu layout)
      uint32_t* rec = &partial[size_t(it.step) * m_recWords];
      rec[0] += c;
      rec[1] += 1;
      const uint32_t* mw = masks + s * size_t(m_words);
      for (uint32_t w = 0; w < m_words; w++) rec[2 + w] |= mw[w];
      m_hitCount += c;
      q->SendAck();                                                      // process.cpp:303-307
      Emit(it.emit, std::move(text));
      for (SampleQueue::MessageType* m : it.msgs) done.push_back(m);
    }
    q->MessageProcessed(done);                                           // process.cpp:309
    m_buffersProcessed += done.size();
    m_perDevice[d] += done.size();
    f.items.clear();
    f.active = false;
  };

  Inbox& in = *m_inbox[d];
  uint32_t cur = 0;
  bool stop = false;
  while (!stop) {
    InFlight& next = slot[cur];
    InFlight& prev = slot[cur ^ 1];
    Item marker;
    bool haveMarker = false;
    {
      std::unique_lock<std::mutex> lock(in.mutex);
      // with a batch in flight never sleep on the inbox: collect it instead
      if (!prev.active) in.ready.wait(lock, [&] { return !in.items.empty(); });
      while (!in.items.empty() && next.items.size() < maxSpectra) {
        if (in.items.front().sweepEnd || in.items.front().stop) {
          if (!next.items.empty()) break;                                // the batch before the marker goes first
          marker = std::move(in.items.front());
          in.items.pop_front();
          haveMarker = true;
          break;
        }
        next.items.push_back(std::move(in.items.front()));
        in.items.pop_front();
      }
    }
    const bool submitted = !next.items.empty();
    if (submitted) submit(next);
    if (prev.active) finish(prev);
    if (submitted) cur ^= 1;
    if (haveMarker) {
      for (auto& s : slot) if (s.active) finish(s);                      // (submitted is false here: nothing newer)
      if (marker.sweepEnd) SweepArrive(d, marker.emit);
      if (marker.stop) stop = true;
    }
  }
  for (auto& s : slot) scn_free_pinned(s.staging);
  if (fullCtx) scn_destroy(fullCtx);
  scn_destroy(ctx);
}

bool SweepProcessor::StartProcessing(SampleQueue& sampleQueue) {
  m_queue = &sampleQueue;
  const uint32_t G = uint32_t(m_devices.size()), S = uint32_t(m_stepFrequencies.size()), K = m_averaging;
  m_inbox.clear();
  m_partial.assign(G, std::vector<uint32_t>(size_t(S) * m_recWords, 0u));
  for (uint32_t d = 0; d < G; d++) m_inbox.emplace_back(new Inbox());
  std::vector<uint32_t> owner(S);
  for (uint32_t s = 0; s < S; s++) owner[s] = GetStepOwner(s);
  for (uint32_t d = 0; d < G; d++) {
    if (m_out) fprintf(m_out, "Starting process thread %u on GPU %d\n", d, m_devices[d]);
    m_workers.emplace_back(&SweepProcessor::Worker, this, d);
  }
  // ---- router: queue order -> (GPU that owns the step, position in the output) ----------------------------------
  std::vector<SampleQueue::MessageType*> batch;
  uint64_t emit = 0;
  bool openSweep = false;
  auto endSweep = [&]() {
    for (uint32_t d = 0; d < G; d++) {
      Item m;
      m.sweepEnd = true;
      m.emit = emit;                         // all GPUs carry the same report position
      Push(d, std::move(m));
    }
    emit++;
    openSweep = false;
  };
  while (uint32_t n = sampleQueue.GetNextBatch(batch, m_maxBatch * K, K)) {
    for (uint32_t i = 0; i + K <= n; i += K) {
      // a scan-start marker on the group's first message closes the sweep before it (messageQueue.h:67-72 stamps
      // the first buffer of every sweep)
      if (batch[i]->GetHeader().m_time != 0 && openSweep) endSweep();
      Item it;
      it.msgs.assign(batch.begin() + i, batch.begin() + i + K);
      it.step = StepOf(batch[i]->GetHeader().m_frequency);
      it.emit = emit++;
      openSweep = true;
      Push(owner[it.step], std::move(it));
    }
    if (n % K) {                             // a trailing partial group at the end of the stream is dropped
      std::vector<SampleQueue::MessageType*> rest(batch.begin() + (n - n % K), batch.begin() + n);
      sampleQueue.MessageProcessed(rest);
    }
  }
  if (openSweep) endSweep();
  for (uint32_t d = 0; d < G; d++) {
    Item m;
    m.stop = true;
    Push(d, std::move(m));
  }
  for (uint32_t d = 0; d < G; d++) {
    m_workers[d].join();
    if (m_out) fprintf(m_out, "Stopped process thread %u\n", d);
  }
  m_workers.clear();
  return true;
}
