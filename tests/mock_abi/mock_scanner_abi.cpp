// mock_scanner_abi.cpp -- TEST INFRASTRUCTURE ONLY.  A CPU stand-in for the part of the C ABI
// (include/scanner_b200.h) that the host plugin layer calls, implemented with the oracle
// (oracle/scanner_oracle.cpp).  It exists so that the HOST logic -- SampleQueue, ProcessSamples' batching /
// ticket / overflow / trigger / record code, the sources, scan_b200's wiring -- can be run against the
// reference's golden stdout on a machine without a GPU and under ThreadSanitizer (tests/test_host_mock.py).
// It is never built into libscanner_b200.so and never shipped: the product library has no CPU path.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scanner_b200.h"

extern "C" {
uint32_t orc_bytes_per_sample(uint32_t kind);
void orc_convert(uint32_t kind, const void* raw, float* dst, uint32_t n, uint32_t enob, uint32_t correct_dc);
void orc_window_build(int type, uint32_t n, float* out);
uint32_t orc_use_window(double use_bandwidth, uint32_t n);
uint64_t orc_hit_frequency(double center, uint32_t sample_rate, uint32_t n, uint32_t i);
uint32_t orc_frequency_table(uint32_t sample_rate, double start, double stop, double use_bw, double dc_ignore,
                             double* out, uint32_t cap);
void orc_pipeline(uint32_t n, uint32_t sample_rate, uint32_t enob, uint32_t kind, uint32_t correct_dc,
                  uint32_t averaging, float threshold, uint32_t use_window, uint32_t dc_ignore_window, int db_variant,
                  const float* window, const void* raw, uint32_t n_spectra, int precision, float* spectra_db,
                  double* spectra_db64, uint32_t* masks, uint32_t* counts, uint32_t threads);
void orc_time_domain(uint32_t n, uint32_t enob, uint32_t kind, uint32_t correct_dc, float threshold, const void* raw,
                     uint32_t n_buffers, uint32_t* trigger, float* max_min);
}

namespace {
thread_local std::string g_error;
int fail(int code, const char* fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}
struct Result {
  bool busy = false;
  uint32_t n = 0;
  // The real library reads the caller's (pinned) buffers ASYNCHRONOUSLY, some time between submit and collect.
  // The mock models the worst case: it remembers only the POINTERS at submit and computes at collect, so a host
  // that recycles a slab message before its ticket is collected reads recycled bytes here and fails the goldens.
  std::vector<const void*> runs;
  std::vector<uint32_t> run_buffers;
  std::vector<float> spectra, tdmm;
  std::vector<uint32_t> masks, counts;
  std::vector<scn_hit> hits;
};
}  // namespace

struct scn_ctx {
  scn_config cfg;
  std::vector<float> window;
  uint32_t K, words, hit_cap;
  size_t buf_bytes;
  std::vector<Result> slots;
  uint64_t launches = 0;
};

static void compute(scn_ctx* c, const void* raw, uint32_t n_spectra, Result& r) {
  const uint32_t N = c->cfg.sample_count;
  r.n = n_spectra;
  r.counts.assign(n_spectra, 0);
  if (c->cfg.mode == SCN_MODE_TIME_DOMAIN) {
    r.tdmm.assign(size_t(n_spectra) * 2, 0.0f);
    orc_time_domain(N, c->cfg.enob, c->cfg.sample_kind, c->cfg.correct_dc_offset, c->cfg.threshold, raw, n_spectra,
                    r.counts.data(), r.tdmm.data());
    return;
  }
  r.spectra.assign(size_t(n_spectra) * N, 0.0f);
  r.masks.assign(size_t(n_spectra) * c->words, 0u);
  orc_pipeline(N, c->cfg.sample_rate, c->cfg.enob, c->cfg.sample_kind, c->cfg.correct_dc_offset, c->K, c->cfg.threshold,
               c->cfg.use_window, c->cfg.dc_ignore_window, 0, c->window.data(), raw, n_spectra, 0, r.spectra.data(),
               nullptr, r.masks.data(), r.counts.data(), 1);
  r.hits.assign(size_t(n_spectra) * c->hit_cap, scn_hit{0, 0.0f});
  for (uint32_t s = 0; s < n_spectra; s++) {
    uint32_t rank = 0;
    for (uint32_t i = 0; i < N && rank < c->hit_cap; i++)
      if (r.masks[size_t(s) * c->words + (i >> 5)] >> (i & 31) & 1u)
        r.hits[size_t(s) * c->hit_cap + rank++] = scn_hit{i, r.spectra[size_t(s) * N + ((i + N / 2) % N)]};
  }
}

extern "C" {
const char* scn_version(void) { return "scanner_b200 MOCK (oracle-backed, tests only)"; }
const char* scn_last_error(void) { return g_error.c_str(); }

int scn_create(const scn_config* cf, scn_ctx** out) {
  if (!cf || !out) return fail(SCN_ERR_INVALID, "null");
  const bool td = cf->mode == SCN_MODE_TIME_DOMAIN;
  if (!td && !cf->window) return fail(SCN_ERR_INVALID, "window table is NULL");
  scn_ctx* c = new scn_ctx;
  c->cfg = *cf;
  if (cf->window) c->window.assign(cf->window, cf->window + cf->sample_count);
  c->cfg.window = nullptr;
  c->K = td ? 1 : (cf->averaging ? cf->averaging : 1);
  c->words = cf->sample_count / 32;
  c->hit_cap = (cf->flags & SCN_OUT_HITS) ? (cf->max_hits_per_spectrum ? cf->max_hits_per_spectrum : cf->sample_count) : 0;
  c->buf_bytes = size_t(cf->sample_count) * orc_bytes_per_sample(cf->sample_kind);
  c->slots.resize(cf->ticket_slots ? cf->ticket_slots : 2);
  *out = c;
  return SCN_OK;
}
int scn_destroy(scn_ctx* c) { delete c; return SCN_OK; }
size_t scn_buffer_bytes(const scn_ctx* c) { return c ? c->buf_bytes : 0; }
uint32_t scn_mask_words(const scn_ctx* c) { return c ? c->words : 0; }
int scn_alloc_pinned(size_t bytes, void** out) { *out = malloc(bytes ? bytes : 1); return *out ? SCN_OK : SCN_ERR_CUDA; }
int scn_free_pinned(void* p) { free(p); return SCN_OK; }

int scn_submit_gather(scn_ctx* c, const void* const* runs, const uint32_t* run_buffers, uint32_t n_runs,
                      uint32_t n_spectra, uint32_t* ticket) {
  if (!c || !runs || !run_buffers || !ticket || n_runs == 0) return fail(SCN_ERR_INVALID, "submit: bad arguments");
  if (n_spectra > c->cfg.max_spectra) return fail(SCN_ERR_CAPACITY, "submit: %u > max_spectra %u", n_spectra, c->cfg.max_spectra);
  uint64_t total = 0;
  for (uint32_t r = 0; r < n_runs; r++) total += run_buffers[r];
  if (total != uint64_t(n_spectra) * c->K) return fail(SCN_ERR_INVALID, "submit_gather: run lengths do not add up");
  for (uint32_t s = 0; s < c->slots.size(); s++)
    if (!c->slots[s].busy) {
      Result& r = c->slots[s];
      r.runs.assign(runs, runs + n_runs);
      r.run_buffers.assign(run_buffers, run_buffers + n_runs);
      r.n = n_spectra;
      r.busy = true;
      c->launches++;
      *ticket = s;
      return SCN_OK;
    }
  return fail(SCN_ERR_BUSY, "submit: no free ticket slot");
}

int scn_submit(scn_ctx* c, const void* raw, uint32_t n_spectra, uint32_t* ticket) {
  if (!c || !raw) return fail(SCN_ERR_INVALID, "submit: bad arguments");
  const uint32_t buffers = n_spectra * c->K;
  return scn_submit_gather(c, &raw, &buffers, 1, n_spectra, ticket);
}

int scn_collect(scn_ctx* c, uint32_t ticket, float* spectra_db, uint32_t* hit_mask, uint32_t* hit_count, scn_hit* hits,
                float* td_max_min) {
  if (!c || ticket >= c->slots.size() || !c->slots[ticket].busy) return fail(SCN_ERR_INVALID, "collect: bad ticket");
  Result& r = c->slots[ticket];
  {                                           // "the DMA and the kernel happen now"
    std::vector<char> staged(size_t(r.n) * c->K * c->buf_bytes);
    size_t off = 0;
    for (size_t i = 0; i < r.runs.size(); i++) {
      memcpy(staged.data() + off, r.runs[i], size_t(r.run_buffers[i]) * c->buf_bytes);
      off += size_t(r.run_buffers[i]) * c->buf_bytes;
    }
    compute(c, staged.data(), r.n, r);
  }
  if (spectra_db && !r.spectra.empty()) memcpy(spectra_db, r.spectra.data(), r.spectra.size() * sizeof(float));
  if (hit_mask && !r.masks.empty()) memcpy(hit_mask, r.masks.data(), r.masks.size() * sizeof(uint32_t));
  if (hit_count) memcpy(hit_count, r.counts.data(), r.counts.size() * sizeof(uint32_t));
  if (hits && !r.hits.empty()) memcpy(hits, r.hits.data(), r.hits.size() * sizeof(scn_hit));
  if (td_max_min && !r.tdmm.empty()) memcpy(td_max_min, r.tdmm.data(), r.tdmm.size() * sizeof(float));
  r.busy = false;
  return SCN_OK;
}

int scn_collect_view(scn_ctx* c, uint32_t ticket, const uint32_t** hit_mask, const uint32_t** hit_count,
                     const scn_hit** hits, const float** td_max_min) {
  int rc = scn_collect(c, ticket, nullptr, nullptr, nullptr, nullptr, nullptr);    // computes into the slot
  if (rc != SCN_OK) return rc;
  Result& r = c->slots[ticket];
  if (hit_mask) *hit_mask = r.masks.empty() ? nullptr : r.masks.data();
  if (hit_count) *hit_count = r.counts.data();
  if (hits) *hits = r.hits.empty() ? nullptr : r.hits.data();
  if (td_max_min) *td_max_min = r.tdmm.empty() ? nullptr : r.tdmm.data();
  return SCN_OK;
}

int scn_process_host(scn_ctx* c, const void* raw, uint32_t n_spectra, float* spectra_db, uint32_t* hit_mask,
                     uint32_t* hit_count, scn_hit* hits, float* td_max_min) {
  if (!c || (n_spectra && !raw)) return fail(SCN_ERR_INVALID, "process_host: bad arguments");
  const uint32_t N = c->cfg.sample_count;
  for (uint32_t first = 0; first < n_spectra; first += c->cfg.max_spectra) {
    const uint32_t count = n_spectra - first < c->cfg.max_spectra ? n_spectra - first : c->cfg.max_spectra;
    uint32_t t = 0;
    int rc = scn_submit(c, static_cast<const char*>(raw) + size_t(first) * c->K * c->buf_bytes, count, &t);
    if (rc != SCN_OK) return rc;
    rc = scn_collect(c, t, spectra_db ? spectra_db + size_t(first) * N : nullptr,
                     hit_mask ? hit_mask + size_t(first) * c->words : nullptr, hit_count ? hit_count + first : nullptr,
                     hits ? hits + size_t(first) * c->hit_cap : nullptr, td_max_min ? td_max_min + size_t(first) * 2 : nullptr);
    if (rc != SCN_OK) return rc;
  }
  return SCN_OK;
}

int scn_convert_host(scn_ctx* c, const void* raw, uint32_t n_buffers, float* out) {
  if (!c || (n_buffers && (!raw || !out))) return fail(SCN_ERR_INVALID, "convert: bad arguments");
  const uint32_t N = c->cfg.sample_count;
  for (uint32_t b = 0; b < n_buffers; b++)
    orc_convert(c->cfg.sample_kind, static_cast<const char*>(raw) + size_t(b) * c->buf_bytes, out + size_t(b) * 2 * N, N,
                c->cfg.enob, c->cfg.correct_dc_offset);
  return SCN_OK;
}

uint32_t scn_use_window(double use_bandwidth, uint32_t n) { return orc_use_window(use_bandwidth, n); }
uint64_t scn_hit_frequency(double center, uint32_t fs, uint32_t n, uint32_t i) { return orc_hit_frequency(center, fs, n, i); }
uint32_t scn_frequency_table(uint32_t fs, double start, double stop, double use_bw, double dc_ignore, double* out,
                             uint32_t cap) {
  return orc_frequency_table(fs, start, stop, use_bw, dc_ignore, out, cap);
}
int scn_window_build(int type, uint32_t n, float* out) { orc_window_build(type, n, out); return SCN_OK; }

void scn_shard_steps(uint32_t n_steps, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* end) {
  if (world == 0) world = 1;
  if (begin) *begin = uint32_t(uint64_t(n_steps) * rank / world);
  if (end) *end = uint32_t(uint64_t(n_steps) * (rank + 1) / world);
}
}  // extern "C"

// ---- record exchange / NCCL gather stand-ins (csrc/host/sweepProcessor.cpp): plain host memory, same semantics ----
struct scn_exchange {
  uint32_t rank, world, rec_total, rec_words;
  uint64_t seq = 0;
  std::vector<scn_exchange*> peers;
  std::vector<std::vector<uint32_t>> rows[4];     // [slot][rank] -> that rank's table
  std::vector<uint64_t> flags[4];
};
struct scn_gather { uint32_t n_devices, rec_total, rec_words; };
static void merge_rows(const uint32_t* const* parts, uint32_t n_parts, uint32_t total, uint32_t rec_words, uint32_t* out) {
  for (uint32_t x = 0; x < total; x++) {
    uint32_t v = 0;
    for (uint32_t p = 0; p < n_parts; p++) v = (x % rec_words) < 2 ? v + parts[p][x] : (v | parts[p][x]);
    out[x] = v;
  }
}
extern "C" {
int scn_exchange_create(int, uint32_t rank, uint32_t world, uint32_t n_steps, uint32_t record_words, scn_exchange** out) {
  scn_exchange* x = new scn_exchange();
  x->rank = rank; x->world = world; x->rec_words = record_words; x->rec_total = n_steps * record_words;
  for (int sl = 0; sl < 4; sl++) { x->rows[sl].assign(world, std::vector<uint32_t>(x->rec_total, 0u)); x->flags[sl].assign(world, 0); }
  x->peers.assign(world, nullptr);
  x->peers[rank] = x;
  *out = x;
  return SCN_OK;
}
int scn_exchange_connect_local(scn_exchange* const* all, uint32_t world) {
  for (uint32_t a = 0; a < world; a++) all[a]->peers.assign(all, all + world);
  return SCN_OK;
}
int scn_exchange_publish_host(scn_exchange* x, const uint32_t* host_records, uint64_t* seq_out) {
  const uint64_t seq = ++x->seq;
  for (scn_exchange* p : x->peers) {
    if (!p) return fail(SCN_ERR_INVALID, "exchange: peers are not connected");
    p->rows[seq % 4][x->rank].assign(host_records, host_records + x->rec_total);
    p->flags[seq % 4][x->rank] = seq;
  }
  if (seq_out) *seq_out = seq;
  return SCN_OK;
}
int scn_exchange_merge_host(scn_exchange* x, uint64_t seq, uint32_t* host_merged) {
  std::vector<const uint32_t*> parts;
  for (uint32_t r = 0; r < x->world; r++) {
    if (x->flags[seq % 4][r] != seq) return fail(SCN_ERR_CUDA, "exchange: rank %u has not published %llu", r, (unsigned long long)seq);
    parts.push_back(x->rows[seq % 4][r].data());
  }
  merge_rows(parts.data(), x->world, x->rec_total, x->rec_words, host_merged);
  return SCN_OK;
}
int scn_exchange_destroy(scn_exchange* x) { delete x; return SCN_OK; }
int scn_nccl_gather_create(const int*, uint32_t n_devices, uint32_t n_steps, uint32_t record_words, scn_gather** out) {
  *out = new scn_gather{n_devices, n_steps * record_words, record_words};
  return SCN_OK;
}
int scn_nccl_gather_merge_host(scn_gather* g, const uint32_t* const* host_partials, uint32_t* host_merged) {
  merge_rows(host_partials, g->n_devices, g->rec_total, g->rec_words, host_merged);
  return SCN_OK;
}
int scn_nccl_gather_destroy(scn_gather* g) { delete g; return SCN_OK; }
}
