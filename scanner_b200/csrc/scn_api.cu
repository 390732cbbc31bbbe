// scn_api.cu -- C ABI (include/scanner_b200.h) over the fused sm_100a kernels.
//
// Replaces the per-message body of ProcessSamples::ThreadWorker (process.cpp:292-299) and
// ProcessSamples::Run (process.cpp:131-144) with one batched launch; see the header for the
// interface-by-interface citations.  There is no CPU fallback anywhere in this file: when
// CUDA is unavailable every entry point fails with SCN_ERR_NO_DEVICE / SCN_ERR_CUDA.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scanner_b200.h"
#include "scn_dispatch.h"
#include "scn_timedomain.cuh"

namespace scn {
cudaError_t launch_summarize(const uint32_t* masks, const uint32_t* counts, uint32_t n_spectra,
                             uint64_t first_unit, uint32_t units_per_step, uint32_t n_steps, uint32_t words,
                             uint32_t* records, int num_sms, cudaStream_t stream);
cudaError_t launch_convert(int kind, const void* raw, float* out, uint32_t n, uint32_t n_buffers, float onebymax,
                           bool correct_dc, int num_sms, cudaStream_t stream);
cudaError_t launch_hackrf_prepass(void* transfers, uint32_t n_transfers, uint32_t valid_length,
                                  uint64_t* frequency_hz, uint32_t* status, cudaStream_t stream);
cudaError_t launch_merge(const uint32_t* parts, uint32_t n_parts, uint32_t n_steps, uint32_t rec_words,
                         uint32_t* out, cudaStream_t stream);
// scn_large.cu: four-step path for N = 2^15, 2^16
cudaError_t launch_dc_sums(uint32_t kind, const uint8_t* raw, uint32_t n_buffers, uint32_t n, int2* dcs,
                           int num_sms, cudaStream_t stream);
cudaError_t launch_columns(uint32_t kind, const uint8_t* raw, uint32_t n_buffers, uint32_t n2_count,
                           const float* window, const float2* wn, const int2* dcs, float2* y, int num_sms,
                           cudaStream_t stream);
cudaError_t launch_finalize(const float* power, uint32_t n_spectra, uint32_t n2_count, float threshold,
                            uint32_t use_window, uint32_t dc_ignore, float* spectra, uint32_t* masks,
                            uint32_t* counts, scn_hit* hits, uint32_t hit_cap, int num_sms, cudaStream_t stream);
constexpr int kMaxLog2NLarge = 16;   // sizes above kMaxLog2N go through the four-step path
}  // namespace scn

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

}  // namespace

namespace scn {
// error text for the other translation units of the C ABI (scn_exchange.cu)
int api_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
}  // namespace scn

namespace {

#define SCN_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver          \
                      ? SCN_ERR_NO_DEVICE : SCN_ERR_CUDA,                               \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

constexpr double kPi = 3.14159265358979323846264338327950288;

uint32_t bytes_per_sample(uint32_t kind) {
  switch (kind) {
    case SCN_KIND_BYTE_COMPLEX: return 2;
    case SCN_KIND_SHORT: return 4;
    case SCN_KIND_SHORT_COMPLEX: return 4;
    case SCN_KIND_FLOAT_COMPLEX: return 8;
    default: return 0;
  }
}

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  void* h_raw = nullptr;        // pinned staging for pageable callers
  void* d_raw = nullptr;
  float* d_spectra = nullptr;
  uint32_t* d_masks = nullptr;
  uint32_t* d_counts = nullptr;
  scn_hit* d_hits = nullptr;
  float* d_td = nullptr;
  uint32_t* h_masks = nullptr;  // pinned
  uint32_t* h_counts = nullptr;
  scn_hit* h_hits = nullptr;
  float* h_td = nullptr;
  uint32_t n_spectra = 0;
  bool busy = false;
  bool ready = false;           // every buffer below was allocated (a half-built slot is freed, never used)
};

}  // namespace

struct scn_ctx {
  scn_config cfg{};
  int log2n = 0;
  uint32_t K = 1;
  bool time_domain = false;
  size_t buf_bytes = 0;
  uint32_t words = 0;
  uint32_t hit_cap = 0;
  float onebymax = 1.0f;
  float* d_window = nullptr;
  bool win_mirror = false;       // window[n] == window[N-1-n] bitwise
  float* d_ones = nullptr;       // unit window for the row transforms of the four-step path
  float2* d_twiddles = nullptr;
  scn::KernelVariant variant{};
  int ctas_per_sm = 1;
  int max_clusters = 0;          // co-resident clusters (cluster variants only)
  int num_sms = 1;
  int regs = 0;
  // four-step path (N = 2^15, 2^16): N = 16 x N2
  bool large = false;
  int log2n2 = 0;
  float2* d_wn = nullptr;        // exp(-2 pi i m / N), m < N
  float2* d_y = nullptr;         // intermediate [chunk buffers][16][N2]
  float* d_p = nullptr;          // power [chunk spectra][16][N2]
  int2* d_dcs = nullptr;         // per-buffer dc
  uint32_t chunk_spectra = 0;
  cudaEvent_t large_done = nullptr;   // the four-step scratch is per context: launches on different streams take turns
  // work counters of the persistent kernels (scn::WorkQueue): one per launch, rotated, self-resetting
  uint32_t* d_work = nullptr;
  uint32_t work_next = 0;
  std::vector<Slot> slots;
  uint32_t next_slot = 0;
  uint64_t launches = 0;
  // scn_convert_host staging (grown on demand)
  void* d_conv_raw = nullptr;
  float* d_conv_out = nullptr;
  uint32_t conv_capacity = 0;
};

namespace {

constexpr uint32_t kWorkCounters = 64;      // launches of one context that may be in flight at once
constexpr uint32_t kWorkStride = 8;         // uint32 words between counters (one 32-byte sector each)
uint32_t* next_work_counter(scn_ctx* c) {
  uint32_t* w = c->d_work + size_t(c->work_next % kWorkCounters) * kWorkStride;
  c->work_next++;
  return w;
}

int ilog2_exact(uint32_t n) {
  if (n == 0 || (n & (n - 1))) return -1;
  int l = 0;
  while ((1u << l) < n) l++;
  return l;
}

// Twiddle tables of the Stockham plan (scn_fft.cuh): passes p >= 1 are radix 16; thread t of pass p
// multiplies input r by W_{16 Ns}^{k r} = exp(-2 pi i k r / (16 Ns)), k = t mod Ns, Ns = product of the
// earlier radices.  Evaluated in double, stored thread-contiguous: tw[((p-1)*15 + r-1)*T + t].
std::vector<float2> build_twiddles(int log2n) {
  const int N = 1 << log2n, T = N / 16;
  std::vector<float2> tw(size_t(scn::total_tw_per_thread(log2n)) * T);
  for (int p = 1; p < scn::num_passes(log2n); p++) {
    const int Ns = 1 << scn::pass_log2ns(log2n, p);
    for (int r = 1; r < 16; r++)
      for (int t = 0; t < T; t++) {
        const int k = t & (Ns - 1);
        const double a = -2.0 * kPi * double(k) * double(r) / (16.0 * double(Ns));
        tw[size_t((p - 1) * 15 + (r - 1)) * T + t] = make_float2(float(std::cos(a)), float(std::sin(a)));
      }
  }
  return tw;
}

int launch_rows(scn_ctx* c, const void* d_y, uint32_t n_rows, float* d_power, cudaStream_t stream);

// N = 2^15 / 2^16: dc sums -> columns -> row FFTs (fused kernel in row mode) -> finalize, chunked so the
// intermediate stays within the context's scratch.
int launch_large(scn_ctx* c, const void* d_raw, uint32_t n_spectra, float* d_spectra, uint32_t* d_masks,
                 uint32_t* d_counts, scn_hit* d_hits, cudaStream_t stream) {
  const uint32_t N = c->cfg.sample_count, N2 = N / 16, K = c->K;
  const bool dc = c->cfg.correct_dc_offset && c->cfg.sample_kind != SCN_KIND_FLOAT_COMPLEX;
  // scratch grows on demand up to the chunk cap (intermediate <= 256 MiB)
  {
    uint64_t cap = (256ull << 20) / (uint64_t(K) * N * sizeof(float2));
    if (cap < 1) cap = 1;
    const uint32_t want = n_spectra < cap ? n_spectra : uint32_t(cap);
    if (want > c->chunk_spectra) {
      SCN_CUDA(cudaStreamSynchronize(stream));
      if (c->d_y) cudaFree(c->d_y);
      if (c->d_p) cudaFree(c->d_p);
      if (c->d_dcs) cudaFree(c->d_dcs);
      c->d_y = nullptr; c->d_p = nullptr; c->d_dcs = nullptr; c->chunk_spectra = 0;
      SCN_CUDA(cudaMalloc(&c->d_y, sizeof(float2) * size_t(want) * K * N));
      SCN_CUDA(cudaMalloc(&c->d_p, sizeof(float) * size_t(want) * N));
      SCN_CUDA(cudaMalloc(&c->d_dcs, sizeof(int2) * size_t(want) * K));
      c->chunk_spectra = want;
    }
  }
  // d_y / d_p / d_dcs belong to the context, not to the stream: two tickets (or two user streams) in flight must
  // not overwrite each other's intermediates, so every launch waits for the previous one's finalize kernel.
  if (!c->large_done) SCN_CUDA(cudaEventCreateWithFlags(&c->large_done, cudaEventDisableTiming));
  else SCN_CUDA(cudaStreamWaitEvent(stream, c->large_done, 0));
  for (uint32_t first = 0; first < n_spectra; first += c->chunk_spectra) {
    const uint32_t ns = (n_spectra - first < c->chunk_spectra) ? (n_spectra - first) : c->chunk_spectra;
    const uint32_t nb = ns * K;
    const uint8_t* raw = static_cast<const uint8_t*>(d_raw) + size_t(first) * K * c->buf_bytes;
    if (dc) {
      SCN_CUDA(scn::launch_dc_sums(c->cfg.sample_kind, raw, nb, N, c->d_dcs, c->num_sms, stream));
      c->launches++;
    }
    SCN_CUDA(scn::launch_columns(c->cfg.sample_kind, raw, nb, N2, c->d_window, c->d_wn, dc ? c->d_dcs : nullptr,
                                 c->d_y, c->num_sms, stream));
    c->launches++;
    int rc = launch_rows(c, c->d_y, ns * 16, c->d_p, stream);
    if (rc != SCN_OK) return rc;
    SCN_CUDA(scn::launch_finalize(c->d_p, ns, N2, c->cfg.threshold, c->cfg.use_window, c->cfg.dc_ignore_window,
                                  d_spectra ? d_spectra + size_t(first) * N : nullptr,
                                  d_masks ? d_masks + size_t(first) * c->words : nullptr,
                                  d_counts ? d_counts + first : nullptr,
                                  d_hits ? d_hits + size_t(first) * c->hit_cap : nullptr, c->hit_cap, c->num_sms,
                                  stream));
    c->launches++;
  }
  SCN_CUDA(cudaEventRecord(c->large_done, stream));
  return SCN_OK;
}

int launch_frequency(scn_ctx* c, const void* d_raw, uint32_t n_spectra, float* d_spectra,
                     uint32_t* d_masks, uint32_t* d_counts, scn_hit* d_hits, cudaStream_t stream) {
  if (c->large) return launch_large(c, d_raw, n_spectra, d_spectra, d_masks, d_counts, d_hits, stream);
  scn::KernelParams p{};

  p.raw = static_cast<const uint8_t*>(d_raw);
  p.window = c->d_window;
  p.twiddles = c->d_twiddles;
  p.spectra = d_spectra;
  p.masks = d_masks;
  p.counts = d_counts;
  p.hits = d_hits;
  p.hit_cap = c->hit_cap;
  p.n_spectra = n_spectra;
  p.averaging = c->K;
  p.inv_averaging = 1.0f / float(c->K);
  p.threshold = c->cfg.threshold;
  p.use_window = c->cfg.use_window;
  p.dc_ignore = c->cfg.dc_ignore_window;
  p.win_mirror = c->win_mirror ? 1u : 0u;
  p.work = next_work_counter(c);
  void* args[] = {&p};
  if (c->variant.cluster > 0) {
    // persistent clusters: as many as fit on the device, each walks spectra cluster_id, + n_clusters, ...
    uint32_t n_clusters = uint32_t(c->max_clusters);
    if (n_clusters > n_spectra) n_clusters = n_spectra;
    if (n_clusters == 0) return SCN_OK;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(n_clusters * uint32_t(c->variant.cluster));
    lc.blockDim = dim3(unsigned(c->variant.threads));
    lc.dynamicSmemBytes = c->variant.smem_bytes;
    lc.stream = stream;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = unsigned(c->variant.cluster); at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    lc.attrs = &at; lc.numAttrs = 1;
    SCN_CUDA(cudaLaunchKernelExC(&lc, c->variant.func, args));
    c->launches++;
    return SCN_OK;
  }
  const uint32_t F = uint32_t(c->variant.transforms_per_cta);
  const uint32_t n_groups = (n_spectra + F - 1) / F;
  uint32_t grid = uint32_t(c->ctas_per_sm) * uint32_t(c->num_sms);
  if (grid > n_groups) grid = n_groups;
  if (grid == 0) return SCN_OK;
  SCN_CUDA(cudaLaunchKernel(c->variant.func, dim3(grid), dim3(c->variant.threads), args,
                            c->variant.smem_bytes, stream));
  c->launches++;
  return SCN_OK;
}

// Row mode of the fused kernel: n_rows independent N2-point transforms of fp32 complex rows, 16 rows per
// intermediate buffer, K-averaged power out.
int launch_rows(scn_ctx* c, const void* d_y, uint32_t n_rows, float* d_power, cudaStream_t stream) {
  scn::KernelParams p{};
  p.raw = static_cast<const uint8_t*>(d_y);
  p.window = c->d_ones;
  p.twiddles = c->d_twiddles;
  p.n_spectra = n_rows;
  p.averaging = c->K;
  p.inv_averaging = 1.0f / float(c->K);
  p.rpb_shift = 4;
  p.power_out = d_power;
  p.work = next_work_counter(c);
  const uint32_t F = uint32_t(c->variant.transforms_per_cta);
  const uint32_t n_groups = (n_rows + F - 1) / F;
  uint32_t grid = uint32_t(c->ctas_per_sm) * uint32_t(c->num_sms);
  if (grid > n_groups) grid = n_groups;
  if (grid == 0) return SCN_OK;
  void* args[] = {&p};
  SCN_CUDA(cudaLaunchKernel(c->variant.func, dim3(grid), dim3(c->variant.threads), args,
                            c->variant.smem_bytes, stream));
  c->launches++;
  return SCN_OK;
}

int launch_time_domain(scn_ctx* c, const void* d_raw, uint32_t n_buffers, uint32_t* d_trigger,
                       float* d_td, cudaStream_t stream) {
  if (n_buffers == 0) return SCN_OK;
  scn::TimeDomainParams p{};
  p.raw = static_cast<const uint8_t*>(d_raw);
  p.n_buffers = n_buffers;
  p.n = c->cfg.sample_count;
  p.threshold = c->cfg.threshold;
  p.trigger = d_trigger;
  p.max_min = d_td;
  p.scale = c->onebymax;
  p.correct_dc = c->cfg.correct_dc_offset;
  uint32_t grid = uint32_t(c->num_sms) * 8;
  if (grid > n_buffers) grid = n_buffers;
  cudaError_t e = scn::launch_time_domain_kernel(c->cfg.sample_kind, grid, stream, p);
  if (e != cudaSuccess) return fail(SCN_ERR_CUDA, "time-domain launch failed: %s", cudaGetErrorString(e));
  c->launches++;
  return SCN_OK;
}

void free_slot(Slot& s) {
  if (s.stream) cudaStreamDestroy(s.stream);
  if (s.done) cudaEventDestroy(s.done);
  if (s.h_raw) cudaFreeHost(s.h_raw);
  if (s.d_raw) cudaFree(s.d_raw);
  if (s.d_spectra) cudaFree(s.d_spectra);
  if (s.d_masks) cudaFree(s.d_masks);
  if (s.d_counts) cudaFree(s.d_counts);
  if (s.d_hits) cudaFree(s.d_hits);
  if (s.d_td) cudaFree(s.d_td);
  if (s.h_masks) cudaFreeHost(s.h_masks);
  if (s.h_counts) cudaFreeHost(s.h_counts);
  if (s.h_hits) cudaFreeHost(s.h_hits);
  if (s.h_td) cudaFreeHost(s.h_td);
  s = Slot{};
}

// Slot buffers are allocated on first use so device-resident-only callers pay nothing.
int ensure_slot_alloc(scn_ctx* c, Slot& s);
int ensure_slot(scn_ctx* c, Slot& s) {
  if (s.ready) return SCN_OK;
  const int rc = ensure_slot_alloc(c, s);
  if (rc != SCN_OK) { free_slot(s); return rc; }     // e.g. out of memory half way: leave nothing half-initialised
  s.ready = true;
  return SCN_OK;
}
int ensure_slot_alloc(scn_ctx* c, Slot& s) {
  const size_t S = c->cfg.max_spectra;
  SCN_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  SCN_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  const size_t n_buffers = S * c->K;
  SCN_CUDA(cudaMalloc(&s.d_raw, n_buffers * c->buf_bytes));
  SCN_CUDA(cudaMalloc(&s.d_counts, sizeof(uint32_t) * n_buffers));
  SCN_CUDA(cudaMallocHost(&s.h_counts, sizeof(uint32_t) * n_buffers));
  if (c->time_domain) {
    SCN_CUDA(cudaMalloc(&s.d_td, sizeof(float) * 2 * n_buffers));
    SCN_CUDA(cudaMallocHost(&s.h_td, sizeof(float) * 2 * n_buffers));
  } else {
    SCN_CUDA(cudaMalloc(&s.d_masks, sizeof(uint32_t) * S * c->words));
    SCN_CUDA(cudaMallocHost(&s.h_masks, sizeof(uint32_t) * S * c->words));
    if (c->cfg.flags & SCN_OUT_SPECTRUM)
      SCN_CUDA(cudaMalloc(&s.d_spectra, sizeof(float) * S * c->cfg.sample_count));
    if (c->cfg.flags & SCN_OUT_HITS) {
      SCN_CUDA(cudaMalloc(&s.d_hits, sizeof(scn_hit) * S * c->hit_cap));
      SCN_CUDA(cudaMallocHost(&s.h_hits, sizeof(scn_hit) * S * c->hit_cap));
    }
  }
  return SCN_OK;
}

bool is_device_accessible_host(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================

extern "C" {

SCN_API const char* scn_version(void) { return "scanner_b200 0.1 (sm_100a)"; }

SCN_API const char* scn_last_error(void) { return g_last_error.c_str(); }

SCN_API int scn_device_count(int* count) {
  if (!count) return fail(SCN_ERR_INVALID, "count is NULL");
  *count = 0;
  SCN_CUDA(cudaGetDeviceCount(count));
  return SCN_OK;
}

SCN_API int scn_create(const scn_config* config, scn_ctx** out) {
  if (!config || !out) return fail(SCN_ERR_INVALID, "config/out is NULL");
  *out = nullptr;
  const scn_config& cf = *config;
  const int log2n = ilog2_exact(cf.sample_count);
  const uint32_t mode = cf.mode ? cf.mode : uint32_t(SCN_MODE_FREQUENCY_DOMAIN);
  if (mode != SCN_MODE_FREQUENCY_DOMAIN && mode != SCN_MODE_TIME_DOMAIN)
    return fail(SCN_ERR_INVALID, "mode %u is not SCN_MODE_TIME_DOMAIN/FREQUENCY_DOMAIN", mode);
  const bool td = mode == SCN_MODE_TIME_DOMAIN;
  if (bytes_per_sample(cf.sample_kind) == 0)
    return fail(SCN_ERR_INVALID, "sample_kind %u is not a SampleKind", cf.sample_kind);
  if (!td && (log2n < scn::kMinLog2N || log2n > scn::kMaxLog2NLarge))
    return fail(SCN_ERR_INVALID, "sample_count %u unsupported: power of two in [%d, %d] required",
                cf.sample_count, 1 << scn::kMinLog2N, 1 << scn::kMaxLog2NLarge);
  if (td && (cf.sample_count == 0 || cf.sample_count % 8 != 0))
    return fail(SCN_ERR_INVALID, "sample_count %u must be a positive multiple of 8", cf.sample_count);
  if (cf.sample_kind != SCN_KIND_FLOAT_COMPLEX) {
    const uint32_t max_enob = cf.sample_kind == SCN_KIND_BYTE_COMPLEX ? 8 : 16;
    if (cf.enob < 1 || cf.enob > max_enob)
      return fail(SCN_ERR_INVALID, "enob %u out of range [1, %u] for sample_kind %u", cf.enob, max_enob,
                  cf.sample_kind);
  }
  if (!td && !cf.window) return fail(SCN_ERR_INVALID, "window table is NULL");
  if (cf.max_spectra == 0) return fail(SCN_ERR_INVALID, "max_spectra is 0");
  const uint32_t K = cf.averaging ? cf.averaging : 1;
  if (td && K != 1) return fail(SCN_ERR_INVALID, "averaging is a frequency-domain option");

  int ndev = 0;
  SCN_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) return fail(SCN_ERR_NO_DEVICE, "no CUDA device");
  if (cf.device < 0 || cf.device >= ndev)
    return fail(SCN_ERR_INVALID, "device %d out of range [0, %d)", cf.device, ndev);
  SCN_CUDA(cudaSetDevice(cf.device));

  scn_ctx* c = new scn_ctx();
  c->cfg = cf;
  c->cfg.mode = mode;
  c->cfg.window = nullptr;
  c->log2n = log2n;
  c->K = K;
  c->time_domain = td;
  c->buf_bytes = size_t(cf.sample_count) * bytes_per_sample(cf.sample_kind);
  c->words = cf.sample_count / 32;
  c->hit_cap = (cf.flags & SCN_OUT_HITS) ? (cf.max_hits_per_spectrum ? cf.max_hits_per_spectrum : cf.sample_count) : 0;
  // onebymax = float(1.0 / intK_t(1 << (enob-1))), utility.cpp:16-17,40-41,64-65 (wraps negative
  // for enob == 8 / 16).
  if (cf.sample_kind == SCN_KIND_BYTE_COMPLEX) {
    const int8_t mx = static_cast<int8_t>(1 << (cf.enob - 1));
    c->onebymax = float(1.0 / mx);
  } else if (cf.sample_kind != SCN_KIND_FLOAT_COMPLEX) {
    const int16_t mx = static_cast<int16_t>(1 << (cf.enob - 1));
    c->onebymax = float(1.0 / mx);
  }

  cudaDeviceProp prop{};
  cudaError_t e = cudaGetDeviceProperties(&prop, cf.device);
  if (e != cudaSuccess) { delete c; return fail(SCN_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
  c->num_sms = prop.multiProcessorCount;

  auto bail = [&](int code) { scn_destroy(c); return code; };

  if (!td) {
    // N >= 2^14: one transform per thread-block cluster, single HBM pass (scn_cluster.cu).  SCN_FOUR_STEP=1 in the
    // environment selects the older four-step path through an HBM intermediate (kept for A/B and as a cross-check).
    // (At N = 2^14 that older path is the 16-points-per-thread kernel; int8 IQ: 212 vs 278 Gsamples/s for the cluster
    // kernel, profiles/r02zh_*.txt.)
    const bool four_step = std::getenv("SCN_FOUR_STEP") != nullptr;
    const bool clustered = log2n >= 14 && !four_step &&
        scn::variant_cluster(int(cf.sample_kind), log2n, cf.correct_dc_offset != 0, K > 1, &c->variant);
    c->large = !clustered && log2n > scn::kMaxLog2N;
    c->log2n2 = c->large ? log2n - 4 : log2n;
    const bool found = clustered ? true : c->large
        ? scn::variant_float_rows(c->log2n2, K > 1, &c->variant)
        : scn::find_variant(int(cf.sample_kind), log2n, cf.correct_dc_offset != 0, K > 1, &c->variant);
    if (!found)
      return bail(fail(SCN_ERR_INVALID, "no kernel variant for kind %u, N %u", cf.sample_kind, cf.sample_count));
    e = cudaFuncSetAttribute(c->variant.func, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(c->variant.smem_bytes));
    if (e != cudaSuccess)
      return bail(fail(SCN_ERR_CUDA, "kernel image unavailable on this device (%s): built for sm_100a only",
                       cudaGetErrorString(e)));
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c->variant.func, c->variant.threads,
                                                      c->variant.smem_bytes);
    if (e != cudaSuccess || occ < 1)
      return bail(fail(SCN_ERR_CUDA, "occupancy query failed: %s", cudaGetErrorString(e)));
    c->ctas_per_sm = occ;
    if (c->variant.cluster > 0) {
      // how many clusters can be co-resident (GPC boundaries may strand a few SMs): the persistent grid
      c->ctas_per_sm = 1;
      cudaLaunchConfig_t lc{};
      lc.gridDim = dim3(unsigned(c->variant.cluster * c->num_sms));
      lc.blockDim = dim3(unsigned(c->variant.threads));
      lc.dynamicSmemBytes = c->variant.smem_bytes;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = unsigned(c->variant.cluster); at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      lc.attrs = &at; lc.numAttrs = 1;
      int n_clusters = 0;
      e = cudaOccupancyMaxActiveClusters(&n_clusters, c->variant.func, &lc);
      if (e != cudaSuccess || n_clusters < 1)
        return bail(fail(SCN_ERR_CUDA, "cluster occupancy query failed: %s", cudaGetErrorString(e)));
      c->max_clusters = n_clusters;
    }
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, c->variant.func) == cudaSuccess) c->regs = fa.numRegs;

    e = cudaMalloc(&c->d_work, sizeof(uint32_t) * kWorkCounters * kWorkStride);
    if (e == cudaSuccess) e = cudaMemset(c->d_work, 0, sizeof(uint32_t) * kWorkCounters * kWorkStride);
    if (e != cudaSuccess) return bail(fail(SCN_ERR_CUDA, "work counter allocation failed: %s", cudaGetErrorString(e)));
    // Window taps, with the converter's power-of-two scale folded in (exact).
    std::vector<float> w(cf.sample_count);
    const float s = (cf.sample_kind == SCN_KIND_FLOAT_COMPLEX) ? 1.0f : c->onebymax;
    for (uint32_t i = 0; i < cf.sample_count; i++) w[i] = cf.window[i] * s;
    c->win_mirror = true;
    for (uint32_t i = 0; i < cf.sample_count / 2 && c->win_mirror; i++)
      c->win_mirror = memcmp(&w[i], &w[cf.sample_count - 1 - i], sizeof(float)) == 0;
    e = cudaMalloc(&c->d_window, sizeof(float) * cf.sample_count);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_window, w.data(), sizeof(float) * cf.sample_count, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(fail(SCN_ERR_CUDA, "window upload failed: %s", cudaGetErrorString(e)));
    std::vector<float2> tw = build_twiddles(c->log2n2);
    if (c->variant.twiddle_layout == 2) {      // scn_p64.cuh: twA[(r-1)*64 + k] = W_4096^(k r); twB[m*64 + kk] = W_8192^(kk + 64 m)
      tw.assign(63 * 64 + 32 * 64, make_float2(0.f, 0.f));
      for (int r = 1; r < 64; r++)
        for (int k = 0; k < 64; k++) {
          const double a = -2.0 * kPi * double(k) * double(r) / 4096.0;
          tw[size_t(r - 1) * 64 + k] = make_float2(float(std::cos(a)), float(std::sin(a)));
        }
      for (int m = 0; m < 32; m++)
        for (int kk = 0; kk < 64; kk++) {
          const double a = -2.0 * kPi * double(kk + 64 * m) / 8192.0;
          tw[size_t(63 * 64) + size_t(m) * 64 + kk] = make_float2(float(std::cos(a)), float(std::sin(a)));
        }
    }
    if (c->variant.twiddle_layout == 3) {      // scn_cluster.cu: twA as above; twC[n1*64 + t] = W_N^(n1 t); twB[n1*64 + q] = W_N^(64 n1 q)
      const int R = int(cf.sample_count / 4096);
      tw.assign(size_t(63 * 64 + 2 * R * 64), make_float2(0.f, 0.f));
      for (int r = 1; r < 64; r++)
        for (int k = 0; k < 64; k++) {
          const double a = -2.0 * kPi * double(k) * double(r) / 4096.0;
          tw[size_t(r - 1) * 64 + k] = make_float2(float(std::cos(a)), float(std::sin(a)));
        }
      for (int n1 = 0; n1 < R; n1++)
        for (int x = 0; x < 64; x++) {
          const double a = -2.0 * kPi * double(n1) * double(x) / double(cf.sample_count);
          tw[size_t(63 * 64) + size_t(n1) * 64 + x] = make_float2(float(std::cos(a)), float(std::sin(a)));
          const double b = -2.0 * kPi * 64.0 * double(n1) * double(x) / double(cf.sample_count);
          tw[size_t(63 * 64 + R * 64) + size_t(n1) * 64 + x] = make_float2(float(std::cos(b)), float(std::sin(b)));
        }
    }
    if (c->variant.twiddle_layout == 1) {      // warp-per-transform kernel: exp(-2 pi i lane r / N), r = 1..63
      tw.assign(63 * 32, make_float2(0.f, 0.f));
      for (int r = 1; r < 64; r++)
        for (int lane = 0; lane < 32; lane++) {
          const double a = -2.0 * kPi * double(lane) * double(r) / double(cf.sample_count);
          tw[size_t(r - 1) * 32 + lane] = make_float2(float(std::cos(a)), float(std::sin(a)));
        }
    }
    e = cudaMalloc(&c->d_twiddles, sizeof(float2) * tw.size());
    if (e == cudaSuccess) e = cudaMemcpy(c->d_twiddles, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(fail(SCN_ERR_CUDA, "twiddle upload failed: %s", cudaGetErrorString(e)));
    if (c->large) {
      const uint32_t N = cf.sample_count, N2 = N / 16;
      std::vector<float> ones(N2, 1.0f);
      std::vector<float2> wn(N);
      for (uint32_t m = 0; m < N; m++) {
        const double a = -2.0 * kPi * double(m) / double(N);
        wn[m] = make_float2(float(std::cos(a)), float(std::sin(a)));
      }
      e = cudaMalloc(&c->d_ones, sizeof(float) * N2);
      if (e == cudaSuccess) e = cudaMemcpy(c->d_ones, ones.data(), sizeof(float) * N2, cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaMalloc(&c->d_wn, sizeof(float2) * N);
      if (e == cudaSuccess) e = cudaMemcpy(c->d_wn, wn.data(), sizeof(float2) * N, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) return bail(fail(SCN_ERR_CUDA, "four-step scratch allocation failed: %s", cudaGetErrorString(e)));
    }
  } else {
    c->variant.name = "time_domain_threshold";
    c->variant.threads = 256;
  }
  c->slots.resize(cf.ticket_slots ? cf.ticket_slots : 2);
  *out = c;
  return SCN_OK;
}

SCN_API int scn_destroy(scn_ctx* c) {
  if (!c) return SCN_OK;
  cudaSetDevice(c->cfg.device);
  for (auto& s : c->slots) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    free_slot(s);
  }
  if (c->d_work) cudaFree(c->d_work);
  if (c->d_window) cudaFree(c->d_window);
  if (c->d_twiddles) cudaFree(c->d_twiddles);
  if (c->d_ones) cudaFree(c->d_ones);
  if (c->d_wn) cudaFree(c->d_wn);
  if (c->d_y) cudaFree(c->d_y);
  if (c->d_p) cudaFree(c->d_p);
  if (c->d_dcs) cudaFree(c->d_dcs);
  if (c->large_done) cudaEventDestroy(c->large_done);
  if (c->d_conv_raw) cudaFree(c->d_conv_raw);
  if (c->d_conv_out) cudaFree(c->d_conv_out);
  delete c;
  return SCN_OK;
}

SCN_API int scn_set_threshold(scn_ctx* c, float threshold) {
  if (!c) return fail(SCN_ERR_INVALID, "ctx is NULL");
  c->cfg.threshold = threshold;
  return SCN_OK;
}

SCN_API size_t scn_buffer_bytes(const scn_ctx* c) { return c ? c->buf_bytes : 0; }
SCN_API uint32_t scn_mask_words(const scn_ctx* c) { return c ? c->words : 0; }
SCN_API uint64_t scn_launch_count(const scn_ctx* c) { return c ? c->launches : 0; }
SCN_API const char* scn_kernel_name(const scn_ctx* c) { return c && c->variant.name ? c->variant.name : ""; }

SCN_API int scn_kernel_info(const scn_ctx* c, int* ctas_per_sm, int* threads, int* smem_bytes,
                            int* regs_per_thread, int* grid) {
  if (!c) return fail(SCN_ERR_INVALID, "ctx is NULL");
  if (ctas_per_sm) *ctas_per_sm = c->ctas_per_sm;
  if (threads) *threads = c->variant.threads;
  if (smem_bytes) *smem_bytes = int(c->variant.smem_bytes);
  if (regs_per_thread) *regs_per_thread = c->regs;
  if (grid) *grid = c->ctas_per_sm * c->num_sms;
  return SCN_OK;
}

SCN_API int scn_alloc_pinned(size_t bytes, void** out) {
  if (!out) return fail(SCN_ERR_INVALID, "out is NULL");
  *out = nullptr;
  SCN_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return SCN_OK;
}

SCN_API int scn_free_pinned(void* p) {
  if (p) SCN_CUDA(cudaFreeHost(p));
  return SCN_OK;
}

SCN_API int scn_launch_device(scn_ctx* c, const void* d_raw, uint32_t n_spectra, float* d_spectra_db,
                              uint32_t* d_hit_mask, uint32_t* d_hit_count, scn_hit* d_hits,
                              float* d_td_max_min, void* stream) {
  if (!c) return fail(SCN_ERR_INVALID, "ctx is NULL");
  if (n_spectra == 0) return SCN_OK;
  if (!d_raw) return fail(SCN_ERR_INVALID, "d_raw is NULL");
  if (reinterpret_cast<uintptr_t>(d_raw) & 15) return fail(SCN_ERR_ALIGNMENT, "d_raw must be 16-byte aligned");
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  if (c->time_domain)
    return launch_time_domain(c, d_raw, n_spectra, d_hit_count, d_td_max_min, static_cast<cudaStream_t>(stream));
  if (d_hits && c->hit_cap == 0)
    return fail(SCN_ERR_INVALID, "d_hits given but the context was created without SCN_OUT_HITS");
  return launch_frequency(c, d_raw, n_spectra, d_spectra_db, d_hit_mask, d_hit_count, d_hits,
                          static_cast<cudaStream_t>(stream));
}

namespace {
// kernel + D2H of the detection records + completion event on the slot's stream, after its raw bytes were enqueued
int finish_submit(scn_ctx* c, Slot& s, uint32_t idx, uint32_t n_spectra, uint32_t* ticket) {
  int rc;
  if (c->time_domain) {
    rc = launch_time_domain(c, s.d_raw, n_spectra, s.d_counts, s.d_td, s.stream);
    if (rc != SCN_OK) return rc;
    SCN_CUDA(cudaMemcpyAsync(s.h_counts, s.d_counts, sizeof(uint32_t) * n_spectra, cudaMemcpyDeviceToHost, s.stream));
    SCN_CUDA(cudaMemcpyAsync(s.h_td, s.d_td, sizeof(float) * 2 * n_spectra, cudaMemcpyDeviceToHost, s.stream));
  } else {
    rc = launch_frequency(c, s.d_raw, n_spectra, s.d_spectra, s.d_masks, s.d_counts, s.d_hits, s.stream);
    if (rc != SCN_OK) return rc;
    SCN_CUDA(cudaMemcpyAsync(s.h_counts, s.d_counts, sizeof(uint32_t) * n_spectra, cudaMemcpyDeviceToHost, s.stream));
    SCN_CUDA(cudaMemcpyAsync(s.h_masks, s.d_masks, sizeof(uint32_t) * size_t(n_spectra) * c->words,
                             cudaMemcpyDeviceToHost, s.stream));
    if (s.d_hits)
      SCN_CUDA(cudaMemcpyAsync(s.h_hits, s.d_hits, sizeof(scn_hit) * size_t(n_spectra) * c->hit_cap,
                               cudaMemcpyDeviceToHost, s.stream));
  }
  SCN_CUDA(cudaEventRecord(s.done, s.stream));
  s.n_spectra = n_spectra;
  s.busy = true;
  *ticket = idx;
  c->next_slot = (idx + 1) % uint32_t(c->slots.size());
  return SCN_OK;
}
}  // namespace

SCN_API int scn_submit(scn_ctx* c, const void* raw, uint32_t n_spectra, uint32_t* ticket) {
  if (!c || !ticket) return fail(SCN_ERR_INVALID, "ctx/ticket is NULL");
  if (n_spectra == 0 || !raw) return fail(SCN_ERR_INVALID, "empty submit");
  if (n_spectra > c->cfg.max_spectra)
    return fail(SCN_ERR_CAPACITY, "n_spectra %u exceeds max_spectra %u", n_spectra, c->cfg.max_spectra);
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  const uint32_t idx = c->next_slot;
  Slot& s = c->slots[idx];
  if (s.busy) return fail(SCN_ERR_BUSY, "ticket slot %u has not been collected", idx);
  int rc = ensure_slot(c, s);
  if (rc != SCN_OK) return rc;
  const size_t bytes = size_t(n_spectra) * c->K * c->buf_bytes;
  const void* src = raw;
  if (!is_device_accessible_host(raw)) {
    // pageable caller memory: stage through pinned memory so the copy is a real async DMA
    if (!s.h_raw) SCN_CUDA(cudaMallocHost(&s.h_raw, size_t(c->cfg.max_spectra) * c->K * c->buf_bytes));
    std::memcpy(s.h_raw, raw, bytes);
    src = s.h_raw;
  }
  SCN_CUDA(cudaMemcpyAsync(s.d_raw, src, bytes, cudaMemcpyHostToDevice, s.stream));
  return finish_submit(c, s, idx, n_spectra, ticket);
}

SCN_API int scn_submit_gather(scn_ctx* c, const void* const* runs, const uint32_t* run_buffers, uint32_t n_runs,
                              uint32_t n_spectra, uint32_t* ticket) {
  if (!c || !ticket || !runs || !run_buffers) return fail(SCN_ERR_INVALID, "submit_gather: NULL argument");
  if (n_spectra == 0 || n_runs == 0) return fail(SCN_ERR_INVALID, "empty submit");
  if (n_spectra > c->cfg.max_spectra)
    return fail(SCN_ERR_CAPACITY, "n_spectra %u exceeds max_spectra %u", n_spectra, c->cfg.max_spectra);
  uint64_t total = 0;
  for (uint32_t r = 0; r < n_runs; r++) {
    if (!runs[r] || run_buffers[r] == 0) return fail(SCN_ERR_INVALID, "submit_gather: run %u is empty", r);
    total += run_buffers[r];
  }
  if (total != uint64_t(n_spectra) * c->K)
    return fail(SCN_ERR_INVALID, "submit_gather: runs hold %llu buffers, n_spectra * averaging is %llu",
                (unsigned long long)total, (unsigned long long)(uint64_t(n_spectra) * c->K));
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  const uint32_t idx = c->next_slot;
  Slot& s = c->slots[idx];
  if (s.busy) return fail(SCN_ERR_BUSY, "ticket slot %u has not been collected", idx);
  int rc = ensure_slot(c, s);
  if (rc != SCN_OK) return rc;
  // one H2D copy per run, back to back on the slot's stream, into consecutive device addresses; pageable runs
  // are packed into the slot's pinned staging buffer first (then it is ONE copy)
  bool pinned = true;
  for (uint32_t r = 0; r < n_runs && pinned; r++) pinned = is_device_accessible_host(runs[r]);
  size_t off = 0;
  if (!pinned) {
    if (!s.h_raw) SCN_CUDA(cudaMallocHost(&s.h_raw, size_t(c->cfg.max_spectra) * c->K * c->buf_bytes));
    for (uint32_t r = 0; r < n_runs; r++) {
      const size_t bytes = size_t(run_buffers[r]) * c->buf_bytes;
      std::memcpy(static_cast<char*>(s.h_raw) + off, runs[r], bytes);
      off += bytes;
    }
    SCN_CUDA(cudaMemcpyAsync(s.d_raw, s.h_raw, off, cudaMemcpyHostToDevice, s.stream));
  } else {
    for (uint32_t r = 0; r < n_runs; r++) {
      const size_t bytes = size_t(run_buffers[r]) * c->buf_bytes;
      SCN_CUDA(cudaMemcpyAsync(static_cast<char*>(s.d_raw) + off, runs[r], bytes, cudaMemcpyHostToDevice, s.stream));
      off += bytes;
    }
  }
  return finish_submit(c, s, idx, n_spectra, ticket);
}

SCN_API int scn_collect(scn_ctx* c, uint32_t ticket, float* spectra_db, uint32_t* hit_mask,
                        uint32_t* hit_count, scn_hit* hits, float* td_max_min) {
  if (!c) return fail(SCN_ERR_INVALID, "ctx is NULL");
  if (ticket >= c->slots.size() || !c->slots[ticket].busy)
    return fail(SCN_ERR_INVALID, "ticket %u is not in flight", ticket);
  Slot& s = c->slots[ticket];
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  SCN_CUDA(cudaEventSynchronize(s.done));
  const size_t n = s.n_spectra;
  if (hit_count) std::memcpy(hit_count, s.h_counts, sizeof(uint32_t) * n);
  if (c->time_domain) {
    if (td_max_min) std::memcpy(td_max_min, s.h_td, sizeof(float) * 2 * n);
  } else {
    if (hit_mask) std::memcpy(hit_mask, s.h_masks, sizeof(uint32_t) * n * c->words);
    if (hits) {
      if (!s.h_hits) { s.busy = false; return fail(SCN_ERR_INVALID, "hits requested without SCN_OUT_HITS"); }
      std::memcpy(hits, s.h_hits, sizeof(scn_hit) * n * c->hit_cap);
    }
    if (spectra_db) {
      if (!s.d_spectra) { s.busy = false; return fail(SCN_ERR_INVALID, "spectra requested without SCN_OUT_SPECTRUM"); }
      // The spectrum stays resident in HBM; it crosses PCIe only when asked for.
      SCN_CUDA(cudaMemcpyAsync(spectra_db, s.d_spectra, sizeof(float) * n * c->cfg.sample_count,
                               cudaMemcpyDeviceToHost, s.stream));
      SCN_CUDA(cudaStreamSynchronize(s.stream));
    }
  }
  s.busy = false;
  return SCN_OK;
}

SCN_API int scn_collect_view(scn_ctx* c, uint32_t ticket, const uint32_t** hit_mask, const uint32_t** hit_count,
                             const scn_hit** hits, const float** td_max_min) {
  if (!c) return fail(SCN_ERR_INVALID, "ctx is NULL");
  if (ticket >= c->slots.size() || !c->slots[ticket].busy)
    return fail(SCN_ERR_INVALID, "ticket %u is not in flight", ticket);
  Slot& s = c->slots[ticket];
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  SCN_CUDA(cudaEventSynchronize(s.done));
  if (hit_mask) *hit_mask = s.h_masks;
  if (hit_count) *hit_count = s.h_counts;
  if (hits) *hits = s.h_hits;
  if (td_max_min) *td_max_min = s.h_td;
  s.busy = false;
  return SCN_OK;
}

SCN_API int scn_process_host(scn_ctx* c, const void* raw, uint32_t n_spectra, float* spectra_db,
                             uint32_t* hit_mask, uint32_t* hit_count, scn_hit* hits, float* td_max_min) {
  if (!c) return fail(SCN_ERR_INVALID, "ctx is NULL");
  if (n_spectra == 0) return SCN_OK;
  // Batches larger than the context capacity are cut into capacity-sized submits that overlap
  // (copy of chunk i+1 against the kernel of chunk i) across the ticket slots.
  const uint32_t cap = c->cfg.max_spectra;
  const size_t N = c->cfg.sample_count;
  const size_t chunk_bytes_per_spec = size_t(c->K) * c->buf_bytes;
  struct Pending { uint32_t ticket, first, count; };
  std::vector<Pending> inflight;
  auto collect_front = [&]() -> int {
    Pending pd = inflight.front();
    inflight.erase(inflight.begin());
    return scn_collect(c, pd.ticket, spectra_db ? spectra_db + size_t(pd.first) * N : nullptr,
                       hit_mask ? hit_mask + size_t(pd.first) * c->words : nullptr,
                       hit_count ? hit_count + pd.first : nullptr,
                       hits ? hits + size_t(pd.first) * c->hit_cap : nullptr,
                       td_max_min ? td_max_min + size_t(pd.first) * 2 : nullptr);
  };
  // on any failure the tickets still in flight are collected (results discarded) so the context stays usable
  auto drain = [&](int rc) {
    const std::string msg = g_last_error;
    for (const Pending& pd : inflight) scn_collect(c, pd.ticket, nullptr, nullptr, nullptr, nullptr, nullptr);
    inflight.clear();
    g_last_error = msg;
    return rc;
  };
  for (uint32_t first = 0; first < n_spectra; first += cap) {
    const uint32_t count = (n_spectra - first < cap) ? (n_spectra - first) : cap;
    if (inflight.size() == c->slots.size()) {
      int rc = collect_front();
      if (rc != SCN_OK) return drain(rc);
    }
    uint32_t ticket = 0;
    int rc = scn_submit(c, static_cast<const uint8_t*>(raw) + size_t(first) * chunk_bytes_per_spec, count, &ticket);
    if (rc != SCN_OK) return drain(rc);
    inflight.push_back({ticket, first, count});
  }
  while (!inflight.empty()) {
    int rc = collect_front();
    if (rc != SCN_OK) return drain(rc);
  }
  return SCN_OK;
}

SCN_API uint32_t scn_record_words(const scn_ctx* c) { return c ? c->words + 2 : 0; }

SCN_API int scn_summarize_steps(scn_ctx* c, const uint32_t* d_hit_mask, const uint32_t* d_hit_count,
                                uint32_t n_spectra, uint64_t first_unit, uint32_t units_per_step,
                                uint32_t n_steps, uint32_t* d_records, void* stream) {
  if (!c || !d_records || units_per_step == 0 || n_steps == 0)
    return fail(SCN_ERR_INVALID, "summarize_steps: bad arguments");
  if (n_spectra && (!d_hit_mask || !d_hit_count)) return fail(SCN_ERR_INVALID, "summarize_steps: NULL inputs");
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  SCN_CUDA(scn::launch_summarize(d_hit_mask, d_hit_count, n_spectra, first_unit, units_per_step, n_steps,
                                 c->words, d_records, c->num_sms, static_cast<cudaStream_t>(stream)));
  if (n_spectra) c->launches++;
  return SCN_OK;
}

SCN_API int scn_merge_step_records(scn_ctx* c, const uint32_t* d_parts, uint32_t n_parts, uint32_t n_steps,
                                   uint32_t* d_out, void* stream) {
  if (!c || !d_parts || !d_out || n_parts == 0) return fail(SCN_ERR_INVALID, "merge_step_records: bad arguments");
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  SCN_CUDA(scn::launch_merge(d_parts, n_parts, n_steps, c->words + 2, d_out, static_cast<cudaStream_t>(stream)));
  c->launches++;
  return SCN_OK;
}

SCN_API int scn_convert_device(scn_ctx* c, const void* d_raw, uint32_t n_buffers, float* d_out, void* stream) {
  if (!c || (n_buffers && (!d_raw || !d_out))) return fail(SCN_ERR_INVALID, "convert: bad arguments");
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  SCN_CUDA(scn::launch_convert(int(c->cfg.sample_kind), d_raw, d_out, c->cfg.sample_count, n_buffers, c->onebymax,
                               c->cfg.correct_dc_offset != 0, c->num_sms, static_cast<cudaStream_t>(stream)));
  if (n_buffers) c->launches++;
  return SCN_OK;
}

SCN_API int scn_convert_host(scn_ctx* c, const void* raw, uint32_t n_buffers, float* out) {
  if (!c || (n_buffers && (!raw || !out))) return fail(SCN_ERR_INVALID, "convert: bad arguments");
  if (n_buffers == 0) return SCN_OK;
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  const size_t N = c->cfg.sample_count;
  if (n_buffers > c->conv_capacity) {
    if (c->d_conv_raw) cudaFree(c->d_conv_raw);
    if (c->d_conv_out) cudaFree(c->d_conv_out);
    c->d_conv_raw = nullptr; c->d_conv_out = nullptr; c->conv_capacity = 0;
    SCN_CUDA(cudaMalloc(&c->d_conv_raw, size_t(n_buffers) * c->buf_bytes));
    SCN_CUDA(cudaMalloc(&c->d_conv_out, size_t(n_buffers) * N * 2 * sizeof(float)));
    c->conv_capacity = n_buffers;
  }
  SCN_CUDA(cudaMemcpy(c->d_conv_raw, raw, size_t(n_buffers) * c->buf_bytes, cudaMemcpyHostToDevice));
  int rc = scn_convert_device(c, c->d_conv_raw, n_buffers, c->d_conv_out, nullptr);
  if (rc != SCN_OK) return rc;
  SCN_CUDA(cudaMemcpy(out, c->d_conv_out, size_t(n_buffers) * N * 2 * sizeof(float), cudaMemcpyDeviceToHost));
  return SCN_OK;
}

SCN_API int scn_hackrf_prepass_device(scn_ctx* c, void* d_transfers, uint32_t n_transfers, uint32_t valid_length,
                                      uint64_t* d_frequency_hz, uint32_t* d_status, void* stream) {
  if (!c || (n_transfers && !d_transfers)) return fail(SCN_ERR_INVALID, "hackrf_prepass: bad arguments");
  if (c->cfg.sample_kind != SCN_KIND_BYTE_COMPLEX)
    return fail(SCN_ERR_INVALID, "hackrf_prepass: HackRF transfers are int8 IQ (SCN_KIND_BYTE_COMPLEX)");
  if (valid_length < 12 || valid_length % (2 * c->cfg.sample_count) != 0)
    return fail(SCN_ERR_INVALID, "hackrf_prepass: valid_length must be a whole number of sample_count buffers");
  SCN_CUDA(cudaSetDevice(c->cfg.device));
  SCN_CUDA(scn::launch_hackrf_prepass(d_transfers, n_transfers, valid_length, d_frequency_hz, d_status,
                                      static_cast<cudaStream_t>(stream)));
  if (n_transfers) c->launches++;
  return SCN_OK;
}

// ---- host helpers that restate reference arithmetic -------------------------------------------

SCN_API uint32_t scn_use_window(double use_bandwidth, uint32_t sample_count) {
  return uint32_t(use_bandwidth * sample_count / 2.0);   // process.cpp:85
}

SCN_API uint64_t scn_hit_frequency(double center_frequency, uint32_t sample_rate, uint32_t sample_count,
                                   uint32_t bin) {
  // process.cpp:38: double start_frequency = header->m_frequency - this->m_sampleRate/2;
  // process.cpp:39: uint32_t bin_step = this->m_sampleRate/this->m_sampleCount;
  // process.cpp:55: double frequency = start_frequency + i*bin_step;   (uint32 product)
  // process.cpp:57: printf("freq %lu ...", uint64_t(frequency), ...)
  const double start_frequency = center_frequency - double(sample_rate / 2u);
  const uint32_t bin_step = sample_count ? sample_rate / sample_count : 0;
  const uint32_t offset = bin * bin_step;
  return uint64_t(start_frequency + double(offset));
}

SCN_API uint32_t scn_frequency_table(uint32_t sample_rate, double start_frequency, double stop_frequency,
                                     double use_bandwidth, double dc_ignore_width, double* out, uint32_t cap) {
  // frequencyTable.cpp:17-36
  const double f1 = start_frequency + use_bandwidth / 2 * sample_rate;
  double step = use_bandwidth;
  if (dc_ignore_width > 0) step = (use_bandwidth - dc_ignore_width) / 2;
  uint32_t count = 0;
  if (stop_frequency == 0.0) {
    count = 1;
  } else {
    while (f1 + count * step * double(sample_rate) < stop_frequency) count++;
  }
  if (out)
    for (uint32_t i = 0; i < count && i < cap; i++) out[i] = f1 + i * step * double(sample_rate);
  return count;
}

SCN_API int scn_window_build(int win_type, uint32_t n, float* out) {
  // gr::fft::window::build(type, N, 0.0) as called at process.cpp:18: symmetric, M = N-1.
  if (!out || n == 0) return fail(SCN_ERR_INVALID, "window_build: bad arguments");
  const double M = double(n) - 1.0;
  for (uint32_t i = 0; i < n; i++) {
    const double x = n > 1 ? double(i) / M : 0.0;
    double w;
    switch (win_type) {
      case SCN_WIN_HAMMING: w = 0.54 - 0.46 * std::cos(2 * kPi * x); break;
      case SCN_WIN_HANN: w = 0.5 - 0.5 * std::cos(2 * kPi * x); break;
      case SCN_WIN_BLACKMAN: w = 0.42 - 0.5 * std::cos(2 * kPi * x) + 0.08 * std::cos(4 * kPi * x); break;
      case SCN_WIN_RECTANGULAR: w = 1.0; break;
      case SCN_WIN_BLACKMAN_HARRIS:
        w = 0.35875 - 0.48829 * std::cos(2 * kPi * x) + 0.14128 * std::cos(4 * kPi * x) -
            0.01168 * std::cos(6 * kPi * x);
        break;
      default: return fail(SCN_ERR_INVALID, "unsupported window type %d", win_type);
    }
    out[i] = float(w);
  }
  return SCN_OK;
}

SCN_API void scn_shard_steps(uint32_t n_steps, uint32_t rank, uint32_t world, uint32_t* begin, uint32_t* end) {
  if (world == 0) world = 1;
  if (begin) *begin = uint32_t(uint64_t(n_steps) * rank / world);
  if (end) *end = uint32_t(uint64_t(n_steps) * (rank + 1) / world);
}

}  // extern "C"
