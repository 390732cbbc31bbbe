/* scanner_b200.h -- C ABI of the B200-native spectrum-sense hot path.
 *
 * This is the drop-in boundary for wpats/scanner's per-retune processing:
 *   convert (utility.cpp:9-84) -> window (process.cpp:28-34) -> FFT (fft.cpp:20-25)
 *   -> dB (utility.cpp:86-98) -> [K-FFT averaging, extension] -> detect (process.cpp:36-64).
 * One call processes a BATCH of raw IQ buffers on one GPU with a fused sm_100a kernel;
 * there is no CPU fallback: every entry point returns SCN_ERR_CUDA / SCN_ERR_NO_DEVICE
 * when the device path is unavailable.
 *
 * Plain C types only.  No exceptions cross this boundary; every function returns an
 * int status (0 == SCN_OK) the way the reference's SDK wrappers do.  A context is
 * thread-compatible: use one context per consumer thread (the reference runs two
 * ProcessSamples::ThreadWorker threads, scan.cpp:217).
 *
 * Reference interfaces replaced (file:line under the reference tree):
 *   scn_config          <- ProcessSamples ctor arguments, process.h:74-85 / process.cpp:66-108,
 *                          plus the SampleQueue ctor's (kind, enob, correctDCOffset),
 *                          messageQueue.h:141-146
 *   scn_process_host    <- ProcessSamples::Run, process.cpp:131-144, and the body of
 *                          ProcessSamples::ThreadWorker's FrequencyDomain branch,
 *                          process.cpp:292-299, applied to a batch of queue messages
 *   scn_submit/collect  <- the same, asynchronous (tickets), for the ThreadWorker loop
 *   scn_submit_gather / scn_collect_view
 *                       <- the same without host copies: the batch is handed over as address runs of the queue's
 *                          pinned slab (what replaces the per-message memcpy of process.cpp:293-295) and the results
 *                          are read in place
 *   scn_exchange_* / scn_nccl_gather_*
 *                       <- no reference counterpart (process.cpp:316-331 runs two threads on one FFT): the per-sweep
 *                          exchange of per-retune-step records between the GPUs that share a sweep
 *   scn_launch_device   <- the same on buffers already resident in HBM
 *   scn_hit_frequency   <- the Hz mapping inside process_fft, process.cpp:38-39,55
 *   scn_use_window      <- m_useWindow initialiser, process.cpp:85
 *   scn_frequency_table <- FrequencyTable::FrequencyTable, frequencyTable.cpp:9-37
 *   scn_window_build    <- FFTWindow::FFTWindow, process.cpp:14-21 (gr::fft::window::build)
 *   scn_time_domain_*   <- ProcessSamples::DoTimeDomainThresholding, process.cpp:203-237
 */
#ifndef SCANNER_B200_H_
#define SCANNER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SCN_API __attribute__((visibility("default")))
#else
#define SCN_API
#endif

/* Status codes. */
enum {
  SCN_OK = 0,
  SCN_ERR_INVALID = 1,     /* bad argument / unsupported configuration */
  SCN_ERR_NO_DEVICE = 2,   /* no CUDA device: there is deliberately no CPU fallback */
  SCN_ERR_CUDA = 3,        /* CUDA runtime error; see scn_last_error() */
  SCN_ERR_CAPACITY = 4,    /* batch larger than the context was created for */
  SCN_ERR_BUSY = 5,        /* no free ticket slot */
  SCN_ERR_ALIGNMENT = 6    /* device pointer not 16-byte aligned */
};

/* Sample kinds -- same numbering as SampleQueue::SampleKind, messageQueue.h:31-37. */
enum {
  SCN_KIND_BYTE_COMPLEX = 1,   /* int8_t  [N][2] interleaved I,Q        (messageQueue.h:218) */
  SCN_KIND_SHORT = 2,          /* int16_t re[N] followed by int16_t im[N] (messageQueue.h:190) */
  SCN_KIND_SHORT_COMPLEX = 3,  /* int16_t [N][2] interleaved I,Q        (messageQueue.h:205) */
  SCN_KIND_FLOAT_COMPLEX = 4   /* float   [N][2] interleaved I,Q        (messageQueue.h:231) */
};

/* Window types -- numbering of gr::fft::window::win_type as used at scan.cpp:215. */
enum {
  SCN_WIN_HAMMING = 0,
  SCN_WIN_HANN = 1,
  SCN_WIN_BLACKMAN = 2,
  SCN_WIN_RECTANGULAR = 3,
  SCN_WIN_BLACKMAN_HARRIS = 5
};

/* Processing mode -- ProcessSamples::Mode, process.h:24-28. */
enum {
  SCN_MODE_TIME_DOMAIN = 1,
  SCN_MODE_FREQUENCY_DOMAIN = 2
};

/* Output selection flags (scn_config.flags). */
enum {
  SCN_OUT_SPECTRUM = 1u << 0,  /* keep/write the averaged dB spectrum, N floats per spectrum */
  SCN_OUT_HITS = 1u << 1       /* write compact {bin, power_db} records per hit */
};

/* One detection: the fields the reference prints at process.cpp:57
 * ("freq %lu power_db %f"), with the frequency kept as the shifted bin index i
 * (ascending i == ascending frequency == the reference's print order);
 * scn_hit_frequency() turns (center frequency, i) into the printed Hz. */
typedef struct scn_hit {
  uint32_t bin;     /* shifted index i of process.cpp:46-47 (0 == lowest frequency) */
  float power_db;   /* magnitudes[j], utility.cpp:97 */
} scn_hit;

typedef struct scn_config {
  int32_t device;               /* CUDA ordinal */
  uint32_t sample_count;        /* N: FFT size == samples per buffer (process.cpp:78); power of two, 256..65536 (2^14..2^16: one transform per thread-block cluster, single HBM pass) */
  uint32_t sample_rate;         /* Hz (process.cpp:79) */
  uint32_t enob;                /* effective number of bits (scan.cpp:138,183,196) */
  uint32_t sample_kind;         /* SCN_KIND_* */
  uint32_t correct_dc_offset;   /* 0/1 (scan.cpp:149,184) */
  uint32_t averaging;           /* K consecutive buffers averaged into one spectrum; 1 == reference behaviour */
  uint32_t mode;                /* SCN_MODE_*; 0 == frequency domain */
  float threshold;              /* dB threshold (process.cpp:82) */
  uint32_t use_window;          /* m_useWindow (process.cpp:85); see scn_use_window() */
  uint32_t dc_ignore_window;    /* m_dcIgnoreWindow (process.cpp:87): the reference hard-codes 4 */
  const float* window;          /* N taps (process.cpp:19-20); copied at create time; must not be NULL in frequency mode */
  uint32_t max_spectra;         /* capacity of one submit, in spectra (== buffers when averaging == 1) */
  uint32_t max_hits_per_spectrum; /* capacity of the per-spectrum hit record list (0 with SCN_OUT_HITS == N) */
  uint32_t flags;               /* SCN_OUT_* */
  uint32_t ticket_slots;        /* number of in-flight submits (0 == 2) */
} scn_config;

typedef struct scn_ctx scn_ctx;

/* Library / device information. */
SCN_API const char* scn_version(void);
SCN_API int scn_device_count(int* count);
SCN_API const char* scn_last_error(void);   /* thread-local text of the last failure */

/* Context lifetime. */
SCN_API int scn_create(const scn_config* config, scn_ctx** out);
SCN_API int scn_destroy(scn_ctx* ctx);
SCN_API int scn_set_threshold(scn_ctx* ctx, float threshold);

/* Bytes of one raw buffer (N samples) of the context's sample kind. */
SCN_API size_t scn_buffer_bytes(const scn_ctx* ctx);
/* Mask words (uint32) per spectrum: N/32. */
SCN_API uint32_t scn_mask_words(const scn_ctx* ctx);

/* Pinned host memory for callers that want zero-copy staging of raw buffers. */
SCN_API int scn_alloc_pinned(size_t bytes, void** out);
SCN_API int scn_free_pinned(void* p);

/* ---- Synchronous host path -------------------------------------------------
 * raw: n_spectra * averaging buffers, contiguous, caller owned, pageable or pinned.
 * Outputs (each nullable): spectra_db [n_spectra][N] in FFT bin order j (the reference's
 * magnitudes[j]); hit_mask [n_spectra][N/32], bit i of the mask == shifted index i hit;
 * hit_count [n_spectra] (the reference's triggerCount); hits [n_spectra][max_hits_per_spectrum]
 * ascending in bin, first min(count, cap) entries valid.
 * Time-domain mode: spectra_db / hit_mask / hits are ignored; hit_count[b] is 1 when buffer b
 * triggers (max magnitude dB >= threshold, process.cpp:226) and td_max_min (nullable,
 * [n_buffers][2]) receives {maxMagnitude, minMagnitude}. */
SCN_API int scn_process_host(scn_ctx* ctx, const void* raw, uint32_t n_spectra, float* spectra_db,
                             uint32_t* hit_mask, uint32_t* hit_count, scn_hit* hits,
                             float* td_max_min);

/* ---- Asynchronous host path (tickets) --------------------------------------
 * scn_submit copies/stages `raw`, enqueues H2D + kernel + D2H of the detection records on
 * the ticket's stream and returns immediately.  scn_collect blocks until that work is done
 * and copies the results out.  Tickets complete in submit order. */
SCN_API int scn_submit(scn_ctx* ctx, const void* raw, uint32_t n_spectra, uint32_t* ticket);
/* scn_submit for a batch that lies in several pieces: runs[r] points at run_buffers[r] consecutive raw buffers
 * (sum == n_spectra * averaging, in batch order).  One H2D copy per run, straight from the caller's memory when it
 * is pinned (SampleQueue's slab: the consumer then never copies a sample on the host); pageable runs are packed
 * through the slot's pinned staging buffer. */
SCN_API int scn_submit_gather(scn_ctx* ctx, const void* const* runs, const uint32_t* run_buffers, uint32_t n_runs,
                              uint32_t n_spectra, uint32_t* ticket);
SCN_API int scn_collect(scn_ctx* ctx, uint32_t ticket, float* spectra_db, uint32_t* hit_mask,
                        uint32_t* hit_count, scn_hit* hits, float* td_max_min);

/* scn_collect without the copies: waits for the ticket and returns POINTERS into the slot's pinned host buffers
 * (layouts as above; NULL for outputs the context does not produce).  They stay valid until the slot is reused, i.e.
 * until the ticket_slots-th scn_submit after the one that returned `ticket`. */
SCN_API int scn_collect_view(scn_ctx* ctx, uint32_t ticket, const uint32_t** hit_mask, const uint32_t** hit_count,
                             const scn_hit** hits, const float** td_max_min);

/* ---- Device-resident path ---------------------------------------------------
 * All pointers are device pointers on ctx's device (16-byte aligned raw); outputs nullable as
 * above.  Enqueues exactly one fused kernel on `stream` (a cudaStream_t, NULL == default
 * stream) and returns without synchronising. */
SCN_API int scn_launch_device(scn_ctx* ctx, const void* d_raw, uint32_t n_spectra,
                              float* d_spectra_db, uint32_t* d_hit_mask, uint32_t* d_hit_count,
                              scn_hit* d_hits, float* d_td_max_min, void* stream);

/* Number of kernels this context has launched since creation. */
SCN_API uint64_t scn_launch_count(const scn_ctx* ctx);
/* Name of the kernel variant the context dispatches to (for logs / profiles). */
SCN_API const char* scn_kernel_name(const scn_ctx* ctx);
/* Occupancy facts of that variant: resident CTAs per SM, threads per CTA, dynamic smem bytes. */
SCN_API int scn_kernel_info(const scn_ctx* ctx, int* ctas_per_sm, int* threads, int* smem_bytes,
                            int* regs_per_thread, int* grid);

/* ---- Per-retune-step records (device resident; what ranks exchange over NCCL) -------------
 * Spectra are numbered globally in step-major order: spectrum u belongs to retune step
 * u / units_per_step of the FrequencyTable (frequencyTable.cpp:17-36).  This context's batch
 * holds units [first_unit, first_unit + n_spectra).  d_records receives, for each of n_steps
 * steps, scn_record_words() uint32: [0] hit total, [1] spectra contributing, [2..] OR of the
 * hit masks.  Steps this batch does not touch get zeros, so per-rank partial records merge by
 * sum/OR (scn_merge_step_records) after an all-gather. */
SCN_API uint32_t scn_record_words(const scn_ctx* ctx);
SCN_API int scn_summarize_steps(scn_ctx* ctx, const uint32_t* d_hit_mask, const uint32_t* d_hit_count,
                                uint32_t n_spectra, uint64_t first_unit, uint32_t units_per_step,
                                uint32_t n_steps, uint32_t* d_records, void* stream);
SCN_API int scn_merge_step_records(scn_ctx* ctx, const uint32_t* d_parts, uint32_t n_parts,
                                   uint32_t n_steps, uint32_t* d_out, void* stream);

/* ---- Record exchange between the GPUs of one box over NVLink peer memory (scn_exchange.cu) ----------------
 * The alternative to all-gathering scn_summarize_steps' partial records with NCCL: every rank owns a window in
 * its HBM that its peers write directly (peer-mapped stores + a flag), so no rank ever waits at a rendezvous
 * and no collective kernel competes with the persistent fused kernel for an SM.
 *   create  -> one window per rank (device = CUDA ordinal; record_words = scn_record_words())
 *   handle / connect_ipc   : ranks in different processes swap the 64-byte CUDA IPC handles (bench.py does it
 *                            with torch.distributed) and map each other's windows
 *   connect_local          : ranks that are devices of ONE process (csrc/host/sweepProcessor.cpp)
 *   publish(d_records)     : stores this rank's [n_steps][record_words] partial records into every peer's
 *                            window; returns the sequence number (1, 2, ...) of the batch
 *   merge(seq, d_merged)   : waits (device side, bounded ~10 s) until every rank has published `seq`, then
 *                            writes the merged records (sum of words 0,1; OR of the mask words)
 * Contract: every rank merges every sequence number and issues merge(s) before publish(s + 2) on the same
 * stream (publish(i) then merge(i - 1) per batch is the intended use: ranks never wait for each other).
 * scn_exchange_status reports a merge that gave up waiting (0 = none). */
#define SCN_IPC_HANDLE_BYTES 64
typedef struct scn_exchange scn_exchange;
SCN_API int scn_exchange_create(int device, uint32_t rank, uint32_t world, uint32_t n_steps,
                                uint32_t record_words, scn_exchange** out);
SCN_API int scn_exchange_handle(scn_exchange* x, unsigned char* handle /* [SCN_IPC_HANDLE_BYTES] */);
SCN_API int scn_exchange_connect_ipc(scn_exchange* x, const unsigned char* handles /* [world][SCN_IPC_HANDLE_BYTES] */);
SCN_API int scn_exchange_connect_local(scn_exchange* const* all /* [world], index == rank */, uint32_t world);
SCN_API int scn_exchange_publish(scn_exchange* x, const uint32_t* d_records, void* stream, uint64_t* seq_out);
SCN_API int scn_exchange_merge(scn_exchange* x, uint64_t seq, uint32_t* d_merged, void* stream);
/* publish(d_records) and merge(previous sequence number) in ONE kernel launch -- the steady-state call of a batch loop
 * (the first call merges nothing; after the last batch one scn_exchange_merge(last seq) closes the loop). */
SCN_API int scn_exchange_step(scn_exchange* x, const uint32_t* d_records, uint32_t* d_merged_previous, void* stream,
                              uint64_t* seq_out);
/* host-pointer forms (synchronous): upload + publish; wait + merge + download */
SCN_API int scn_exchange_publish_host(scn_exchange* x, const uint32_t* host_records, uint64_t* seq_out);
SCN_API int scn_exchange_merge_host(scn_exchange* x, uint64_t seq, uint32_t* host_merged);
SCN_API int scn_exchange_status(scn_exchange* x, uint32_t* timed_out_seq);
SCN_API uint32_t scn_exchange_slots(void);
SCN_API int scn_exchange_destroy(scn_exchange* x);

/* ---- In-process NCCL gather of the per-step records (scn_nccl.cu) -----------------------
 * One process driving several GPUs (csrc/host/sweepProcessor.cpp): ncclCommInitAll over `devices`, then per sweep
 * one grouped ncclAllGather of every device's partial record table [n_steps][record_words] over NVLink and the merge
 * kernel on every device.  host_partials[d] is device d's table (host memory); host_merged receives the merged
 * table (every device ends up with the same one; they are compared).  NCCL is dlopen'ed at the first call. */
typedef struct scn_gather scn_gather;
SCN_API int scn_nccl_gather_create(const int* devices, uint32_t n_devices, uint32_t n_steps, uint32_t record_words,
                                   scn_gather** out);
SCN_API int scn_nccl_gather_merge_host(scn_gather* g, const uint32_t* const* host_partials, uint32_t* host_merged);
SCN_API int scn_nccl_gather_destroy(scn_gather* g);

/* ---- Standalone sample conversion (replaces Utility::*_to_float_complex, utility.cpp:9-84, as called by
 * MessageQueue::AppendSamples, messageQueue.h:190-237) -------------------------------------------------
 * raw: n_buffers buffers of the context's sample kind; out: n_buffers * sample_count fftwf_complex
 * (interleaved float re, im), bit for bit what the reference stores in its queue messages and writes to its
 * recording files (messageQueue.h:126-131): enob wrap, optional DC correction with the reference's
 * unsigned division.  The fused kernel does not need this; the trigger/record path does. */
SCN_API int scn_convert_device(scn_ctx* ctx, const void* d_raw, uint32_t n_buffers, float* d_out, void* stream);
SCN_API int scn_convert_host(scn_ctx* ctx, const void* raw, uint32_t n_buffers, float* out);

/* ---- HackRF sweep-frame pre-pass (replaces HackRFSource::interpolateSamples, hackRFSource.cpp:186-222) ----
 * d_transfers: n_transfers sweep-mode transfers of valid_length bytes each, int8 IQ, device
 * resident, patched IN PLACE exactly as the reference patches them (frame header 0x7F 0x7F + LE64
 * frequency parsed from the first block; samples 0..4 overwritten with sample 5, including the
 * reference's later-iteration quirk).  d_frequency_hz[t] = header frequency (0 if no header); the
 * centre frequency the reference hands the queue is double(frequency + m_scanOffset),
 * m_scanOffset = uint32(uint32(0.75 * sample_rate) / 2.0) (hackRFSource.cpp:111-112,221).
 * d_status[t]: bit 0 = header seen, bits 8.. = number of "frequencyHz != thisFrequencyHz" lines
 * the reference would have printed.  Either output may be NULL.  Afterwards every transfer is
 * valid_length / (2 * sample_count) consecutive buffers for scn_launch_device. */
SCN_API int scn_hackrf_prepass_device(scn_ctx* ctx, void* d_transfers, uint32_t n_transfers,
                                      uint32_t valid_length, uint64_t* d_frequency_hz,
                                      uint32_t* d_status, void* stream);

/* ---- Host-side helpers that restate reference arithmetic --------------------- */
/* uint32_t(useBandWidth * N / 2.0), process.cpp:85. */
SCN_API uint32_t scn_use_window(double use_bandwidth, uint32_t sample_count);
/* uint64_t((center - double(fs/2u)) + double(i * (fs/N))), process.cpp:38-39,55,57. */
SCN_API uint64_t scn_hit_frequency(double center_frequency, uint32_t sample_rate,
                                   uint32_t sample_count, uint32_t bin);
/* Centre-frequency list, frequencyTable.cpp:9-37.  Returns the count; writes min(count, cap). */
SCN_API uint32_t scn_frequency_table(uint32_t sample_rate, double start_frequency,
                                     double stop_frequency, double use_bandwidth,
                                     double dc_ignore_width, double* out, uint32_t cap);
/* Window taps as float, symmetric (M = N-1), process.cpp:18. */
SCN_API int scn_window_build(int win_type, uint32_t sample_count, float* out);
/* Contiguous range of retune steps owned by `rank` of `world`: [begin, end). */
SCN_API void scn_shard_steps(uint32_t n_steps, uint32_t rank, uint32_t world, uint32_t* begin,
                             uint32_t* end);

#ifdef __cplusplus
}
#endif
#endif /* SCANNER_B200_H_ */
