// scan_b200 -- wires source -> SampleQueue -> ProcessSamples the way the reference's main() does
// (scan.cpp:137-239), with the GPU consumer.  No SDR SDKs: the source is synthetic or a replay.
//
//   scan_b200 replay <kind> <N> <fs> <enob> <dc> <threshold> <win_type> <mode> <buffers_per_sweep>
//                    <raw_file> <freq_file> [threads] [legacy]
//       same arguments and input files as oracle/_ref/ref_tool scan; prints what the reference prints.
//       "legacy" routes through SampleBuffer + GpuProcessInterface instead of SampleQueue.
//   scan_b200 synth <kind> <N> <fs> <enob> <dc> <threshold> <start> <stop> <buffers_per_step>
//                   <iterations> <seed> [threads] [averaging]
//       seeded SyntheticSource sweep (first sweep dropped like the reference: needs iterations >= 2).
//   scan_b200 record <kind> <N> <fs> <enob> <dc> <threshold> <win_type> <buffers_per_sweep> <raw_file> <freq_file>
//                    <file_base> <pre_trigger> <post_trigger> [threads] [averaging]
//       same arguments as oracle/_ref/ref_tool record: triggered recording on (SampleQueue doWrite, ProcessSamples
//       fileNameBase / preTrigger / postTrigger); file-name stamps 1500000000 + 1000 k like ref_tool's clock.
//   scan_b200 hackrf <N> <fs> <start> <stop> <threshold> <iterations> <valid_length> <stream_file> [threads]
//       same arguments and capture file as oracle/_ref/ref_tool hackrf_scan: HackRFSweepSource replays the
//       sweep-mode transfers with scan.cpp:177-188's settings (enob 8, DC correction on, ByteComplex).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "buffer.h"
#include "hackrfSweepSource.h"
#include "process.h"
#include "sampleBuffer.h"
#include "sampleQueue.h"
#include "sweepProcessor.h"
#include "syntheticSource.h"

static std::vector<char> ReadFile(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  std::vector<char> data;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
  fclose(f);
  return data;
}

static int RunLegacy(int kind, uint32_t n, uint32_t fs, uint32_t enob, bool dc, float thr, int win,
                     const std::vector<char>& raw, const double* freqs, size_t nbuf) {
  // SampleBuffer + GpuProcessInterface: the legacy surface named by north_star.
  SampleBuffer::SampleKind sk = kind == SCN_KIND_SHORT ? SampleBuffer::Short
                              : kind == SCN_KIND_SHORT_COMPLEX ? SampleBuffer::ShortComplex : SampleBuffer::FloatComplex;
  SampleBuffer sb(sk, enob, n, 64);
  std::vector<float> window(n);
  scn_window_build(win, n, window.data());
  scn_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.sample_count = n; cfg.sample_rate = fs; cfg.enob = enob; cfg.sample_kind = sb.GetScnKind();
  cfg.correct_dc_offset = dc; cfg.averaging = 1; cfg.threshold = thr;
  cfg.use_window = scn_use_window(0.75, n); cfg.dc_ignore_window = 4; cfg.window = window.data();
  cfg.max_spectra = 16; cfg.flags = 0;
  scn_ctx* ctx = nullptr;
  if (scn_create(&cfg, &ctx) != SCN_OK) { fprintf(stderr, "scn_create: %s\n", scn_last_error()); return 1; }
  std::thread producer([&] {
    const size_t bb = sb.GetBufferBytes();
    for (size_t b = 0; b < nbuf; b++) {
      char* p = const_cast<char*>(raw.data()) + b * bb;
      if (sk == SampleBuffer::Short) sb.AppendSamples(reinterpret_cast<int16_t*>(p), reinterpret_cast<int16_t*>(p) + n, freqs[b]);
      else if (sk == SampleBuffer::ShortComplex) sb.AppendSamples(reinterpret_cast<int16_t(*)[2]>(p), freqs[b]);
      else sb.AppendSamples(reinterpret_cast<fftwf_complex*>(p), freqs[b]);
    }
    sb.SetIsDone();
  });
  GpuProcessInterface gpu(ctx, 16, 1, n);
  std::vector<double> f;
  const uint32_t words = scn_mask_words(ctx);
  while (uint32_t got = sb.GetNextSamples(&gpu, f, 16)) {
    for (uint32_t s = 0; s < got; s++)
      for (uint32_t i = 0; i < n; i++)
        if (gpu.GetHitMasks()[size_t(s) * words + (i >> 5)] >> (i & 31) & 1u)
          printf("freq %lu\n", (unsigned long)scn_hit_frequency(f[s], fs, n, i));
  }
  producer.join();
  scn_destroy(ctx);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: scan_b200 <replay|synth> ...\n"); return 2; }
  const std::string cmd = argv[1];
  if (cmd == "replay" && argc >= 13) {
    const int kind = atoi(argv[2]);
    const uint32_t n = atoi(argv[3]), fs = uint32_t(atof(argv[4])), enob = atoi(argv[5]);
    const bool dc = atoi(argv[6]) != 0;
    const float thr = float(atof(argv[7]));
    const int win = atoi(argv[8]), mode = atoi(argv[9]);
    const uint32_t perSweep = atoi(argv[10]);
    std::vector<char> raw = ReadFile(argv[11]);
    std::vector<char> fr = ReadFile(argv[12]);
    const uint32_t threads = argc > 13 ? atoi(argv[13]) : 1;
    const bool legacy = argc > 14 && std::string(argv[14]) == "legacy";
    const size_t bb = SyntheticSource::BufferBytes(SampleQueue::SampleKind(kind), n);
    const size_t nbuf = raw.size() / bb;
    const double* freqs = reinterpret_cast<const double*>(fr.data());
    if (legacy) return RunLegacy(kind, n, fs, enob, dc, thr, win, raw, freqs, nbuf);
    ReplaySource source(SampleQueue::SampleKind(kind), raw.data(), freqs, nbuf, perSweep, fs, n);
    if (getenv("SCN_APPEND_BATCH")) source.SetAppendBatch(atoi(getenv("SCN_APPEND_BATCH")));   // AppendSamplesBatch path
    ProcessSamples process(n, fs, enob, thr, win, ProcessSamples::Mode(mode), threads);
    SampleQueue queue(SampleQueue::SampleKind(kind), enob, n, 1024, dc, false);
    source.Start();
    source.StartStreaming(1, queue);
    if (getenv("SCN_STAGING_COPY")) process.SetZeroCopy(false);
    process.StartProcessing(queue);
    source.Join();
    fflush(stdout);
    return 0;
  }
  if (cmd == "record" && argc >= 15) {
    const int kind = atoi(argv[2]);
    const uint32_t n = atoi(argv[3]), fs = uint32_t(atof(argv[4])), enob = atoi(argv[5]);
    const bool dc = atoi(argv[6]) != 0;
    const float thr = float(atof(argv[7]));
    const int win = atoi(argv[8]);
    const uint32_t perSweep = atoi(argv[9]);
    std::vector<char> raw = ReadFile(argv[10]);
    std::vector<char> fr = ReadFile(argv[11]);
    const std::string base = argv[12];
    const uint32_t pre = atoi(argv[13]), post = atoi(argv[14]);
    const uint32_t threads = argc > 15 ? atoi(argv[15]) : 1;
    const uint32_t averaging = argc > 16 ? atoi(argv[16]) : 1;
    const size_t bb = SyntheticSource::BufferBytes(SampleQueue::SampleKind(kind), n);
    const size_t nbuf = raw.size() / bb;
    const double* freqs = reinterpret_cast<const double*>(fr.data());
    ReplaySource source(SampleQueue::SampleKind(kind), raw.data(), freqs, nbuf, perSweep, fs, n);
    if (getenv("SCN_APPEND_BATCH")) source.SetAppendBatch(atoi(getenv("SCN_APPEND_BATCH")));
    ProcessSamples process(n, fs, enob, thr, win, ProcessSamples::FrequencyDomain, threads, base, 0.75, 0.0, pre, post);
    process.SetAveraging(averaging);
    uint64_t stamps = 0;
    process.SetClock([&stamps]() { return time_t(1500000000 + 1000 * stamps++); });
    {
      SampleQueue queue(SampleQueue::SampleKind(kind), enob, n, 1024, dc, true);
      source.Start();
      source.StartStreaming(1, queue);
      if (getenv("SCN_STAGING_COPY")) process.SetZeroCopy(false);
    process.StartProcessing(queue);
      source.Join();
    }                                                           // ~SampleQueue flushes and joins the writer
    fflush(stdout);
    return 0;
  }
  if (cmd == "hackrf" && argc >= 10) {
    const uint32_t n = atoi(argv[2]), fs = uint32_t(atof(argv[3]));
    const double start = atof(argv[4]), stop = atof(argv[5]);
    const float thr = float(atof(argv[6]));
    const uint32_t iterations = atoi(argv[7]);
    const uint32_t valid = uint32_t(strtoul(argv[8], nullptr, 0));
    std::vector<char> stream = ReadFile(argv[9]);
    const uint32_t threads = argc > 10 ? atoi(argv[10]) : 1;
    HackRFSweepSource source("hackrf", fs, n, start, stop);
    source.SetCapture(reinterpret_cast<uint8_t*>(stream.data()), stream.size(), valid);
    source.SetReplayClock(1500000000, 1000);
    ProcessSamples process(n, fs, 8, thr, SCN_WIN_BLACKMAN_HARRIS, ProcessSamples::FrequencyDomain, threads);
    SampleQueue queue(SampleQueue::ByteComplex, 8, n, 1024, true, false);
    source.StartStreaming(iterations, queue);
    if (getenv("SCN_STAGING_COPY")) process.SetZeroCopy(false);
    process.StartProcessing(queue);
    source.Join();
    fflush(stdout);
    return 0;
  }
  if (cmd == "synth" && argc >= 13) {
    const int kind = atoi(argv[2]);
    const uint32_t n = atoi(argv[3]), fs = uint32_t(atof(argv[4])), enob = atoi(argv[5]);
    const bool dc = atoi(argv[6]) != 0;
    const float thr = float(atof(argv[7]));
    const double start = atof(argv[8]), stop = atof(argv[9]);
    const uint32_t perStep = atoi(argv[10]), iterations = atoi(argv[11]);
    const uint64_t seed = strtoull(argv[12], nullptr, 0);
    const uint32_t threads = argc > 13 ? atoi(argv[13]) : 2;       // the reference runs 2 (scan.cpp:217)
    const uint32_t averaging = argc > 14 ? atoi(argv[14]) : 1;
    SyntheticSource source(SampleQueue::SampleKind(kind), enob, seed, perStep, fs, n, start, stop);
    ProcessSamples process(n, fs, enob, thr, SCN_WIN_BLACKMAN_HARRIS, ProcessSamples::FrequencyDomain, threads);
    process.SetAveraging(averaging);
    SampleQueue queue(SampleQueue::SampleKind(kind), enob, n, 1024, dc, false);
    const auto t0 = std::chrono::steady_clock::now();
    source.Start();
    source.StartStreaming(iterations, queue);
    if (getenv("SCN_STAGING_COPY")) process.SetZeroCopy(false);
    process.StartProcessing(queue);
    source.Join();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("Elapsed time = %f ms\n", ms);                         // scan.cpp:43-47
    fprintf(stderr, "buffers %lu hits %lu launches %lu zero-copy batches %lu\n", (unsigned long)process.GetBuffersProcessed(),
            (unsigned long)process.GetHitCount(), (unsigned long)process.GetLaunchCount(),
            (unsigned long)process.GetZeroCopyBatches());
    return 0;
  }
  if (cmd == "sweep" && argc >= 14) {
    // scan_b200 sweep <kind> <N> <fs> <enob> <dc> <threshold> <start> <stop> <buffers_per_step> <iterations> <seed>
    //                 <gpus> [averaging] [nccl|peer] [report]
    // The synth scan through SweepProcessor: one worker per GPU, retune steps split across them, per-sweep records
    // exchanged with NCCL (default) or the NVLink peer-memory windows.  Prints what `synth` prints, whatever <gpus>.
    const int kind = atoi(argv[2]);
    const uint32_t n = atoi(argv[3]), fs = uint32_t(atof(argv[4])), enob = atoi(argv[5]);
    const bool dc = atoi(argv[6]) != 0;
    const float thr = float(atof(argv[7]));
    const double start = atof(argv[8]), stop = atof(argv[9]);
    const uint32_t perStep = atoi(argv[10]), iterations = atoi(argv[11]);
    const uint64_t seed = strtoull(argv[12], nullptr, 0);
    const uint32_t gpus = atoi(argv[13]) > 0 ? atoi(argv[13]) : 1;
    const uint32_t averaging = argc > 14 ? atoi(argv[14]) : 1;
    const bool peer = argc > 15 && std::string(argv[15]) == "peer";
    const bool report = argc > 16 && atoi(argv[16]) != 0;
    SyntheticSource source(SampleQueue::SampleKind(kind), enob, seed, perStep, fs, n, start, stop);
    std::vector<double> steps(scn_frequency_table(fs, start, stop, 0.75, 0.0, nullptr, 0));
    scn_frequency_table(fs, start, stop, 0.75, 0.0, steps.data(), uint32_t(steps.size()));
    std::vector<int> devices(gpus);
    for (uint32_t d = 0; d < gpus; d++) devices[d] = int(d);
    SweepProcessor process(n, fs, enob, thr, SCN_WIN_BLACKMAN_HARRIS, steps, devices,
                           peer ? SweepProcessor::PeerMemory : SweepProcessor::NcclAllGather);
    process.SetAveraging(averaging);
    process.SetSweepReport(report);
    SampleQueue queue(SampleQueue::SampleKind(kind), enob, n, 1024, dc, false);
    const auto t0 = std::chrono::steady_clock::now();
    source.Start();
    source.StartStreaming(iterations, queue);
    process.StartProcessing(queue);
    source.Join();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("Elapsed time = %f ms\n", ms);
    fprintf(stderr, "buffers %lu hits %lu launches %lu sweeps %u per-gpu", (unsigned long)process.GetBuffersProcessed(),
            (unsigned long)process.GetHitCount(), (unsigned long)process.GetLaunchCount(), process.GetSweepCount());
    for (uint32_t d = 0; d < gpus; d++) fprintf(stderr, " %lu", (unsigned long)process.GetBuffersOnDevice(d));
    uint64_t recHits = 0, recSpectra = 0;
    const std::vector<uint32_t>& rec = process.GetLastSweepRecords();
    for (size_t s = 0; s * (n / 32 + 2) < rec.size(); s++) { recHits += rec[s * (n / 32 + 2)]; recSpectra += rec[s * (n / 32 + 2) + 1]; }
    fprintf(stderr, " last-sweep-records hits %lu spectra %lu\n", (unsigned long)recHits, (unsigned long)recSpectra);
    return 0;
  }
  if (cmd == "bench" && argc >= 8) {
    // scan_b200 bench <kind> <N> <enob> <dc> <distinct_buffers> <total_buffers> [threads] [max_batch] [producers]
    //                 [append_batch] [linger_us]
    // Throughput of the plugin surface itself: `producers` ReplaySource threads (think: one per SDR / USB transfer
    // thread) push `total` raw buffers, `append_batch` per AppendSamplesBatch call (1 == the reference's one
    // AppendSamples per buffer), through ONE SampleQueue into GPU ProcessSamples workers (printing off).
    const int kind = atoi(argv[2]);
    const uint32_t n = atoi(argv[3]), enob = atoi(argv[4]);
    const bool dc = atoi(argv[5]) != 0;
    const size_t distinct = strtoull(argv[6], nullptr, 0), total = strtoull(argv[7], nullptr, 0);
    const uint32_t threads = argc > 8 ? atoi(argv[8]) : 2;
    const uint32_t maxBatch = argc > 9 ? atoi(argv[9]) : 4096;
    const uint32_t producers = argc > 10 && atoi(argv[10]) > 0 ? atoi(argv[10]) : 1;
    const uint32_t appendBatch = argc > 11 && atoi(argv[11]) > 0 ? atoi(argv[11]) : 1;
    const uint32_t lingerUs = argc > 12 ? atoi(argv[12]) : 0;
    const uint32_t fs = 20000000;
    const size_t bb = SyntheticSource::BufferBytes(SampleQueue::SampleKind(kind), n);
    std::vector<char> pool(distinct * bb);
    {
      SyntheticSource gen(SampleQueue::SampleKind(kind), enob ? enob : 12, 1234, 1, fs, n, 2.4e9, 0.0);
      for (size_t b = 0; b < distinct; b++) gen.Generate(0, 0, uint32_t(b), pool.data() + b * bb);
    }
    // every producer replays its own capture ring (>= 128 MB: far beyond the CPU caches, so the hand-off's memcpy
    // reads DRAM as it would behind a USB/PCIe DMA), copied into memory the producer thread touches first
    size_t ringBuffers = (size_t(128) << 20) / bb;
    if (ringBuffers > total / producers) ringBuffers = total / producers ? total / producers : 1;
    std::vector<char> raw(ringBuffers * bb);
    std::vector<double> freqs(ringBuffers);
    for (size_t b = 0; b < ringBuffers; b++) {
      memcpy(raw.data() + b * bb, pool.data() + (b % distinct) * bb, bb);
      freqs[b] = 2.4075e9 + 15e6 * double(b % 50);
    }
    std::vector<std::unique_ptr<ReplaySource>> sources;
    for (uint32_t pr = 0; pr < producers; pr++) {
      const size_t share = total * (pr + 1) / producers - total * pr / producers;
      sources.emplace_back(new ReplaySource(SampleQueue::SampleKind(kind), raw.data(), freqs.data(), ringBuffers, 0, fs, n));
      sources.back()->SetAppendBatch(appendBatch);
      sources.back()->SetRepeat(share);
      sources.back()->SetLocalCopy(true);
    }
    ProcessSamples process(n, fs, enob, 25.0f, SCN_WIN_BLACKMAN_HARRIS, ProcessSamples::FrequencyDomain, threads);
    process.SetOutput(nullptr);
    process.SetMaxBatch(maxBatch);
    if (lingerUs) process.SetBatchLinger(maxBatch / 2, lingerUs);
    if (getenv("SCN_STAGING_COPY")) process.SetZeroCopy(false);
    SampleQueue queue(SampleQueue::SampleKind(kind), enob, n, 4 * maxBatch, dc, false);
    queue.SetDropFirstSweep(false);
    queue.SetProducerCount(producers);
    const auto t0 = std::chrono::steady_clock::now();
    for (auto& src : sources) {
      src->Start();
      src->StartStreaming(1, queue);
    }
    process.StartProcessing(queue);
    for (auto& src : sources) src->Join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("plugin-surface throughput: %.1f Msamples/s (%zu buffers of %u samples, kind %d, %u producer threads x %u "
           "buffers per append, %u worker threads, batch <= %u, linger %u us, %lu hits, %lu launches of which %lu "
           "straight from the pinned slab, %.3f s)\n",
           double(total) * n / sec / 1e6, total, n, kind, producers, appendBatch, threads, maxBatch, lingerUs,
           (unsigned long)process.GetHitCount(), (unsigned long)process.GetLaunchCount(),
           (unsigned long)process.GetZeroCopyBatches(), sec);
    return 0;
  }
  fprintf(stderr, "scan_b200: bad arguments\n");
  return 2;
}
