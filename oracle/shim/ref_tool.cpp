// ref_tool -- drives the reference's OWN hot-path sources (compiled unmodified from
// /root/reference with the shim headers in this directory) so the oracle restatement and the GPU
// path can be checked against what the reference code itself computes.  TEST INFRASTRUCTURE.
//
//   ref_tool convert  <kind> <N> <enob> <dc>            stdin: raw buffer(s)   stdout: float32 IQ
//   ref_tool magnitude <N>                              stdin: complex64[N]    stdout: float32[N]
//   ref_tool freqtable <fs> <start> <stop> <useBW> <dcIgnore>    stdout: the reference's own table dump
//   ref_tool scan <kind> <N> <fs> <enob> <dc> <threshold> <win_type> <mode> <buffers_per_sweep>
//                 <raw_file> <freq_file>
//        feeds every buffer through SampleQueue::AppendSamples (messageQueue.h:190-237) on a
//        producer thread and runs ProcessSamples::StartProcessing (process.cpp:316) with ONE worker
//        (the reference's two workers race on the shared FFT object, fft.cpp:20-25);
//        stdout is the reference's own output ("Start scan at ...", "freq %lu power_db %f", ...).
//   ref_tool record <kind> <N> <fs> <enob> <dc> <threshold> <win_type> <buffers_per_sweep> <raw_file> <freq_file>
//                   <file_base> <pre_trigger> <post_trigger>
//        like scan (frequency domain, one worker) with triggered recording on: SampleQueue(doWrite = true) and
//        ProcessSamples(fileNameBase, preTrigger, postTrigger); the reference's writer thread
//        (messageQueue.h:98-139) produces the files.  SetIsDone is held back until the queue has drained,
//        because the reference's writer stops at SetIsDone (messageQueue.h:100-124).
//   ref_tool hackrf_prepass <N> <fs> <start> <stop> <valid_length> <in_file> <out_file>
//        runs HackRFSource::interpolateSamples (hackRFSource.cpp:186-222) on every transfer of
//        <in_file>; <out_file> = per transfer { double returned centre frequency, patched bytes };
//        stdout carries the function's own mismatch prints.
//   ref_tool hackrf_scan <N> <fs> <start> <stop> <threshold> <iterations> <valid_length> <stream_file>
//        the whole HackRF flow with the settings of scan.cpp:177-188,211-223: transfers are delivered
//        to the rx callback the source registers (hackRFSource.cpp:180-264) -> SampleQueue ->
//        ProcessSamples (one worker); stdout is the reference's own output.
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <iostream>
#include <limits>
#include <cassert>
#include <cstdint>
#include <memory>
#include <algorithm>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include "fft.h"
#define private public   // ref_tool polls MessageQueue::IsEmpty() (access only; the header is unmodified)
#include "messageQueue.h"
#undef private
#include "signalSource.h"
#include "process.h"
#define class struct   // interpolateSamples is a private member of `class HackRFSource` (access only; the source is unmodified)
#include "hackRFSource.h"
#undef class

static std::vector<char> read_all(FILE* f) {
  std::vector<char> data;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
  return data;
}

static std::vector<char> read_file(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "ref_tool: cannot open %s\n", path); exit(2); }
  std::vector<char> d = read_all(f);
  fclose(f);
  return d;
}

static size_t bytes_per_sample(int kind) {
  switch (kind) {
    case SampleQueue::ByteComplex: return 2;
    case SampleQueue::Short: return 4;
    case SampleQueue::ShortComplex: return 4;
    case SampleQueue::FloatComplex: return 8;
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: ref_tool <convert|magnitude|freqtable|scan> ...\n"); return 2; }
  std::string cmd = argv[1];
  if (cmd == "convert" && argc == 6) {
    int kind = atoi(argv[2]); uint32_t n = atoi(argv[3]), enob = atoi(argv[4]); bool dc = atoi(argv[5]) != 0;
    std::vector<char> raw = read_all(stdin);
    size_t bb = n * bytes_per_sample(kind);
    std::vector<float> out(2 * size_t(n));
    for (size_t off = 0; off + bb <= raw.size(); off += bb) {
      fftwf_complex* dst = reinterpret_cast<fftwf_complex*>(out.data());
      if (kind == SampleQueue::ByteComplex)
        Utility::byte_complex_to_float_complex(reinterpret_cast<int8_t(*)[2]>(raw.data() + off), dst, n, enob, dc);
      else if (kind == SampleQueue::ShortComplex)
        Utility::short_complex_to_float_complex(reinterpret_cast<int16_t(*)[2]>(raw.data() + off), dst, n, enob, dc);
      else if (kind == SampleQueue::Short)
        Utility::short_complex_to_float_complex(reinterpret_cast<int16_t*>(raw.data() + off),
                                                reinterpret_cast<int16_t*>(raw.data() + off) + n, dst, n, enob, dc);
      else
        memcpy(out.data(), raw.data() + off, bb);
      fwrite(out.data(), sizeof(float), out.size(), stdout);
    }
    return 0;
  }
  if (cmd == "magnitude" && argc == 3) {
    uint32_t n = atoi(argv[2]);
    std::vector<char> raw = read_all(stdin);
    std::vector<float> out(n);
    Utility::complex_to_magnitude(reinterpret_cast<fftwf_complex*>(raw.data()), out.data(), n);
    fwrite(out.data(), sizeof(float), n, stdout);
    return 0;
  }
  if (cmd == "freqtable" && argc == 7) {
    FrequencyTable table(uint32_t(atof(argv[2])), atof(argv[3]), atof(argv[4]), atof(argv[5]), atof(argv[6]));
    printf("count %u\n", table.GetFrequencyCount());
    return 0;
  }
  if (cmd == "scan" && argc == 13) {
    int kind = atoi(argv[2]);
    uint32_t n = atoi(argv[3]), fs = uint32_t(atof(argv[4])), enob = atoi(argv[5]);
    bool dc = atoi(argv[6]) != 0;
    float threshold = float(atof(argv[7]));
    int win = atoi(argv[8]), mode = atoi(argv[9]);
    uint32_t per_sweep = atoi(argv[10]);
    std::vector<char> raw = read_file(argv[11]);
    std::vector<char> fr = read_file(argv[12]);
    const double* freqs = reinterpret_cast<const double*>(fr.data());
    size_t bb = n * bytes_per_sample(kind);
    size_t nbuf = raw.size() / bb;
    if (fr.size() / sizeof(double) < nbuf) { fprintf(stderr, "ref_tool: too few frequencies\n"); return 2; }
    SampleQueue queue(SampleQueue::SampleKind(kind), enob, n, 1024, dc, false);
    ProcessSamples process(n, fs, enob, threshold, gr::fft::window::win_type(win), ProcessSamples::Mode(mode), 1);
    std::thread producer([&]() {
      for (size_t b = 0; b < nbuf; b++) {
        char* p = raw.data() + b * bb;
        time_t tm = (per_sweep && (b % per_sweep) == 0) ? time_t(1000000000 + b) : 0;
        if (kind == SampleQueue::ByteComplex)
          queue.AppendSamples(reinterpret_cast<int8_t(*)[2]>(p), freqs[b], tm);
        else if (kind == SampleQueue::ShortComplex)
          queue.AppendSamples(reinterpret_cast<int16_t(*)[2]>(p), freqs[b], tm);
        else if (kind == SampleQueue::Short)
          queue.AppendSamples(reinterpret_cast<int16_t*>(p), reinterpret_cast<int16_t*>(p) + n, freqs[b], tm);
        else
          queue.AppendSamples(reinterpret_cast<fftwf_complex*>(p), freqs[b], tm);
      }
      queue.SetIsDone();
    });
    process.StartProcessing(queue);
    producer.join();
    fflush(stdout);
    return 0;
  }
  if (cmd == "record" && argc == 15) {
    int kind = atoi(argv[2]);
    uint32_t n = atoi(argv[3]), fs = uint32_t(atof(argv[4])), enob = atoi(argv[5]);
    bool dc = atoi(argv[6]) != 0;
    float threshold = float(atof(argv[7]));
    int win = atoi(argv[8]);
    uint32_t per_sweep = atoi(argv[9]);
    std::vector<char> raw = read_file(argv[10]);
    std::vector<char> fr = read_file(argv[11]);
    std::string base = argv[12];
    uint32_t pre = atoi(argv[13]), post = atoi(argv[14]);
    const double* freqs = reinterpret_cast<const double*>(fr.data());
    size_t bb = n * bytes_per_sample(kind);
    size_t nbuf = raw.size() / bb;
    {
      ProcessSamples process(n, fs, enob, threshold, gr::fft::window::win_type(win), ProcessSamples::FrequencyDomain, 1,
                             base, 0.75, 0.0, pre, post);
      SampleQueue queue(SampleQueue::SampleKind(kind), enob, n, 1024, dc, true);
      std::thread producer([&]() {
        for (size_t b = 0; b < nbuf; b++) {
          char* p = raw.data() + b * bb;
          time_t tm = (per_sweep && (b % per_sweep) == 0) ? time_t(1000000000 + b) : 0;
          if (kind == SampleQueue::ByteComplex)
            queue.AppendSamples(reinterpret_cast<int8_t(*)[2]>(p), freqs[b], tm);
          else if (kind == SampleQueue::ShortComplex)
            queue.AppendSamples(reinterpret_cast<int16_t(*)[2]>(p), freqs[b], tm);
          else if (kind == SampleQueue::Short)
            queue.AppendSamples(reinterpret_cast<int16_t*>(p), reinterpret_cast<int16_t*>(p) + n, freqs[b], tm);
          else
            queue.AppendSamples(reinterpret_cast<fftwf_complex*>(p), freqs[b], tm);
          usleep(2000);                    // a live source's pace: the writer thread is back in its wait between triggers
        }
        while (!queue.IsEmpty()) usleep(1000);
        usleep(300000);                    // last message processed, writer caught up
        queue.SetIsDone();
      });
      process.StartProcessing(queue);
      producer.join();
      fflush(stdout);
      // every window was closed by the writer (messageQueue.h:132-135).  ~MessageQueue would fclose the last
      // FILE* a second time (messageQueue.h:186-188: m_writeFile is never reset) and abort in glibc, so leave here.
      _exit(0);
    }
  }
  if (cmd == "hackrf_prepass" && argc == 9) {
    uint32_t n = atoi(argv[2]), fs = uint32_t(atof(argv[3]));
    double start = atof(argv[4]), stop = atof(argv[5]);
    size_t valid = strtoul(argv[6], nullptr, 0);
    std::vector<char> stream = read_file(argv[7]);
    FILE* out = fopen(argv[8], "wb");
    if (!out) { fprintf(stderr, "ref_tool: cannot write %s\n", argv[8]); return 2; }
    HackRFSource source("hackrf", fs, n, start, stop);
    for (size_t off = 0; off + valid <= stream.size(); off += valid) {
      hackrf_transfer t;
      memset(&t, 0, sizeof(t));
      t.buffer = reinterpret_cast<uint8_t*>(stream.data() + off);
      t.buffer_length = t.valid_length = int(valid);
      double f = source.interpolateSamples(&t);
      fwrite(&f, sizeof(f), 1, out);
      fwrite(t.buffer, 1, valid, out);
    }
    fclose(out);
    fflush(stdout);
    _exit(0);   // ~HackRFSource prints through a bad format; nothing left to flush
  }
  if (cmd == "hackrf_scan" && argc == 10) {
    uint32_t n = atoi(argv[2]), fs = uint32_t(atof(argv[3]));
    double start = atof(argv[4]), stop = atof(argv[5]);
    float threshold = float(atof(argv[6]));
    uint32_t iterations = atoi(argv[7]);
    size_t valid = strtoul(argv[8], nullptr, 0);
    std::vector<char> stream = read_file(argv[9]);
    HackRFSource* source = new HackRFSource("hackrf", fs, n, start, stop);
    // scan.cpp:182-188: enob 8, DC correction on, ByteComplex, dcIgnoreWidth forced to 0
    ProcessSamples process(n, fs, 8, threshold, gr::fft::window::WIN_BLACKMAN_HARRIS, ProcessSamples::FrequencyDomain, 1);
    SampleQueue queue(SampleQueue::ByteComplex, 8, n, 1024, true, false);
    source->StartStreaming(iterations, queue);
    std::thread feeder([&]() {
      for (size_t off = 0; off + valid <= stream.size() && !shim_hackrf_rx_stopped(); off += valid)
        shim_hackrf_deliver(reinterpret_cast<uint8_t*>(stream.data() + off), int(valid));
      for (int spin = 0; spin < 2000 && !shim_hackrf_rx_stopped(); spin++) usleep(1000);
      if (!shim_hackrf_rx_stopped()) { fprintf(stderr, "ref_tool: stream ended before the source finished\n"); _exit(3); }
    });
    process.StartProcessing(queue);
    feeder.join();
    fflush(stdout);
    _exit(0);
  }
  fprintf(stderr, "ref_tool: bad arguments\n");
  return 2;
}
