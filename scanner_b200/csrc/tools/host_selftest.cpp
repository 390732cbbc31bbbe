// host_selftest -- CPU-only checks of the plugin-surface plumbing (no GPU calls): FrequencyTable
// iteration, SampleQueue semantics (first-sweep drop, sequence ids, FIFO, bounded blocking, batch
// multiples, raw layout per kind), SyntheticSource determinism, SampleBuffer visitor protocol.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include <thread>
#include <vector>

#include "buffer.h"
#include "frequencyTable.h"
#include "hackrfSweepSource.h"
#include "sampleBuffer.h"
#include "sampleQueue.h"
#include "syntheticSource.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

struct CountingVisitor : ProcessInterface<uint8_t> {
  int begins = 0, ends = 0; uint64_t seq = 0; uint32_t total = 0, seen = 0, blocks = 0;
  CountingVisitor() : ProcessInterface<uint8_t>(false) {}
  void Begin(uint64_t s, uint32_t t) override { begins++; seq = s; total = t; seen = 0; blocks = 0; }
  void Process(const uint8_t*, uint32_t c) override { seen += c; blocks++; }
  void End() override { ends++; CHECK(seen == total); }
};

static std::vector<char> ReadAll(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  std::vector<char> data;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
  fclose(f);
  return data;
}

// host_selftest hackrf_prepass <N> <fs> <start> <stop> <valid_length> <in_file> <out_file>
//   HackRFSweepSource::InterpolateSamples on every transfer; output format of `ref_tool hackrf_prepass`.
// host_selftest hackrf_queue <N> <fs> <start> <stop> <iterations> <valid_length> <stream_file> <out_file>
//   replays the capture through RxCallback into a SampleQueue (no GPU) and dumps, per accepted message,
//   { double frequency, int64 time, uint64 sequenceId, raw bytes }.
// host_selftest queuebench <kind> <N> <total_buffers> [consumers] [max_batch] [busy_us_per_batch]
//   throughput of the hand-off alone (no GPU): one producer appends `total` buffers, consumers drain batches and
//   return the messages at once.  Tells how much of the plugin-surface time is the queue itself.
static int QueueBench(int argc, char** argv) {
  const int kind = atoi(argv[2]);
  const uint32_t n = atoi(argv[3]);
  const size_t total = strtoull(argv[4], nullptr, 0);
  const uint32_t consumers = argc > 5 ? atoi(argv[5]) : 1;
  const uint32_t maxBatch = argc > 6 ? atoi(argv[6]) : 1024;
  const uint32_t busyUs = argc > 7 ? atoi(argv[7]) : 0;       // stand-in for the launch + collect time of a batch
  SampleQueue q(SampleQueue::SampleKind(kind), 8, n, 1024, false, false);
  q.SetDropFirstSweep(false);
  const size_t bb = q.GetBufferBytes();
  std::vector<char> raw(bb * 64, 1);
  const auto t0 = std::chrono::steady_clock::now();
  std::thread producer([&] {
    for (size_t b = 0; b < total; b++) {
      char* p = raw.data() + (b % 64) * bb;
      if (kind == SampleQueue::ByteComplex) q.AppendSamples(reinterpret_cast<int8_t(*)[2]>(p), 1e6, 0);
      else if (kind == SampleQueue::ShortComplex) q.AppendSamples(reinterpret_cast<int16_t(*)[2]>(p), 1e6, 0);
      else q.AppendSamples(reinterpret_cast<fftwf_complex*>(p), 1e6, 0);
    }
    q.SetIsDone();
  });
  std::atomic<uint64_t> seen{0};
  std::vector<std::thread> pool;
  for (uint32_t c = 0; c < consumers; c++)
    pool.emplace_back([&] {
      std::vector<SampleQueue::MessageType*> batch;
      std::vector<char> staging(bb * maxBatch);
      while (uint32_t got = q.GetNextBatch(batch, maxBatch, 1, true)) {
        for (uint32_t i = 0; i < got; i++) memcpy(staging.data() + size_t(i) * bb, batch[i]->GetData(), bb);   // the consumer's copy
        if (busyUs) {
          const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(busyUs);
          while (std::chrono::steady_clock::now() < until) {}
        }
        q.MessageProcessed(batch);
        seen += got;
      }
    });
  producer.join();
  for (auto& t : pool) t.join();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("queuebench kind %d N %u: %lu buffers in %.3f s = %.1f Msamples/s = %.2f GB/s (%u consumer(s), batch <= %u)\n", kind, n,
         (unsigned long)seen.load(), sec, double(seen) * n / sec / 1e6, double(seen) * bb / sec / 1e9, consumers, maxBatch);
  return 0;
}

static int HackrfModes(int argc, char** argv) {
  const std::string cmd = argv[1];
  if (cmd == "queuebench" && argc >= 5) return QueueBench(argc, argv);
  const uint32_t n = atoi(argv[2]), fs = uint32_t(atof(argv[3]));
  const double start = atof(argv[4]), stop = atof(argv[5]);
  if (cmd == "hackrf_prepass" && argc == 9) {
    const uint32_t valid = uint32_t(strtoul(argv[6], nullptr, 0));
    std::vector<char> stream = ReadAll(argv[7]);
    FILE* out = fopen(argv[8], "wb");
    CHECK(out != nullptr);
    HackRFSweepSource source("hackrf", fs, n, start, stop);
    for (size_t off = 0; off + valid <= stream.size(); off += valid) {
      uint8_t* p = reinterpret_cast<uint8_t*>(stream.data() + off);
      const double f = source.InterpolateSamples(p, valid);
      fwrite(&f, sizeof(f), 1, out);
      fwrite(p, 1, valid, out);
    }
    fclose(out);
    return 0;
  }
  if (cmd == "hackrf_queue" && argc == 10) {
    const uint32_t iterations = atoi(argv[6]);
    const uint32_t valid = uint32_t(strtoul(argv[7], nullptr, 0));
    std::vector<char> stream = ReadAll(argv[8]);
    FILE* out = fopen(argv[9], "wb");
    CHECK(out != nullptr);
    HackRFSweepSource source("hackrf", fs, n, start, stop);
    source.SetCapture(reinterpret_cast<uint8_t*>(stream.data()), stream.size(), valid);
    source.SetReplayClock(1500000000, 1000);
    SampleQueue queue(SampleQueue::ByteComplex, 8, n, uint32_t(stream.size() / (2 * n)) + 8, true, false);
    source.StartStreaming(iterations, queue);
    source.Join();
    CHECK(source.GetStreamDone());
    while (SampleQueue::MessageType* m = queue.GetNextSamples()) {
      const double f = m->GetHeader().m_frequency;
      const int64_t t = int64_t(m->GetHeader().m_time);
      const uint64_t seq = m->GetHeader().m_sequenceId;
      fwrite(&f, sizeof(f), 1, out);
      fwrite(&t, sizeof(t), 1, out);
      fwrite(&seq, sizeof(seq), 1, out);
      fwrite(m->GetData(), 1, m->GetDataBytes(), out);
      queue.MessageProcessed(m);
    }
    fclose(out);
    return 0;
  }
  fprintf(stderr, "host_selftest: bad arguments\n");
  return 2;
}

int main(int argc, char** argv) {
  if (argc > 1) return HackrfModes(argc, argv);
  // ---- FrequencyTable (frequencyTable.cpp:9-47)
  FrequencyTable ft(20000000, 2.4e9, 3.15e9, 0.75, 0.0, false);
  CHECK(ft.GetFrequencyCount() == 50);
  CHECK(ft.GetCurrentFrequency() == 2407500000.0 && ft.GetIsScanStart());
  CHECK(ft.GetNextFrequency() == 2422500000.0 && !ft.GetIsScanStart());
  for (int i = 0; i < 49; i++) ft.GetNextFrequency();
  CHECK(ft.GetIsScanStart() && ft.GetIterationCount() == 1 && ft.GetCurrentFrequency() == 2407500000.0);
  CHECK(ft.GetStartFrequency() == 2407500000.0 && ft.GetStopFrequency() == 2407500000.0 + 49 * 15e6);
  FrequencyTable one(8000000, 3e8, 0.0, 0.75, 0.0, false);
  CHECK(one.GetFrequencyCount() == 1 && one.GetCurrentFrequency() == 303000000.0);
  one.GetNextFrequency();
  CHECK(one.GetIterationCount() == 1 && one.GetIsScanStart());

  // ---- SampleQueue: first-sweep drop, sequence ids, FIFO, raw layout
  {
    const uint32_t N = 64;
    SampleQueue q(SampleQueue::ShortComplex, 12, N, 8, false, false);
    std::vector<int16_t> buf(2 * N);
    for (int b = 0; b < 9; b++) {                       // 3 sweeps of 3 buffers
      for (uint32_t i = 0; i < 2 * N; i++) buf[i] = int16_t(b * 100 + i);
      q.AppendSamples(reinterpret_cast<int16_t(*)[2]>(buf.data()), 1e6 * b, (b % 3 == 0) ? time_t(1000 + b) : 0);
    }
    q.SetIsDone();
    CHECK(q.GetDroppedCount() == 3 && q.GetAcceptedCount() == 6);       // messageQueue.h:67-72
    for (int b = 3; b < 9; b++) {
      SampleQueue::MessageType* m = q.GetNextSamples();
      CHECK(m != nullptr);
      CHECK(m->GetHeader().m_sequenceId == uint64_t(b - 3));
      CHECK(m->GetHeader().m_frequency == 1e6 * b);
      CHECK((m->GetHeader().m_time != 0) == (b % 3 == 0));
      CHECK(static_cast<int16_t*>(m->GetData())[5] == int16_t(b * 100 + 5));
      CHECK(m->GetDataBytes() == 4 * N);
      q.MessageProcessed(m);
    }
    CHECK(q.GetNextSamples() == nullptr && q.GetIsDone());
  }
  {
    // split int16 -> re block then im block; int8 and float pass through
    const uint32_t N = 32;
    SampleQueue q(SampleQueue::Short, 12, N, 4, false, false);
    q.SetDropFirstSweep(false);
    std::vector<int16_t> re(N), im(N);
    for (uint32_t i = 0; i < N; i++) { re[i] = int16_t(i); im[i] = int16_t(-int(i)); }
    q.AppendSamples(re.data(), im.data(), 5.0, 0);
    auto* m = q.GetNextSamples();
    const int16_t* d = static_cast<const int16_t*>(m->GetData());
    CHECK(d[3] == 3 && d[N + 3] == -3);
    q.MessageProcessed(m);
    q.SetIsDone();
  }
  {
    // bounded: the producer blocks at bufferCount and resumes as the consumer drains; batch multiples
    const uint32_t N = 16;
    SampleQueue q(SampleQueue::ByteComplex, 8, N, 4, true, false);
    q.SetDropFirstSweep(false);
    std::atomic<int> appended{0};
    std::thread producer([&] {
      std::vector<int8_t> buf(2 * N, 1);
      for (int b = 0; b < 10; b++) { q.AppendSamples(reinterpret_cast<int8_t(*)[2]>(buf.data()), b, 0); appended++; }
      q.SetIsDone();
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
    CHECK(appended == 4);                                 // queue depth 4 reached, producer parked
    std::vector<SampleQueue::MessageType*> batch;
    uint64_t next = 0; uint32_t total = 0;
    while (uint32_t n = q.GetNextBatch(batch, 4, 2)) {
      if (!q.GetIsDone()) CHECK(n % 2 == 0);
      for (auto* m : batch) { CHECK(m->GetHeader().m_sequenceId == next++); q.MessageProcessed(m); }
      total += n;
    }
    producer.join();
    CHECK(total == 10);
    CHECK(!q.ReceivedAck()); q.SendAck(); CHECK(q.ReceivedAck()); q.ClearAck(); CHECK(!q.ReceivedAck());
  }

  {
    // a consumer asleep waiting for a GROUP of buffers (K-FFT averaging) is woken when the group completes, not
    // only on the empty -> non-empty edge (regression: it used to sleep forever once it had seen 1 < K buffers)
    const uint32_t N = 16;
    SampleQueue q(SampleQueue::ByteComplex, 8, N, 64, true, false);
    q.SetDropFirstSweep(false);
    std::thread producer([&] {
      std::vector<int8_t> buf(2 * N, 1);
      for (int b = 0; b < 8; b++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(15));     // the consumer is asleep again by now
        q.AppendSamples(reinterpret_cast<int8_t(*)[2]>(buf.data()), b, 0);
      }
    });
    std::vector<SampleQueue::MessageType*> batch;
    CHECK(q.GetNextBatch(batch, 4, 4) == 4);
    q.MessageProcessed(batch);
    CHECK(q.GetNextBatch(batch, 8, 4) == 4);
    q.MessageProcessed(batch);
    producer.join();
    q.SetIsDone();
    CHECK(q.GetNextBatch(batch, 8, 4) == 0);
  }

  // ---- SyntheticSource: deterministic, kind layouts, drives a queue through sweeps
  {
    const uint32_t N = 256;
    SyntheticSource a(SampleQueue::ByteComplex, 8, 42, 2, 20000000, N, 2.4e9, 2.45e9);
    SyntheticSource b(SampleQueue::ByteComplex, 8, 42, 2, 20000000, N, 2.4e9, 2.45e9);
    std::vector<int8_t> x(2 * N), y(2 * N), z(2 * N);
    a.Generate(1, 2, 0, x.data()); b.Generate(1, 2, 0, y.data()); a.Generate(1, 2, 1, z.data());
    CHECK(memcmp(x.data(), y.data(), 2 * N) == 0 && memcmp(x.data(), z.data(), 2 * N) != 0);
    CHECK(a.GetFrequencyCount() == 3);
    SampleQueue q(SampleQueue::ByteComplex, 8, N, 64, true, false);
    a.Start();
    a.StartStreaming(3, q);                             // 3 sweeps x 3 steps x 2 buffers, first sweep dropped
    uint32_t got = 0, starts = 0;
    while (auto* m = q.GetNextSamples()) { got++; starts += m->GetHeader().m_time != 0; q.MessageProcessed(m); }
    a.Join();
    CHECK(got == 12 && starts == 2 && q.GetDroppedCount() == 6);
  }

  // ---- SampleBuffer + visitor protocol (buffer.cpp:360-372: one Begin, one Process per block, one End)
  {
    const uint32_t N = 128;
    SampleBuffer sb(SampleBuffer::ShortComplex, 12, N, 8);
    std::vector<int16_t> buf(2 * N, 7);
    for (int b = 0; b < 5; b++) sb.AppendSamples(reinterpret_cast<int16_t(*)[2]>(buf.data()), 100.0 + b);
    sb.SetIsDone();
    CountingVisitor v;
    std::vector<double> f;
    CHECK(sb.GetNextSamples(&v, f, 3) == 3 && v.begins == 1 && v.ends == 1 && v.blocks == 3 && v.seq == 0);
    CHECK(v.total == 3 * 4 * N && f.size() == 3 && f[2] == 102.0);
    CHECK(sb.GetNextSamples(&v, f, 3) == 2 && v.seq == 3 * N && f[0] == 103.0);
    CHECK(sb.GetNextSamples(&v, f, 3) == 0);
  }
  // ---- SampleBuffer: the reference's own drain signature (sampleBuffer.h:44) on fc32 buffers
  {
    const uint32_t N = 64;
    SampleBuffer sb(SampleBuffer::FloatComplex, 0, N, 4);
    std::vector<float> in(2 * N), out(2 * N);
    for (int b = 0; b < 3; b++) {
      for (uint32_t i = 0; i < 2 * N; i++) in[i] = float(b * 100 + int(i));
      sb.AppendSamples(reinterpret_cast<fftwf_complex*>(in.data()), 1e6 * (b + 1));
    }
    sb.SetIsDone();
    double f = 0;
    for (int b = 0; b < 3; b++) {
      CHECK(sb.GetNextSamples(reinterpret_cast<fftwf_complex*>(out.data()), f));
      CHECK(f == 1e6 * (b + 1) && out[9] == float(b * 100 + 9));
    }
    CHECK(!sb.GetNextSamples(reinterpret_cast<fftwf_complex*>(out.data()), f));
  }

  // ---- triggered recording (messageQueue.h:98-139, 259-288) on fc32 messages (no conversion, no GPU)
  {
    const uint32_t N = 256;
    char path[] = "/tmp/scn_selftest_rec_XXXXXX";
    int fd = mkstemp(path);
    CHECK(fd >= 0);
    close(fd);
    std::vector<float> buf(2 * N);
    {
      SampleQueue q(SampleQueue::FloatComplex, 0, N, 40, false, true);     // history = 4 messages
      q.SetDropFirstSweep(false);
      for (int b = 0; b < 12; b++) {
        for (uint32_t i = 0; i < 2 * N; i++) buf[i] = float(b * 1000 + int(i));
        q.AppendSamples(reinterpret_cast<fftwf_complex*>(buf.data()), 1e6, 0);
      }
      q.SetIsDone();
      std::vector<SampleQueue::MessageType*> held;
      for (int b = 0; b < 12; b++) {
        SampleQueue::MessageType* m = q.GetNextSamples();
        CHECK(m != nullptr && m->GetHeader().m_sequenceId == uint64_t(b));
        if (b == 5) q.BeginWrite(3, path);                                   // trigger at 5, pre-trigger 2
        if (b == 9) q.EndWrite(9);
        // workers may finish out of order: 6 is parked after 7
        if (b == 6) { held.push_back(m); continue; }
        q.MessageProcessed(m);
        if (b == 7) { q.MessageProcessed(held[0]); held.clear(); }
      }
      CHECK(q.GetNextSamples() == nullptr);
    }                                                                        // destructor joins the writer
    std::vector<char> rec = ReadAll(path);
    CHECK(rec.size() == size_t(6) * N * 8);                                  // messages 3..8
    const float* f = reinterpret_cast<const float*>(rec.data());
    for (int k = 0; k < 6; k++) CHECK(f[size_t(k) * 2 * N + 7] == float((3 + k) * 1000 + 7));
    // a start that has already left the history is skipped forward, and an open window is flushed at shutdown
    {
      SampleQueue q(SampleQueue::FloatComplex, 0, N, 20, false, true);     // history = 2 messages
      q.SetDropFirstSweep(false);
      for (int b = 0; b < 8; b++) {
        for (uint32_t i = 0; i < 2 * N; i++) buf[i] = float(b * 1000 + int(i));
        q.AppendSamples(reinterpret_cast<fftwf_complex*>(buf.data()), 1e6, 0);
      }
      q.SetIsDone();
      for (int b = 0; b < 8; b++) {
        SampleQueue::MessageType* m = q.GetNextSamples();
        if (b == 6) q.BeginWrite(1, path);                                   // 1..3 are gone: history holds 4, 5
        q.MessageProcessed(m);
      }
    }
    rec = ReadAll(path);
    CHECK(rec.size() == size_t(4) * N * 8);                                  // 4, 5, 6, 7
    f = reinterpret_cast<const float*>(rec.data());
    CHECK(f[3] == 4003.0f && f[size_t(3) * 2 * N + 3] == 7003.0f);
    // the writer is a thread of its own: with a limit it stays below the planned end of an OPEN window however far
    // the workers are ahead (without one it would have written 6..11 here before EndWrite arrived); LimitWrite extends
    for (int extend = 0; extend < 2; extend++) {
      {
        SampleQueue q(SampleQueue::FloatComplex, 0, N, 400, false, true);  // history = 40 messages
        q.SetDropFirstSweep(false);
        for (int b = 0; b < 12; b++) {
          for (uint32_t i = 0; i < 2 * N; i++) buf[i] = float(b * 1000 + int(i));
          q.AppendSamples(reinterpret_cast<fftwf_complex*>(buf.data()), 1e6, 0);
        }
        q.SetIsDone();
        for (int b = 0; b < 12; b++) {
          SampleQueue::MessageType* m = q.GetNextSamples();
          if (b == 4) q.BeginWrite(2, path, 6);                              // trigger at 4, pre 2, post 1: [2, 6)
          if (b == 7 && extend) q.LimitWrite(9);                             // a later trigger: [2, 9)
          q.MessageProcessed(m);
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(50));          // the writer has had every chance
        q.EndWrite(extend ? 9 : 6);
      }
      rec = ReadAll(path);
      const int want = extend ? 7 : 4;
      CHECK(rec.size() == size_t(want) * N * 8);
      f = reinterpret_cast<const float*>(rec.data());
      CHECK(f[5] == 2005.0f && f[size_t(want - 1) * 2 * N + 5] == float((2 + want - 1) * 1000 + 5));
    }
    // integer kinds without a converter must not write garbage: covered by the GPU tests (SetWriteConverter)
    remove(path);
  }

  printf("host_selftest ok\n");
  return 0;
}
