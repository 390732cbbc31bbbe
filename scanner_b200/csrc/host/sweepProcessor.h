// SweepProcessor -- the multi-GPU consumer: one worker thread, one scn_ctx and one CUDA stream set per GPU, inside
// ONE process.  It replaces what the reference does with two worker threads sharing one FFT
// object (ProcessSamples::StartProcessing, process.cpp:316-331):
//
//   * the frequency table's retune steps are split into contiguous ranges, one per GPU (scn_shard_steps); a queue
//     message goes to the GPU that owns the step of its centre frequency, so no sample ever crosses between GPUs;
//   * every GPU runs the same batched submit/collect loop as ProcessSamples::ThreadWorker and produces, per message
//     in sequence order, exactly the lines the reference prints (process.cpp:57,285); an ordered emitter interleaves
//     the GPUs' output back into queue order, so the text equals the single-GPU (and the reference's) output;
//   * per retune step every GPU keeps a small record (hit total, spectra, OR of the hit masks -- scn_records.cu's
//     layout); at the end of every sweep (the next scan-start marker, messageQueue.h:67-72) the partial tables are
//     exchanged over NVLink -- ncclCommInitAll + one grouped ncclAllGather (scn_nccl_gather_*), or the peer-memory
//     windows of scn_exchange_* -- and merged, so every GPU (and the host) holds the whole sweep's records.
//
// Only the C ABI is used (no CUDA headers here), which is also what lets tests/mock_abi run this class on a CPU.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sampleQueue.h"
#include "scanner_b200.h"

class SweepProcessor {
 public:
  enum Exchange { NcclAllGather, PeerMemory };

  // stepFrequencies: the centre frequencies of the FrequencyTable (frequencyTable.cpp:17-36), in table order;
  // devices: CUDA ordinals, one worker each.  Remaining arguments as ProcessSamples (process.h:74-85).
  SweepProcessor(uint32_t numSamples, uint32_t sampleRate, uint32_t enob, float threshold, int windowType,
                 const std::vector<double>& stepFrequencies, const std::vector<int>& devices,
                 Exchange exchange = NcclAllGather, double useBandWidth = 0.75);
  ~SweepProcessor();

  void SetAveraging(uint32_t k) { m_averaging = k ? k : 1; }
  void SetMaxBatch(uint32_t maxBatch) { m_maxBatch = maxBatch ? maxBatch : 1; }
  void SetOutput(FILE* out) { m_out = out; }                       // nullptr silences printing
  // one line per retune step after every sweep, from the MERGED records: "sweep <n> step <i> freq <Hz> spectra <c> hits <h>"
  void SetSweepReport(bool on) { m_sweepReport = on; }

  // Drains the queue (blocks until the source is done), like ProcessSamples::StartProcessing.
  bool StartProcessing(SampleQueue& sampleQueue);

  uint32_t GetDeviceCount() const { return uint32_t(m_devices.size()); }
  uint32_t GetStepOwner(uint32_t step) const;
  uint64_t GetBuffersProcessed() const { return m_buffersProcessed; }
  uint64_t GetHitCount() const { return m_hitCount; }
  uint64_t GetLaunchCount() const { return m_launches; }
  uint32_t GetSweepCount() const { return m_sweepsDone; }
  uint64_t GetBuffersOnDevice(uint32_t d) const { return m_perDevice[d]; }
  // merged records of the last completed sweep: n_steps x (N/32 + 2) words, identical on every GPU after the exchange
  const std::vector<uint32_t>& GetLastSweepRecords() const { return m_lastRecords; }

 private:
  struct Item {                       // one spectrum (K consecutive messages of one step), or a sweep-end marker
    std::vector<SampleQueue::MessageType*> msgs;
    uint64_t emit = 0;                // position of its output in queue order
    uint32_t step = 0;
    bool sweepEnd = false;
    bool stop = false;
  };
  struct Inbox {
    std::mutex mutex;
    std::condition_variable ready;
    std::deque<Item> items;
  };

  void Worker(uint32_t d);
  void Push(uint32_t d, Item&& item);
  void Emit(uint64_t index, std::string&& text);
  void SweepArrive(uint32_t d, uint64_t reportEmit);
  void ExchangeAndReport(uint64_t reportEmit);
  uint32_t StepOf(double frequency) const;
  scn_ctx* CreateContext(int device, uint32_t maxSpectra, uint32_t hitCap);

  uint32_t m_sampleCount, m_sampleRate, m_enob;
  float m_threshold;
  std::vector<double> m_stepFrequencies;
  std::map<double, uint32_t> m_stepIndex;
  std::vector<int> m_devices;
  Exchange m_exchange;
  uint32_t m_useWindow, m_words, m_recWords;
  std::vector<float> m_window;
  uint32_t m_averaging = 1, m_maxBatch = 1024;
  FILE* m_out = stdout;
  bool m_sweepReport = false;
  SampleQueue* m_queue = nullptr;

  std::vector<std::unique_ptr<Inbox>> m_inbox;
  std::vector<std::thread> m_workers;
  std::vector<std::vector<uint32_t>> m_partial;          // per device: n_steps x record words, current sweep
  std::vector<uint32_t> m_lastRecords;
  // sweep barrier: the last worker to arrive runs the exchange for all devices
  std::mutex m_sweepMutex;
  std::condition_variable m_sweepCv;
  uint32_t m_sweepArrived = 0;
  uint64_t m_sweepGeneration = 0;
  scn_gather* m_gather = nullptr;
  std::vector<scn_exchange*> m_windows;
  // ordered output
  std::mutex m_emitMutex;
  std::map<uint64_t, std::string> m_pending;
  uint64_t m_nextEmit = 0;

  std::atomic<uint64_t> m_buffersProcessed{0}, m_hitCount{0}, m_launches{0};
  std::vector<std::atomic<uint64_t>> m_perDevice;
  std::atomic<uint32_t> m_sweepsDone{0};
};
