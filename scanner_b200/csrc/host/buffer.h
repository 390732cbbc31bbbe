// ProcessInterface -- the Begin / Process / End visitor of the reference's legacy ring API
// (buffer.h:9-24), and GpuProcessInterface, the visitor that turns a visited run of raw sample
// blocks into one fused GPU launch.  In the reference, CircularBuffer::ProcessItems drives a
// visitor with one Begin, one Process per contiguous block, one End (buffer.cpp:360-372) and the
// shipped visitor (CopyBufferProcessInterface, processInterface.cpp:63-90) memcpy's blocks into
// an output buffer; here the blocks are raw device samples staged in pinned memory and End()
// submits them through the C ABI.
#pragma once
#include <cstdint>
#include <vector>

#include "scanner_b200.h"

template <typename ElementType>
class ProcessInterface {
  bool m_doMergeRequests;
 public:
  explicit ProcessInterface(bool doMergeRequests) : m_doMergeRequests(doMergeRequests) {}
  virtual ~ProcessInterface() {}
  bool GetDoMergeRequests() { return m_doMergeRequests; }
  virtual void Begin(uint64_t sequenceId, uint32_t totalItemCount) = 0;
  virtual void Process(const ElementType* items, uint32_t count) = 0;
  virtual void End() = 0;
};

// Items are raw bytes of the context's sample kind; totalItemCount must be a whole number of
// spectra (averaging * buffer bytes each).  Errors follow the reference's convention for this
// layer: assert / message + exit(1) (processInterface.cpp:42-45,57), never exceptions.
class GpuProcessInterface : public ProcessInterface<uint8_t> {
 public:
  GpuProcessInterface(scn_ctx* ctx, uint32_t maxSpectra, uint32_t averaging, uint32_t sampleCount);
  ~GpuProcessInterface() override;
  void Begin(uint64_t sequenceId, uint32_t totalItemCount) override;
  void Process(const uint8_t* items, uint32_t count) override;
  void End() override;

  uint64_t GetSequenceId() const { return m_sequenceId; }
  uint32_t GetSpectrumCount() const { return m_spectra; }
  const std::vector<uint32_t>& GetHitCounts() const { return m_counts; }
  const std::vector<uint32_t>& GetHitMasks() const { return m_masks; }     // [spectrum][N/32]

 private:
  scn_ctx* m_ctx;
  uint32_t m_maxSpectra, m_averaging, m_sampleCount;
  size_t m_spectrumBytes;
  void* m_staging = nullptr;     // pinned
  uint64_t m_sequenceId = 0;
  uint32_t m_expected = 0, m_count = 0, m_spectra = 0;
  std::vector<uint32_t> m_counts, m_masks;
};
