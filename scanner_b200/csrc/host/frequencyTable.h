// Sweep plan: the list of centre frequencies a source retunes through.
// Same public interface as the reference's FrequencyTable (frequencyTable.h:3-29); the table
// arithmetic is scn_frequency_table() (frequencyTable.cpp:17-36).  The retune steps of this
// table are the unit that shards across GPUs.
#pragma once
#include <cstdint>
#include <vector>

class FrequencyTable {
 public:
  FrequencyTable(uint32_t sampleRate, double startFrequency, double stopFrequency,
                 double useBandWidth, double dcIgnoreWidth, bool printTable = true);
  double GetNextFrequency(void** pinfo = nullptr);       // advances; wraps and counts sweeps
  double GetCurrentFrequency(void** pinfo = nullptr);
  uint32_t GetFrequencyCount();
  double GetFrequencyFromIndex(uint32_t index);
  void SetFrequencyInfoForIndex(uint32_t index, void* info);
  uint32_t GetIterationCount();
  bool GetIsScanStart();
  double GetStartFrequency();
  double GetStopFrequency();
  uint32_t GetCurrentIndex() const { return index_; }

 private:
  struct Entry { double frequency; void* info; };
  std::vector<Entry> table_;
  uint32_t index_ = 0;
  uint32_t sweeps_ = 0;
};
