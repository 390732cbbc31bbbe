#include "frequencyTable.h"

#include <cassert>
#include <cstdio>

#include "scanner_b200.h"

FrequencyTable::FrequencyTable(uint32_t sampleRate, double startFrequency, double stopFrequency,
                               double useBandWidth, double dcIgnoreWidth, bool printTable) {
  const uint32_t count = scn_frequency_table(sampleRate, startFrequency, stopFrequency, useBandWidth,
                                             dcIgnoreWidth, nullptr, 0);
  std::vector<double> f(count);
  scn_frequency_table(sampleRate, startFrequency, stopFrequency, useBandWidth, dcIgnoreWidth, f.data(), count);
  table_.reserve(count);
  for (uint32_t i = 0; i < count; i++) {
    if (printTable) printf("Frequency %d: %.0f\n", i, f[i]);   // the reference dumps its table (frequencyTable.cpp:34)
    table_.push_back(Entry{f[i], nullptr});
  }
}

double FrequencyTable::GetNextFrequency(void** pinfo) {
  if (++index_ >= table_.size()) {
    index_ = 0;
    sweeps_++;
  }
  return GetCurrentFrequency(pinfo);
}

double FrequencyTable::GetCurrentFrequency(void** pinfo) {
  const Entry& e = table_[index_];
  if (pinfo) *pinfo = e.info;
  return e.frequency;
}

uint32_t FrequencyTable::GetFrequencyCount() { return uint32_t(table_.size()); }

double FrequencyTable::GetFrequencyFromIndex(uint32_t index) {
  assert(index < table_.size());
  return table_[index].frequency;
}

void FrequencyTable::SetFrequencyInfoForIndex(uint32_t index, void* info) {
  assert(index < table_.size());
  table_[index].info = info;
}

uint32_t FrequencyTable::GetIterationCount() { return sweeps_; }
bool FrequencyTable::GetIsScanStart() { return index_ == 0; }
double FrequencyTable::GetStartFrequency() { return table_.front().frequency; }
double FrequencyTable::GetStopFrequency() { return table_.back().frequency; }
