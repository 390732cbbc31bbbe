#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "buffer.h"

GpuProcessInterface::GpuProcessInterface(scn_ctx* ctx, uint32_t maxSpectra, uint32_t averaging,
                                         uint32_t sampleCount)
    : ProcessInterface<uint8_t>(false), m_ctx(ctx), m_maxSpectra(maxSpectra),
      m_averaging(averaging ? averaging : 1), m_sampleCount(sampleCount),
      m_spectrumBytes(scn_buffer_bytes(ctx) * (averaging ? averaging : 1)) {
  if (scn_alloc_pinned(m_spectrumBytes * maxSpectra, &m_staging) != SCN_OK) {
    fprintf(stderr, "GpuProcessInterface: %s\n", scn_last_error());
    exit(1);
  }
  m_counts.resize(maxSpectra);
  m_masks.resize(size_t(maxSpectra) * scn_mask_words(ctx));
}

GpuProcessInterface::~GpuProcessInterface() { scn_free_pinned(m_staging); }

void GpuProcessInterface::Begin(uint64_t sequenceId, uint32_t totalItemCount) {
  assert(totalItemCount % m_spectrumBytes == 0);
  assert(totalItemCount / m_spectrumBytes <= m_maxSpectra);
  m_sequenceId = sequenceId;
  m_expected = totalItemCount;
  m_count = 0;
  m_spectra = 0;
}

void GpuProcessInterface::Process(const uint8_t* items, uint32_t count) {
  assert(m_count + count <= m_expected);
  memcpy(static_cast<uint8_t*>(m_staging) + m_count, items, count);
  m_count += count;
}

void GpuProcessInterface::End() {
  assert(m_expected == m_count);
  m_spectra = uint32_t(m_count / m_spectrumBytes);
  if (m_spectra) {
    uint32_t ticket = 0;
    if (scn_submit(m_ctx, m_staging, m_spectra, &ticket) != SCN_OK ||
        scn_collect(m_ctx, ticket, nullptr, m_masks.data(), m_counts.data(), nullptr, nullptr) != SCN_OK) {
      fprintf(stderr, "GpuProcessInterface: %s\n", scn_last_error());
      exit(1);
    }
  }
  m_count = m_expected = 0;
}
