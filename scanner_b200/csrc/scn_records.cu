// scn_records.cu -- per-retune-step detection records.
//
// The reference reports detections per buffer as it goes (process.cpp:46-61).  A batched,
// multi-GPU sweep wants one small record per retune step of the FrequencyTable
// (frequencyTable.cpp:17-36) instead: how many bins triggered during the dwell and which bins
// ever triggered.  These records are what ranks exchange over NCCL;
// the raw IQ and the spectra never leave the GPU that owns the step.
//
// Record layout (uint32 words): [0] hit total, [1] spectra that contributed, [2 .. 2+W) OR of the
// hit masks, W = N/32.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/scanner_b200.h"

namespace scn {

constexpr int kRecThreads = 256;

// grid = (n_steps, splits).  Spectra are indexed globally in step-major order: spectrum u belongs
// to step u / units_per_step; this rank holds units [first_unit, first_unit + n_spectra).
__global__ void __launch_bounds__(kRecThreads)
summarize_steps_kernel(const uint32_t* __restrict__ masks, const uint32_t* __restrict__ counts,
                       uint32_t n_spectra, uint64_t first_unit, uint32_t units_per_step, uint32_t words,
                       uint32_t* __restrict__ records) {
  const uint32_t step = blockIdx.x;
  const uint64_t g0 = uint64_t(step) * units_per_step, g1 = g0 + units_per_step;
  const uint64_t lo = g0 > first_unit ? g0 : first_unit;
  const uint64_t hi_all = first_unit + n_spectra;
  const uint64_t hi = g1 < hi_all ? g1 : hi_all;
  if (lo >= hi) return;
  const uint32_t n_local = uint32_t(hi - lo);
  const uint32_t s_begin = uint32_t(lo - first_unit);
  // this CTA's slice of the step's spectra
  const uint32_t per = (n_local + gridDim.y - 1) / gridDim.y;
  const uint32_t a = blockIdx.y * per;
  if (a >= n_local) return;
  const uint32_t b = (a + per < n_local) ? a + per : n_local;
  uint32_t* rec = records + size_t(step) * (words + 2);

  const uint32_t rows = kRecThreads / 32;           // one warp walks one spectrum's words at a time
  const uint32_t lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  uint32_t hit_sum = 0;
  for (uint32_t w0 = 0; w0 < words; w0 += 32) {
    const uint32_t w = w0 + lane;
    uint32_t acc = 0;
    if (w < words)
      for (uint32_t s = a + row; s < b; s += rows) acc |= __ldg(masks + size_t(s_begin + s) * words + w);
    if (acc) atomicOr(rec + 2 + w, acc);
  }
  for (uint32_t s = a + threadIdx.x; s < b; s += kRecThreads) hit_sum += __ldg(counts + s_begin + s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hit_sum += __shfl_xor_sync(0xffffffffu, hit_sum, o);
  if (lane == 0 && hit_sum) atomicAdd(rec, hit_sum);
  if (threadIdx.x == 0) atomicAdd(rec + 1, b - a);
}

// out[step] = merge over parts: sums for words 0,1; OR for the mask words.
__global__ void merge_records_kernel(const uint32_t* __restrict__ parts, uint32_t n_parts, uint32_t n_steps,
                                     uint32_t rec_words, uint32_t* __restrict__ out) {
  const uint32_t total = n_steps * rec_words;
  for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < total; x += gridDim.x * blockDim.x) {
    const bool is_sum = (x % rec_words) < 2;
    uint32_t v = 0;
    for (uint32_t p = 0; p < n_parts; p++) {
      const uint32_t y = __ldg(parts + size_t(p) * total + x);
      v = is_sum ? v + y : (v | y);
    }
    out[x] = v;
  }
}

cudaError_t launch_summarize(const uint32_t* masks, const uint32_t* counts, uint32_t n_spectra,
                             uint64_t first_unit, uint32_t units_per_step, uint32_t n_steps, uint32_t words,
                             uint32_t* records, int num_sms, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(records, 0, sizeof(uint32_t) * size_t(n_steps) * (words + 2), stream);
  if (e != cudaSuccess) return e;
  if (n_spectra == 0) return cudaSuccess;
  // CTAs per step: spread 4 x SMs CTAs over the steps this batch actually TOUCHES (a rank of an 8-GPU sweep holds
  // ~1/8 of the steps: dividing by n_steps left 7 x 11 CTAs to read all the masks, and the kernel -- 15 us on one
  // GPU -- took ~100 us per batch at 8 GPUs, which was most of the weak-scaling loss the driver measured).
  const uint64_t last_unit = first_unit + n_spectra - 1;
  uint32_t touched = uint32_t(last_unit / units_per_step - first_unit / units_per_step) + 1;
  if (touched > n_steps) touched = n_steps;
  uint32_t splits = uint32_t(4 * num_sms) / (touched ? touched : 1);
  if (splits < 1) splits = 1;
  const uint32_t max_useful = (units_per_step + 63) / 64;
  if (splits > max_useful) splits = max_useful ? max_useful : 1;
  summarize_steps_kernel<<<dim3(n_steps, splits), kRecThreads, 0, stream>>>(
      masks, counts, n_spectra, first_unit, units_per_step, words, records);
  return cudaGetLastError();
}

cudaError_t launch_merge(const uint32_t* parts, uint32_t n_parts, uint32_t n_steps, uint32_t rec_words,
                         uint32_t* out, cudaStream_t stream) {
  const uint32_t total = n_steps * rec_words;
  if (total == 0) return cudaSuccess;
  const uint32_t grid = (total + 255) / 256;
  merge_records_kernel<<<grid < 1024 ? grid : 1024, 256, 0, stream>>>(parts, n_parts, n_steps, rec_words, out);
  return cudaGetLastError();
}

}  // namespace scn
