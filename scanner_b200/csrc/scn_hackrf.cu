// scn_hackrf.cu -- HackRF sweep-frame pre-pass on device.
//
// In sweep mode the HackRF firmware starts every 16 384-byte block of a transfer with a 10-byte
// frame header (0x7F 0x7F + little-endian uint64 tuned frequency) that overwrites the first five
// IQ samples.  The reference's rx callback parses the header for the centre frequency and patches
// those samples before the transfer is cut into FFT buffers (HackRFSource::interpolateSamples,
// hackRFSource.cpp:186-222).  A captured sweep stream that is already in HBM gets the same
// treatment here, in place, so the fused kernel can consume the transfer as N-sample int8 buffers
// without the host touching the sample bytes.
//
// The reference's loop is reproduced as written, quirks included: it runs valid_length/2/8192
// times but never advances `ubuf` (:192), so every iteration looks at the FIRST block; iterations
// after the first only act when the patched bytes themselves read 0x7F 0x7F (sample 5 saturated).
// The patch value is sample 5, averaged for i > 0 with sample i-1 in int arithmetic (truncation
// toward zero, :209-210) and narrowed to int8.  One thread per transfer: the work is ~12 bytes and
// a data-dependent chain of at most 16 iterations -- nothing to parallelise inside a transfer.
#include <cuda_runtime.h>
#include <stdint.h>

namespace scn {

__global__ void __launch_bounds__(128)
hackrf_sweep_prepass_kernel(uint8_t* __restrict__ transfers, uint32_t n_transfers, uint32_t valid_length,
                            unsigned long long* __restrict__ frequency_hz, uint32_t* __restrict__ status) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_transfers) return;
  uint8_t* ubuf = transfers + size_t(t) * valid_length;
  const uint32_t count = valid_length / 2;
  // head of the transfer in registers: marker, frequency, sample 5
  uint8_t head[12];
#pragma unroll
  for (int k = 0; k < 12; k++) head[k] = ubuf[k];
  unsigned long long freq = 0;
  uint32_t st = 0;
  bool dirty = false;
  for (uint32_t i = 0; i < count; i += 8192) {
    if (head[0] != 0x7F || head[1] != 0x7F) break;          // nothing can change head[0..1] any more
    unsigned long long cur = 0;
#pragma unroll
    for (int k = 7; k >= 0; k--) cur = (cur << 8) | head[2 + k];
    if (freq != 0 && freq != cur) st += 1u << 8;            // the reference prints a mismatch line here (:202-206)
    freq = cur;
    st |= 1u;
    int p0 = int(int8_t(head[10])), p1 = int(int8_t(head[11]));
    if (i > 0) {
      p0 = int(int8_t((p0 + int(int8_t(ubuf[2 * (i - 1)]))) / 2));
      p1 = int(int8_t((p1 + int(int8_t(ubuf[2 * (i - 1) + 1]))) / 2));
    }
#pragma unroll
    for (int j = 0; j < 5; j++) { head[2 * j] = uint8_t(p0); head[2 * j + 1] = uint8_t(p1); }
    dirty = true;
  }
  if (dirty) {
#pragma unroll
    for (int k = 0; k < 10; k++) ubuf[k] = head[k];
  }
  if (frequency_hz) frequency_hz[t] = freq;
  if (status) status[t] = st;
}

cudaError_t launch_hackrf_prepass(void* transfers, uint32_t n_transfers, uint32_t valid_length,
                                  uint64_t* frequency_hz, uint32_t* status, cudaStream_t stream) {
  if (n_transfers == 0) return cudaSuccess;
  hackrf_sweep_prepass_kernel<<<(n_transfers + 127) / 128, 128, 0, stream>>>(
      static_cast<uint8_t*>(transfers), n_transfers, valid_length,
      reinterpret_cast<unsigned long long*>(frequency_hz), status);
  return cudaGetLastError();
}

}  // namespace scn
