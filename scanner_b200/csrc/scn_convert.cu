// scn_convert.cu -- the reference's sample converters as a standalone kernel.
//
// Utility::byte_complex_to_float_complex (utility.cpp:34-56), Utility::short_complex_to_float_complex
// for interleaved (utility.cpp:58-84) and split (utility.cpp:9-32) int16, and the fc32 pass-through of
// MessageQueue::AppendSamples (messageQueue.h:231-237): raw device samples -> fftwf_complex, bit for
// bit what the reference keeps in its queue messages and writes to its recording files
// (messageQueue.h:126-131).  The fused spectrum-sense kernel never materialises these floats; this
// kernel exists for the trigger/record path, which has to write them.
//
//   max = intK_t(1 << (enob-1)) (wraps: enob 8 -> -128), s = float(1.0 / max)            utility.cpp:14-15
//   dc  = int32(uint32(sum) / uint32(N)) per rail when DC correction is on (unsigned!)    utility.cpp:25-26
//   out = float(int(x) - dc) * s                                                          utility.cpp:28-31
// int -> float rounds to nearest-even exactly like the reference's cast (it only rounds at all on the
// unsigned-division quirk path, where |x - dc| can reach 2^24 + 2^15); the multiply by a signed power of
// two is exact.
//
// One CTA per buffer (grid-stride): 16-byte loads, int32 rail sums by packed dot products + block
// reduction, then the conversion re-reads the buffer (an L1/L2 hit: 2..4 bytes per sample) and writes
// 8 bytes per sample -- an HBM stream of B_in + 8 bytes per sample.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/scanner_b200.h"

namespace scn {

constexpr int kConvThreads = 256;

template <int KIND>
__global__ void __launch_bounds__(kConvThreads)
convert_kernel(const uint8_t* __restrict__ raw, float2* __restrict__ out, uint32_t n, uint32_t n_buffers,
               float onebymax, int correct_dc) {
  __shared__ int red[2][kConvThreads / 32];
  constexpr int kBytes = KIND == SCN_KIND_BYTE_COMPLEX ? 2 : 4;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (uint32_t b = blockIdx.x; b < n_buffers; b += gridDim.x) {
    const uint8_t* buf = raw + size_t(b) * n * kBytes;
    float2* dst = out + size_t(b) * n;
    int dci = 0, dcq = 0;
    if (correct_dc) {
      int si = 0, sq = 0;
      const uint32_t* w32 = reinterpret_cast<const uint32_t*>(buf);
      if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
        for (uint32_t i = t; i < n / 2; i += kConvThreads) {          // 2 samples per word: I Q I Q
          const int w = int(__ldg(w32 + i));
          si = __dp4a(w, 0x00010001, si);
          sq = __dp4a(w, 0x01000100, sq);
        }
      } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
        for (uint32_t i = t; i < n; i += kConvThreads) {              // 1 sample per word: I Q
          const int w = int(__ldg(w32 + i));
          si += int(short(w & 0xffff));
          sq += w >> 16;
        }
      } else {                                                        // split: n int16 I, then n int16 Q
        for (uint32_t i = t; i < n / 2; i += kConvThreads) {
          const int wi = int(__ldg(w32 + i)), wq = int(__ldg(w32 + n / 2 + i));
          si += int(short(wi & 0xffff)) + (wi >> 16);
          sq += int(short(wq & 0xffff)) + (wq >> 16);
        }
      }
      si = __reduce_add_sync(0xffffffffu, si);
      sq = __reduce_add_sync(0xffffffffu, sq);
      if (lane == 0) { red[0][warp] = si; red[1][warp] = sq; }
      __syncthreads();
      si = 0; sq = 0;
#pragma unroll
      for (int w = 0; w < kConvThreads / 32; w++) { si += red[0][w]; sq += red[1][w]; }
      dci = int(unsigned(si) / n);                                    // the reference divides unsigned (utility.cpp:25-26)
      dcq = int(unsigned(sq) / n);
      __syncthreads();                                                // red[] is reused by the next buffer
    }
    if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
      const unsigned short* s16 = reinterpret_cast<const unsigned short*>(buf);
      for (uint32_t i = t; i < n; i += kConvThreads) {
        const unsigned short w = __ldg(s16 + i);
        const int xi = int(int8_t(w & 0xff)), xq = int(int8_t(w >> 8));
        dst[i] = make_float2(__fmul_rn(__int2float_rn(xi - dci), onebymax), __fmul_rn(__int2float_rn(xq - dcq), onebymax));
      }
    } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
      const uint32_t* w32 = reinterpret_cast<const uint32_t*>(buf);
      for (uint32_t i = t; i < n; i += kConvThreads) {
        const int w = int(__ldg(w32 + i));
        const int xi = int(short(w & 0xffff)), xq = w >> 16;
        dst[i] = make_float2(__fmul_rn(__int2float_rn(xi - dci), onebymax), __fmul_rn(__int2float_rn(xq - dcq), onebymax));
      }
    } else {
      const short* re = reinterpret_cast<const short*>(buf);
      const short* im = re + n;
      for (uint32_t i = t; i < n; i += kConvThreads) {
        const int xi = int(__ldg(re + i)), xq = int(__ldg(im + i));
        dst[i] = make_float2(__fmul_rn(__int2float_rn(xi - dci), onebymax), __fmul_rn(__int2float_rn(xq - dcq), onebymax));
      }
    }
  }
}

cudaError_t launch_convert(int kind, const void* raw, float* out, uint32_t n, uint32_t n_buffers, float onebymax,
                           bool correct_dc, int num_sms, cudaStream_t stream) {
  if (n_buffers == 0) return cudaSuccess;
  if (kind == SCN_KIND_FLOAT_COMPLEX)                                 // messageQueue.h:231-237: stored as is
    return cudaMemcpyAsync(out, raw, size_t(n_buffers) * n * sizeof(float2), cudaMemcpyDeviceToDevice, stream);
  uint32_t grid = uint32_t(num_sms) * 8;
  if (grid > n_buffers) grid = n_buffers;
  const uint8_t* r = static_cast<const uint8_t*>(raw);
  float2* o = reinterpret_cast<float2*>(out);
  const int dc = correct_dc ? 1 : 0;
  switch (kind) {
    case SCN_KIND_BYTE_COMPLEX:
      convert_kernel<SCN_KIND_BYTE_COMPLEX><<<grid, kConvThreads, 0, stream>>>(r, o, n, n_buffers, onebymax, dc); break;
    case SCN_KIND_SHORT_COMPLEX:
      convert_kernel<SCN_KIND_SHORT_COMPLEX><<<grid, kConvThreads, 0, stream>>>(r, o, n, n_buffers, onebymax, dc); break;
    case SCN_KIND_SHORT:
      convert_kernel<SCN_KIND_SHORT><<<grid, kConvThreads, 0, stream>>>(r, o, n, n_buffers, onebymax, dc); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace scn
