// scn_large.cu -- N = 2^15 and 2^16: one transform no longer fits one CTA (N/16 > 1024 threads), so
// the transform is split four-step style, N = 16 x N2, through an HBM/L2-resident intermediate:
//
//   X[k1 + 16 k2] = sum_n2 W_N2^(n2 k2) * [ W_N^(n2 k1) * sum_n1 x[n1 N2 + n2] W_16^(n1 k1) ]
//
//   dc_sums_kernel      (DC on)  per-buffer int32 sums -> dc = int32(uint32(sum) / N)      utility.cpp:44-50
//   columns_kernel      convert + window (utility.cpp:52-55, process.cpp:28-34), radix-16 over n1,
//                       twiddle W_N^(n2 k1), write Y[buffer][k1][n2] (fp32 complex, coalesced)
//   spectrum_sense_kernel<log2 N2, fp32 IQ> in row mode: N2-point FFT of every row, |X|^2, K-average,
//                       write power P[spectrum][k1][k2]
//   finalize_kernel     dB (utility.cpp:91-97), bin k = k1 + 16 k2, detection (process.cpp:46-61), mask,
//                       count, ordered hit records
//
// Same results contract as the in-CTA family (tests/test_gpu_large.py); ~30 B/sample of extra HBM traffic,
// so this path is bandwidth-bound well below the fused kernel -- it exists for completeness of the
// 256..65536 size range (BASELINE.json configs[4]).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/scanner_b200.h"
#include "scn_fft.cuh"

namespace scn {

constexpr float kDbPerLog2L = 1.5051499783199060f;

// ---- per-buffer DC ---------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) dc_sums_kernel(const uint8_t* __restrict__ raw, uint32_t n_buffers,
                                                      uint32_t n, int2* __restrict__ dcs) {
  __shared__ int s_red[2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kBytes = (KIND == SCN_KIND_BYTE_COMPLEX) ? 2 : 4;
  for (uint32_t b = blockIdx.x; b < n_buffers; b += gridDim.x) {
    const uint8_t* buf = raw + size_t(b) * n * kBytes;
    int si = 0, sq = 0;
    const int4* v4 = reinterpret_cast<const int4*>(buf);
    const uint32_t n16 = n * kBytes / 16;
    for (uint32_t i = tid; i < n16; i += 256) {
      const int4 v = __ldg(v4 + i);
      const int ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
          si = __dp4a(ws[k], 0x00010001, si);
          sq = __dp4a(ws[k], 0x01000100, sq);
        } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
          si = __dp2a_lo(ws[k], 0x00000001, si);
          sq = __dp2a_lo(ws[k], 0x00000100, sq);
        } else {   // split: first half of the buffer is I, second half Q
          const int s2 = __dp2a_lo(ws[k], 0x00000101, 0);
          if (i < n16 / 2) si += s2; else sq += s2;
        }
      }
    }
    si = __reduce_add_sync(0xffffffffu, si);
    sq = __reduce_add_sync(0xffffffffu, sq);
    if (lane == 0) { s_red[0][warp] = si; s_red[1][warp] = sq; }
    __syncthreads();
    if (tid == 0) {
      int ti = 0, tq = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) { ti += s_red[0][i]; tq += s_red[1][i]; }
      dcs[b] = make_int2(int(unsigned(ti) / n), int(unsigned(tq) / n));   // the unsigned-division quirk
    }
    __syncthreads();
  }
}

// ---- columns: convert + window + radix-16 over n1 + twiddle ------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) columns_kernel(const uint8_t* __restrict__ raw, uint32_t n_buffers,
                                                      uint32_t n2_count /* N2 */, const float* __restrict__ window,
                                                      const float2* __restrict__ wn /* exp(-2 pi i m / N) */,
                                                      const int2* __restrict__ dcs /* nullable */,
                                                      float2* __restrict__ y) {
  const uint32_t n = n2_count * 16;
  const uint32_t per_buf = n2_count / 256;                    // CTAs per buffer
  const uint32_t total = n_buffers * per_buf;
  for (uint32_t w = blockIdx.x; w < total; w += gridDim.x) {
    const uint32_t b = w / per_buf;
    const uint32_t n2 = (w - b * per_buf) * 256 + threadIdx.x;
    float2 v[kPts];
    if constexpr (KIND == SCN_KIND_FLOAT_COMPLEX) {
      const float2* src = reinterpret_cast<const float2*>(raw) + size_t(b) * n;
#pragma unroll
      for (int r = 0; r < 16; r++) {
        const float2 x = __ldg(src + n2 + r * n2_count);
        const float wv = __ldg(window + n2 + r * n2_count);
        v[r] = make_float2(__fmul_rn(x.x, wv), __fmul_rn(x.y, wv));
      }
    } else {
      const int2 dc = dcs ? dcs[b] : make_int2(0, 0);
#pragma unroll
      for (int r = 0; r < 16; r++) {
        const uint32_t s = n2 + r * n2_count;
        int xi, xq;
        if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
          const unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(raw) + size_t(b) * n + s);
          xi = int(static_cast<signed char>(u & 0xff));
          xq = int(static_cast<signed char>(u >> 8));
        } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
          const unsigned u = __ldg(reinterpret_cast<const unsigned*>(raw) + size_t(b) * n + s);
          xi = int(static_cast<short>(u & 0xffff));
          xq = int(static_cast<short>(u >> 16));
        } else {
          const short* p = reinterpret_cast<const short*>(raw) + size_t(b) * n * 2;
          xi = int(__ldg(p + s));
          xq = int(__ldg(p + n + s));
        }
        const float wv = __ldg(window + s);      // onebymax folded in (exact)
        v[r] = make_float2(__fmul_rn(float(xi - dc.x), wv), __fmul_rn(float(xq - dc.y), wv));
      }
    }
    dft16(v);                                     // v[k1] = sum_n1 x[n1 N2 + n2] W_16^(n1 k1)
    float2* dst = y + (size_t(b) * 16) * n2_count + n2;
    dst[0] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < 16; k1++) dst[size_t(k1) * n2_count] = cmul(v[k1], __ldg(wn + n2 * k1));
  }
}

// ---- finalize: dB, detection, mask, count, ordered hit records ----------------------------------------------
// P layout [spectrum][k1][k2]; output bin k = k1 + 16 k2.  One CTA per spectrum.
__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ power, uint32_t n_spectra,
                                                       uint32_t n2_count, float threshold, uint32_t use_window,
                                                       uint32_t dc_ignore, float* __restrict__ spectra,
                                                       uint32_t* __restrict__ masks, uint32_t* __restrict__ counts,
                                                       scn_hit* __restrict__ hits, uint32_t hit_cap) {
  extern __shared__ uint32_t smem_words[];        // [words] mask, then [words] exclusive prefix
  __shared__ uint32_t s_scan[256];
  const uint32_t n = n2_count * 16, half = n / 2, words = n / 32;
  uint32_t* smask = smem_words;
  uint32_t* sprefix = smem_words + words;
  const int tid = threadIdx.x;
  for (uint32_t s = blockIdx.x; s < n_spectra; s += gridDim.x) {
    const float* P = power + size_t(s) * n;
    // phase 1: dB out, mask words into smem
    for (uint32_t k2 = tid; k2 < n2_count; k2 += 256) {
      float db[16];
      uint32_t bits = 0;
#pragma unroll
      for (int k1 = 0; k1 < 16; k1++) {
        db[k1] = kDbPerLog2L * __log2f(__ldg(P + size_t(k1) * n2_count + k2));
        const uint32_t j = 16 * k2 + k1, i = j ^ half;
        bool cand = !(j < dc_ignore || (n - j) < dc_ignore);
        cand = cand && !(i < (half - use_window) || i > (half + use_window));
        bits |= ((cand && db[k1] > threshold) ? 1u : 0u) << k1;
      }
      if (spectra != nullptr) {
        float4* out = reinterpret_cast<float4*>(spectra + size_t(s) * n + 16 * size_t(k2));
#pragma unroll
        for (int x = 0; x < 4; x++) out[x] = make_float4(db[4 * x], db[4 * x + 1], db[4 * x + 2], db[4 * x + 3]);
      }
      // 16-bit group index in shifted order: g = (16 k2 ^ half) >> 4; two groups per mask word
      const uint32_t partner = __shfl_xor_sync(0xffffffffu, bits, 1);
      if ((k2 & 1u) == 0) smask[((k2 ^ (half >> 4)) >> 1)] = bits | (partner << 16);
    }
    __syncthreads();
    // phase 2: global mask, exclusive prefix over the words, total
    const uint32_t per = (words + 255) / 256;
    uint32_t local = 0;
    for (uint32_t x = 0; x < per; x++) {
      const uint32_t wi = tid * per + x;
      if (wi < words) local += __popc(smask[wi]);
    }
    s_scan[tid] = local;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {            // Hillis-Steele inclusive scan over 256 partials
      const uint32_t add = (tid >= o) ? s_scan[tid - o] : 0u;
      __syncthreads();
      s_scan[tid] += add;
      __syncthreads();
    }
    uint32_t run = s_scan[tid] - local;
    for (uint32_t x = 0; x < per; x++) {
      const uint32_t wi = tid * per + x;
      if (wi < words) { sprefix[wi] = run; run += __popc(smask[wi]); }
    }
    if (tid == 255 && counts != nullptr) counts[s] = s_scan[255];
    __syncthreads();
    if (masks != nullptr)
      for (uint32_t wi = tid; wi < words; wi += 256) masks[size_t(s) * words + wi] = smask[wi];
    // phase 3: hit records in ascending shifted bin (rare)
    if (hits != nullptr && s_scan[255] != 0) {
      for (uint32_t wi = tid; wi < words; wi += 256) {
        uint32_t mw = smask[wi];
        uint32_t rank = sprefix[wi];
        while (mw) {
          const uint32_t bit = __ffs(mw) - 1;
          mw &= mw - 1;
          const uint32_t i = wi * 32 + bit, j = i ^ half;
          if (rank < hit_cap) {
            scn_hit h;
            h.bin = i;
            h.power_db = kDbPerLog2L * __log2f(__ldg(P + size_t(j & 15) * n2_count + (j >> 4)));
            hits[size_t(s) * hit_cap + rank] = h;
          }
          rank++;
        }
      }
    }
    __syncthreads();
  }
}

cudaError_t launch_dc_sums(uint32_t kind, const uint8_t* raw, uint32_t n_buffers, uint32_t n, int2* dcs,
                           int num_sms, cudaStream_t stream) {
  uint32_t grid = uint32_t(num_sms) * 8;
  if (grid > n_buffers) grid = n_buffers;
  switch (kind) {
    case SCN_KIND_BYTE_COMPLEX: dc_sums_kernel<SCN_KIND_BYTE_COMPLEX><<<grid, 256, 0, stream>>>(raw, n_buffers, n, dcs); break;
    case SCN_KIND_SHORT: dc_sums_kernel<SCN_KIND_SHORT><<<grid, 256, 0, stream>>>(raw, n_buffers, n, dcs); break;
    case SCN_KIND_SHORT_COMPLEX: dc_sums_kernel<SCN_KIND_SHORT_COMPLEX><<<grid, 256, 0, stream>>>(raw, n_buffers, n, dcs); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_columns(uint32_t kind, const uint8_t* raw, uint32_t n_buffers, uint32_t n2_count,
                           const float* window, const float2* wn, const int2* dcs, float2* y, int num_sms,
                           cudaStream_t stream) {
  const uint32_t total = n_buffers * (n2_count / 256);
  uint32_t grid = uint32_t(num_sms) * 8;
  if (grid > total) grid = total;
  switch (kind) {
    case SCN_KIND_BYTE_COMPLEX: columns_kernel<SCN_KIND_BYTE_COMPLEX><<<grid, 256, 0, stream>>>(raw, n_buffers, n2_count, window, wn, dcs, y); break;
    case SCN_KIND_SHORT: columns_kernel<SCN_KIND_SHORT><<<grid, 256, 0, stream>>>(raw, n_buffers, n2_count, window, wn, dcs, y); break;
    case SCN_KIND_SHORT_COMPLEX: columns_kernel<SCN_KIND_SHORT_COMPLEX><<<grid, 256, 0, stream>>>(raw, n_buffers, n2_count, window, wn, dcs, y); break;
    case SCN_KIND_FLOAT_COMPLEX: columns_kernel<SCN_KIND_FLOAT_COMPLEX><<<grid, 256, 0, stream>>>(raw, n_buffers, n2_count, window, wn, dcs, y); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_finalize(const float* power, uint32_t n_spectra, uint32_t n2_count, float threshold,
                            uint32_t use_window, uint32_t dc_ignore, float* spectra, uint32_t* masks,
                            uint32_t* counts, scn_hit* hits, uint32_t hit_cap, int num_sms, cudaStream_t stream) {
  uint32_t grid = uint32_t(num_sms) * 4;
  if (grid > n_spectra) grid = n_spectra;
  const size_t smem = sizeof(uint32_t) * 2 * (size_t(n2_count) * 16 / 32);
  finalize_kernel<<<grid, 256, smem, stream>>>(power, n_spectra, n2_count, threshold, use_window, dc_ignore,
                                               spectra, masks, counts, hits, hit_cap);
  return cudaGetLastError();
}

}  // namespace scn
