// scn_timedomain.cuh -- time-domain threshold mode (the reference CLI's default mode,
// scan.cpp:87): ProcessSamples::DoTimeDomainThresholding, process.cpp:203-237.
//
// Per buffer: convert (utility.cpp:9-84, incl. the unsigned-division DC quirk), then
//   mag = sqrt(re*re + im*im); magnitude = 10*log2(mag)/log2(10)
//   maxMagnitude = max over samples, seeded with numeric_limits<float>::min()  (process.cpp:207)
//   minMagnitude = min over samples, seeded with numeric_limits<float>::max()  (process.cpp:208)
//   trigger      = maxMagnitude >= threshold                                    (process.cpp:226)
// sqrt and log2 are monotone, so max/min are taken over the fp32 power p and converted to
// dB once per buffer.  One CTA streams one buffer with 128-bit loads; HBM-bound.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#include "../../include/scanner_b200.h"

namespace scn {

struct TimeDomainParams {
  const uint8_t* __restrict__ raw;
  uint32_t n_buffers;
  uint32_t n;                       // samples per buffer, multiple of 8
  float threshold;
  uint32_t* __restrict__ trigger;   // nullable [n_buffers]
  float* __restrict__ max_min;      // nullable [n_buffers][2]
  float scale;                      // onebymax (1 for the float kind)
  uint32_t correct_dc;
};

constexpr int kTdThreads = 256;

__device__ __forceinline__ float td_db(float p) {
  // float mag = sqrt(p); 10 * log2(mag) / log2(10.0)   (utility.cpp:91-97 arithmetic, float overloads)
  const float mag = __fsqrt_rn(p);
  return static_cast<float>(static_cast<double>(10.0f * log2f(mag)) / 3.3219280948873622);
}

__device__ __forceinline__ void td_acc(float re, float im, float& pmax, float& pmin) {
  const float pw = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
  pmax = fmaxf(pmax, pw);
  pmin = fminf(pmin, pw);
}

template <int KIND>
__global__ void __launch_bounds__(kTdThreads) time_domain_kernel(const TimeDomainParams p) {
  __shared__ int s_red[2][kTdThreads / 32];
  __shared__ float s_mm[2][kTdThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kBytes = (KIND == SCN_KIND_BYTE_COMPLEX) ? 2 : (KIND == SCN_KIND_FLOAT_COMPLEX ? 8 : 4);
  const size_t buf_bytes = size_t(p.n) * kBytes;

  for (uint32_t b = blockIdx.x; b < p.n_buffers; b += gridDim.x) {
    const uint8_t* buf = p.raw + size_t(b) * buf_bytes;
    int dci = 0, dcq = 0;
    if constexpr (KIND != SCN_KIND_FLOAT_COMPLEX) {
      if (p.correct_dc) {
        int si = 0, sq = 0;
        if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
          const int4* v4 = reinterpret_cast<const int4*>(buf);
          for (uint32_t i = tid; i < p.n / 8; i += kTdThreads) {
            const int4 v = __ldg(v4 + i);
            const int ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              si = __dp4a(ws[k], 0x00010001, si);
              sq = __dp4a(ws[k], 0x01000100, sq);
            }
          }
        } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
          const int4* v4 = reinterpret_cast<const int4*>(buf);
          for (uint32_t i = tid; i < p.n / 4; i += kTdThreads) {
            const int4 v = __ldg(v4 + i);
            const int ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              si += static_cast<short>(ws[k] & 0xffff);
              sq += static_cast<short>(static_cast<unsigned>(ws[k]) >> 16);
            }
          }
        } else {
          const int4* re4 = reinterpret_cast<const int4*>(buf);
          const int4* im4 = reinterpret_cast<const int4*>(buf + size_t(p.n) * 2);
          for (uint32_t i = tid; i < p.n / 8; i += kTdThreads) {
            const int4 a = __ldg(re4 + i), c = __ldg(im4 + i);
            const int ra[4] = {a.x, a.y, a.z, a.w}, ia[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
              si += static_cast<short>(ra[k] & 0xffff) + static_cast<short>(static_cast<unsigned>(ra[k]) >> 16);
              sq += static_cast<short>(ia[k] & 0xffff) + static_cast<short>(static_cast<unsigned>(ia[k]) >> 16);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          si += __shfl_xor_sync(0xffffffffu, si, o);
          sq += __shfl_xor_sync(0xffffffffu, sq, o);
        }
        if (lane == 0) { s_red[0][warp] = si; s_red[1][warp] = sq; }
        __syncthreads();
        si = 0; sq = 0;
#pragma unroll
        for (int i = 0; i < kTdThreads / 32; i++) { si += s_red[0][i]; sq += s_red[1][i]; }
        // int32 /= uint32 (utility.cpp:25-26,49-50,77-78): unsigned division, modular conversion back
        dci = static_cast<int>(static_cast<unsigned>(si) / p.n);
        dcq = static_cast<int>(static_cast<unsigned>(sq) / p.n);
      }
    }

    float pmax = 0.0f, pmin = FLT_MAX;
    bool any = false;
    if constexpr (KIND == SCN_KIND_BYTE_COMPLEX) {
      const int4* v4 = reinterpret_cast<const int4*>(buf);
      for (uint32_t i = tid; i < p.n / 8; i += kTdThreads) {
        const int4 v = __ldg(v4 + i);
        const int ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int xi = static_cast<signed char>((ws[k] >> (16 * h)) & 0xff);
            const int xq = static_cast<signed char>((ws[k] >> (16 * h + 8)) & 0xff);
            td_acc(__fmul_rn(static_cast<float>(xi - dci), p.scale),
                   __fmul_rn(static_cast<float>(xq - dcq), p.scale), pmax, pmin);
          }
        }
        any = true;
      }
    } else if constexpr (KIND == SCN_KIND_SHORT_COMPLEX) {
      const int4* v4 = reinterpret_cast<const int4*>(buf);
      for (uint32_t i = tid; i < p.n / 4; i += kTdThreads) {
        const int4 v = __ldg(v4 + i);
        const int ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int xi = static_cast<short>(ws[k] & 0xffff);
          const int xq = static_cast<short>(static_cast<unsigned>(ws[k]) >> 16);
          td_acc(__fmul_rn(static_cast<float>(xi - dci), p.scale),
                 __fmul_rn(static_cast<float>(xq - dcq), p.scale), pmax, pmin);
        }
        any = true;
      }
    } else if constexpr (KIND == SCN_KIND_SHORT) {
      const int4* re4 = reinterpret_cast<const int4*>(buf);
      const int4* im4 = reinterpret_cast<const int4*>(buf + size_t(p.n) * 2);
      for (uint32_t i = tid; i < p.n / 8; i += kTdThreads) {
        const int4 a = __ldg(re4 + i), c = __ldg(im4 + i);
        const int ra[4] = {a.x, a.y, a.z, a.w}, ia[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int xi = static_cast<short>((static_cast<unsigned>(ra[k]) >> (16 * h)) & 0xffff);
            const int xq = static_cast<short>((static_cast<unsigned>(ia[k]) >> (16 * h)) & 0xffff);
            td_acc(__fmul_rn(static_cast<float>(xi - dci), p.scale),
                   __fmul_rn(static_cast<float>(xq - dcq), p.scale), pmax, pmin);
          }
        }
        any = true;
      }
    } else {
      const float4* v4 = reinterpret_cast<const float4*>(buf);
      for (uint32_t i = tid; i < p.n / 2; i += kTdThreads) {
        const float4 v = __ldg(v4 + i);
        td_acc(v.x, v.y, pmax, pmin);
        td_acc(v.z, v.w, pmax, pmin);
        any = true;
      }
    }
    if (!any) { pmax = 0.0f; pmin = FLT_MAX; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
      pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
    }
    if (lane == 0) { s_mm[0][warp] = pmax; s_mm[1][warp] = pmin; }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < kTdThreads / 32; i++) { pmax = fmaxf(pmax, s_mm[0][i]); pmin = fminf(pmin, s_mm[1][i]); }
      // seeds of process.cpp:207-208
      const float max_db = fmaxf(FLT_MIN, td_db(pmax));
      const float min_db = fminf(FLT_MAX, td_db(pmin));
      if (p.max_min) { p.max_min[2 * size_t(b)] = max_db; p.max_min[2 * size_t(b) + 1] = min_db; }
      if (p.trigger) p.trigger[b] = (max_db >= p.threshold) ? 1u : 0u;
    }
    __syncthreads();   // s_red / s_mm reused by the next buffer
  }
}

inline cudaError_t launch_time_domain_kernel(uint32_t kind, uint32_t grid, cudaStream_t stream,
                                             const TimeDomainParams& p) {
  switch (kind) {
    case SCN_KIND_BYTE_COMPLEX: time_domain_kernel<SCN_KIND_BYTE_COMPLEX><<<grid, kTdThreads, 0, stream>>>(p); break;
    case SCN_KIND_SHORT: time_domain_kernel<SCN_KIND_SHORT><<<grid, kTdThreads, 0, stream>>>(p); break;
    case SCN_KIND_SHORT_COMPLEX: time_domain_kernel<SCN_KIND_SHORT_COMPLEX><<<grid, kTdThreads, 0, stream>>>(p); break;
    case SCN_KIND_FLOAT_COMPLEX: time_domain_kernel<SCN_KIND_FLOAT_COMPLEX><<<grid, kTdThreads, 0, stream>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace scn
