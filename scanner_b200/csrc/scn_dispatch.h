// scn_dispatch.h -- host-side dispatch table over the fused-kernel instantiations.
#pragma once
#include <cuda_runtime.h>
#include "scn_kernel.cuh"

namespace scn {

constexpr int kMinLog2N = 8;    // 256
constexpr int kMaxLog2N = 14;   // 16384

struct KernelVariant {
  const void* func;       // __global__ entry
  int threads;
  size_t smem_bytes;
  int transforms_per_cta;
  const char* name;
  int twiddle_layout;     // 0: pass tables of scn_fft.cuh; 1: warp-per-transform table [63][32] of scn_wpt.cuh; 2: scn_p64.cuh tables; 3: scn_cluster.cu
  int cluster = 0;        // > 0: thread-block cluster of this many CTAs per transform (scn_cluster.cu), one CTA per SM
};

// One translation unit per sample kind (compiled in parallel); each fills its rows.
bool variant_byte_complex(int log2n, bool dc, bool avg, KernelVariant* out);
bool variant_short(int log2n, bool dc, bool avg, KernelVariant* out);
bool variant_short_complex(int log2n, bool dc, bool avg, KernelVariant* out);
bool variant_float_complex(int log2n, bool dc, bool avg, KernelVariant* out);
// row mode of the four-step path: fp32 complex rows of 2^11 / 2^12 points, power out
bool variant_float_rows(int log2n, bool avg, KernelVariant* out);
// N = 2^14 .. 2^16: one transform per cluster of 1 / 2 / 4 CTAs in distributed shared memory (scn_cluster.cu)
bool variant_cluster(int kind, int log2n, bool dc, bool avg, KernelVariant* out);

// avg: K > 1 (keeps the 16 accumulators live across the K loop; K == 1 kernels do not pay for them)
inline bool find_variant(int kind, int log2n, bool dc, bool avg, KernelVariant* out) {
  switch (kind) {
    case SCN_KIND_BYTE_COMPLEX: return variant_byte_complex(log2n, dc, avg, out);
    case SCN_KIND_SHORT: return variant_short(log2n, dc, avg, out);
    case SCN_KIND_SHORT_COMPLEX: return variant_short_complex(log2n, dc, avg, out);
    case SCN_KIND_FLOAT_COMPLEX: return variant_float_complex(log2n, false, avg, out);
    default: return false;
  }
}

#define SCN_VARIANT_CASE(L, KIND, DC, AVG, NAME)                                                   \
  case L: {                                                                                   \
    out->func = reinterpret_cast<const void*>(&spectrum_sense_kernel<L, KIND, DC, AVG>);          \
    out->threads = Geometry<L>::THREADS;                                                      \
    out->smem_bytes = Geometry<L>::kSmemBytes;                                                \
    out->transforms_per_cta = Geometry<L>::F;                                                 \
    out->name = NAME "<N=2^" #L ">";                                                          \
    out->twiddle_layout = 0;                                                                  \
    return true;                                                                              \
  }

#ifdef SCN_ONLY_LOG2N   /* experiment builds: a single transform size */
#define SCN_VARIANT_TABLE(KIND, DC, AVG, NAME)                                                \
  switch (log2n) {                                                                            \
    SCN_VARIANT_CASE(SCN_ONLY_LOG2N, KIND, DC, AVG, NAME)                                     \
    default: return false;                                                                    \
  }
#else
#define SCN_VARIANT_TABLE(KIND, DC, AVG, NAME)                                                     \
  switch (log2n) {                                                                            \
    SCN_VARIANT_CASE(8, KIND, DC, AVG, NAME)                                                       \
    SCN_VARIANT_CASE(9, KIND, DC, AVG, NAME)                                                       \
    SCN_VARIANT_CASE(10, KIND, DC, AVG, NAME)                                                      \
    SCN_VARIANT_CASE(11, KIND, DC, AVG, NAME)                                                      \
    SCN_VARIANT_CASE(12, KIND, DC, AVG, NAME)                                                      \
    SCN_VARIANT_CASE(13, KIND, DC, AVG, NAME)                                                      \
    SCN_VARIANT_CASE(14, KIND, DC, AVG, NAME)                                                      \
    default: return false;                                                                    \
  }
#endif

}  // namespace scn
