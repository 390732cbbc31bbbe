// fp32 interleaved IQ (B210 / Airspy style, messageQueue.h:231) instantiations.
#include "scn_dispatch.h"
#include "scn_p64.cuh"
#ifndef SCN_P64
#define SCN_P64 1      // 64-points-per-thread kernels for N = 4096 / 8192 (BASELINE configs[2], configs[3])
#endif
namespace scn {
bool variant_float_complex(int log2n, bool /*dc*/, bool avg, KernelVariant* out) {
  if (SCN_P64 && (log2n == 12 || log2n == 13)) {
    if (log2n == 12) {
      out->func = avg ? reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<12, SCN_KIND_FLOAT_COMPLEX, false, true>)
                      : reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<12, SCN_KIND_FLOAT_COMPLEX, false, false>);
      out->threads = P64Geometry<12>::T;
      out->smem_bytes = p64_smem_bytes<12, SCN_KIND_FLOAT_COMPLEX>();
      out->name = "spectrum_sense_p64<fp32 IQ><N=2^12>";
    } else {
      out->func = avg ? reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<13, SCN_KIND_FLOAT_COMPLEX, false, true>)
                      : reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<13, SCN_KIND_FLOAT_COMPLEX, false, false>);
      out->threads = P64Geometry<13>::T;
      out->smem_bytes = p64_smem_bytes<13, SCN_KIND_FLOAT_COMPLEX>();
      out->name = "spectrum_sense_p64<fp32 IQ><N=2^13>";
    }
    out->transforms_per_cta = 1;
    out->twiddle_layout = 2;
    return true;
  }
  if (avg) { SCN_VARIANT_TABLE(SCN_KIND_FLOAT_COMPLEX, false, true, "spectrum_sense<fp32 IQ, avg>") }
  SCN_VARIANT_TABLE(SCN_KIND_FLOAT_COMPLEX, false, false, "spectrum_sense<fp32 IQ>")
}

#define SCN_ROWS_CASE(L, AVG, NAME)                                                                            \
  case L: {                                                                                                    \
    out->func = reinterpret_cast<const void*>(&spectrum_sense_kernel<L, SCN_KIND_FLOAT_COMPLEX, false, AVG, true>); \
    out->threads = Geometry<L>::THREADS;                                                                       \
    out->smem_bytes = Geometry<L>::kSmemBytes;                                                                 \
    out->transforms_per_cta = Geometry<L>::F;                                                                  \
    out->name = NAME "<N2=2^" #L ">";                                                                          \
    out->twiddle_layout = 0;                                                                                   \
    return true;                                                                                               \
  }
bool variant_float_rows(int log2n, bool avg, KernelVariant* out) {
  if (avg) {
    switch (log2n) {
      SCN_ROWS_CASE(11, true, "four_step: columns + spectrum_sense_rows<avg> + finalize")
      SCN_ROWS_CASE(12, true, "four_step: columns + spectrum_sense_rows<avg> + finalize")
      default: return false;
    }
  }
  switch (log2n) {
    SCN_ROWS_CASE(11, false, "four_step: columns + spectrum_sense_rows + finalize")
    SCN_ROWS_CASE(12, false, "four_step: columns + spectrum_sense_rows + finalize")
    default: return false;
  }
}
}  // namespace scn
