// int16 interleaved IQ (BladeRF style, messageQueue.h:205) instantiations.
#include "scn_dispatch.h"
#include "scn_p64.cuh"
#ifndef SCN_P64
#define SCN_P64 1      // 64-points-per-thread kernel with TMA-staged raw buffers for N = 8192, K = 1
#endif
namespace scn {
bool variant_short_complex(int log2n, bool dc, bool avg, KernelVariant* out) {
  if (SCN_P64 && log2n == 13 && !avg) {
    out->func = dc ? reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<SCN_KIND_SHORT_COMPLEX, true>)
                   : reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<SCN_KIND_SHORT_COMPLEX, false>);
    out->threads = kP64Threads;
    out->smem_bytes = p64_smem_bytes<SCN_KIND_SHORT_COMPLEX>();
    out->transforms_per_cta = 1;
    out->name = dc ? "spectrum_sense_p64<int16 IQ, dc, tma-staged><N=2^13>" : "spectrum_sense_p64<int16 IQ, tma-staged><N=2^13>";
    out->twiddle_layout = 2;
    return true;
  }
  if (dc && avg) { SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, true, true, "spectrum_sense<int16 IQ, dc, avg>") }
  if (dc) { SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, true, false, "spectrum_sense<int16 IQ, dc>") }
  if (avg) { SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, false, true, "spectrum_sense<int16 IQ, avg>") }
  SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, false, false, "spectrum_sense<int16 IQ>")
}
}  // namespace scn
