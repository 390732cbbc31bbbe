// int8 interleaved IQ (HackRF / RTL style, messageQueue.h:218) instantiations.
#include "scn_dispatch.h"
#include "scn_p64.cuh"
#ifndef SCN_P64
#define SCN_P64 1      // 64-points-per-thread kernels with TMA-staged raw buffers for N = 4096 / 8192
#endif
#include "scn_wpt.cuh"
#ifndef SCN_WPT
#define SCN_WPT 1      // warp-per-transform kernel for N = 2048, K = 1 (the headline workload)
#endif
namespace scn {
bool variant_byte_complex(int log2n, bool dc, bool avg, KernelVariant* out) {
  if (SCN_P64 && (log2n == 12 || log2n == 13)) {
    if (log2n == 12) {
      out->func = dc ? (avg ? reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<12, SCN_KIND_BYTE_COMPLEX, true, true>)
                           : reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<12, SCN_KIND_BYTE_COMPLEX, true, false>))
                    : (avg ? reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<12, SCN_KIND_BYTE_COMPLEX, false, true>)
                           : reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<12, SCN_KIND_BYTE_COMPLEX, false, false>));
      out->threads = P64Geometry<12>::T;
      out->smem_bytes = p64_smem_bytes<12, SCN_KIND_BYTE_COMPLEX>();
      out->name = "spectrum_sense_p64<int8 IQ, tma-staged><N=2^12>";
    } else {
      out->func = dc ? (avg ? reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<13, SCN_KIND_BYTE_COMPLEX, true, true>)
                           : reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<13, SCN_KIND_BYTE_COMPLEX, true, false>))
                    : (avg ? reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<13, SCN_KIND_BYTE_COMPLEX, false, true>)
                           : reinterpret_cast<const void*>(&spectrum_sense_p64_kernel<13, SCN_KIND_BYTE_COMPLEX, false, false>));
      out->threads = P64Geometry<13>::T;
      out->smem_bytes = p64_smem_bytes<13, SCN_KIND_BYTE_COMPLEX>();
      out->name = "spectrum_sense_p64<int8 IQ, tma-staged><N=2^13>";
    }
    out->transforms_per_cta = 1;
    out->twiddle_layout = 2;
    return true;
  }
  if (SCN_WPT && log2n == 11 && !avg) {
    out->func = dc ? reinterpret_cast<const void*>(&spectrum_sense_wpt_kernel<true>)
                   : reinterpret_cast<const void*>(&spectrum_sense_wpt_kernel<false>);
    out->threads = 32 * kWptWarpsPerCta;
    out->smem_bytes = kWptSmemBytes;
    out->transforms_per_cta = kWptWarpsPerCta;
    out->name = dc ? "spectrum_sense_wpt<int8 IQ, dc><N=2^11>" : "spectrum_sense_wpt<int8 IQ><N=2^11>";
    out->twiddle_layout = 1;
    return true;
  }
  if (dc && avg) { SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, true, true, "spectrum_sense<int8 IQ, dc, avg>") }
  if (dc) { SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, true, false, "spectrum_sense<int8 IQ, dc>") }
  if (avg) { SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, false, true, "spectrum_sense<int8 IQ, avg>") }
  SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, false, false, "spectrum_sense<int8 IQ>")
}
}  // namespace scn
