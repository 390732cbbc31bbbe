// int16 interleaved IQ (BladeRF style, messageQueue.h:205) instantiations.
#include "scn_dispatch.h"
namespace scn {
bool variant_short_complex(int log2n, bool dc, bool avg, KernelVariant* out) {
  if (dc && avg) { SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, true, true, "spectrum_sense<int16 IQ, dc, avg>") }
  if (dc) { SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, true, false, "spectrum_sense<int16 IQ, dc>") }
  if (avg) { SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, false, true, "spectrum_sense<int16 IQ, avg>") }
  SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, false, false, "spectrum_sense<int16 IQ>")
}
}  // namespace scn
