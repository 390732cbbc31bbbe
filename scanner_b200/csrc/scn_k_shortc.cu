// int16 interleaved IQ (BladeRF style, messageQueue.h:205) instantiations.
#include "scn_dispatch.h"
namespace scn {
bool variant_short_complex(int log2n, bool dc, KernelVariant* out) {
  if (dc) { SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, true, "spectrum_sense<int16 IQ, dc>") }
  SCN_VARIANT_TABLE(SCN_KIND_SHORT_COMPLEX, false, "spectrum_sense<int16 IQ>")
}
}  // namespace scn
