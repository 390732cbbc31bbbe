// int8 interleaved IQ (HackRF / RTL style, messageQueue.h:218) instantiations.
#include "scn_dispatch.h"
namespace scn {
bool variant_byte_complex(int log2n, bool dc, bool avg, KernelVariant* out) {
  if (dc && avg) { SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, true, true, "spectrum_sense<int8 IQ, dc, avg>") }
  if (dc) { SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, true, false, "spectrum_sense<int8 IQ, dc>") }
  if (avg) { SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, false, true, "spectrum_sense<int8 IQ, avg>") }
  SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, false, false, "spectrum_sense<int8 IQ>")
}
}  // namespace scn
