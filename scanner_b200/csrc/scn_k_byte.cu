// int8 interleaved IQ (HackRF / RTL style, messageQueue.h:218) instantiations.
#include "scn_dispatch.h"
namespace scn {
bool variant_byte_complex(int log2n, bool dc, KernelVariant* out) {
  if (dc) { SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, true, "spectrum_sense<int8 IQ, dc>") }
  SCN_VARIANT_TABLE(SCN_KIND_BYTE_COMPLEX, false, "spectrum_sense<int8 IQ>")
}
}  // namespace scn
