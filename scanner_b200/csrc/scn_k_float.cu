// fp32 interleaved IQ (B210 / Airspy style, messageQueue.h:231) instantiations.
#include "scn_dispatch.h"
namespace scn {
bool variant_float_complex(int log2n, bool /*dc*/, bool avg, KernelVariant* out) {
  if (avg) { SCN_VARIANT_TABLE(SCN_KIND_FLOAT_COMPLEX, false, true, "spectrum_sense<fp32 IQ, avg>") }
  SCN_VARIANT_TABLE(SCN_KIND_FLOAT_COMPLEX, false, false, "spectrum_sense<fp32 IQ>")
}
}  // namespace scn
