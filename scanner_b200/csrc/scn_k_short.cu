// int16 split I / Q arrays (SDRplay style, messageQueue.h:190) instantiations.
#include "scn_dispatch.h"
namespace scn {
bool variant_short(int log2n, bool dc, bool avg, KernelVariant* out) {
  if (dc && avg) { SCN_VARIANT_TABLE(SCN_KIND_SHORT, true, true, "spectrum_sense<int16 split, dc, avg>") }
  if (dc) { SCN_VARIANT_TABLE(SCN_KIND_SHORT, true, false, "spectrum_sense<int16 split, dc>") }
  if (avg) { SCN_VARIANT_TABLE(SCN_KIND_SHORT, false, true, "spectrum_sense<int16 split, avg>") }
  SCN_VARIANT_TABLE(SCN_KIND_SHORT, false, false, "spectrum_sense<int16 split>")
}
}  // namespace scn
