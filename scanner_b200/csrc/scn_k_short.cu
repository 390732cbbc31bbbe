// int16 split I / Q arrays (SDRplay style, messageQueue.h:190) instantiations.
#include "scn_dispatch.h"
namespace scn {
bool variant_short(int log2n, bool dc, KernelVariant* out) {
  if (dc) { SCN_VARIANT_TABLE(SCN_KIND_SHORT, true, "spectrum_sense<int16 split, dc>") }
  SCN_VARIANT_TABLE(SCN_KIND_SHORT, false, "spectrum_sense<int16 split>")
}
}  // namespace scn
