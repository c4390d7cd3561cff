// dkb_scan_inst.cu - the instantiations of kernel 2 for ONE probe stride (DKB_INST_D = 1, 2,
// 4, 8 or 16), compiled as its own translation unit so that the strides build in parallel.
// Shared-memory filter: 1..4 filter bits; L2 filter (strides 2..16): 1..2 bits; each with and
// without the profiling counters.  DKB_AB_BUILD (scripts/ab_build.sh): 2 filter bits only.
#include "dkb_scan.cuh"

#ifndef DKB_INST_D
#error "compile with -DDKB_INST_D=<stride>"
#endif
#define DKB_CAT2(a, b) a##b
#define DKB_CAT(a, b) DKB_CAT2(a, b)

namespace dkb {
typedef void (*scan_fn)(const ScanParams);

scan_fn DKB_CAT(pick_scan_d, DKB_INST_D)(int NH, bool gf, bool prof) {
  constexpr int D = DKB_INST_D;
#define PICK(h, g)                                                                      \
  if (NH == h && gf == g)                                                               \
    return prof ? (scan_fn)k_scan<D, h, g, true> : (scan_fn)k_scan<D, h, g, false>;
  PICK(2, false)
#if DKB_INST_D >= 2
  PICK(2, true)
#endif
#ifndef DKB_AB_BUILD
  PICK(1, false) PICK(3, false) PICK(4, false)
#if DKB_INST_D >= 2
  PICK(1, true)
#endif
#endif
#undef PICK
  return nullptr;
}
}  // namespace dkb
