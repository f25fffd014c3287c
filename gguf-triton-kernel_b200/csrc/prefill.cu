// prefill.cu — tensor-bound family (placeholder until the tcgen05 kernel lands in this file).
#include "../../include/ggq.h"
#include "common.cuh"

namespace ggq {
bool prefill_supports(int, const MmArgs&) { return false; }
int launch_prefill(int, const MmArgs&) { return GGQ_E_FAMILY; }
}  // namespace ggq
