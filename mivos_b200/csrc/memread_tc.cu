// tcgen05 candidate generation for the memory read (stage A2).  Placeholder until the kernel
// lands: AUTO resolves to the exact SIMT path while this reports unavailable.
#include "memread.h"
namespace mivos {
bool memread_tc_available() { return false; }
int memread_tc_run(const float*, const float*, int64_t, int, int64_t, const float*, int, int, float*, int,
                   int, int, int, int32_t*, float*, void*, cudaStream_t) {
  set_last_error("memory_read: tcgen05 path not built yet");
  return MIVOS_ERR_INVALID;
}
}  // namespace mivos
