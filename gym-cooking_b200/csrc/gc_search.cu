// gc_search.cu - path B, part 2: exact level-0 subtask values (work in progress)
#include "gc_device.cuh"
#include "gc_host.h"

extern "C" {

int gc_subtask_q(const gc_level*, int, const uint8_t*, const uint32_t*, const uint8_t*, int, float*, float*,
                 uint8_t*, int64_t, int, void*) {
  return gc_fail(GC_E_ARG, "gc_subtask_q: not built yet");
}

}  // extern "C"
