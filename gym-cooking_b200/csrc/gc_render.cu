// gc_render.cu - path A': optional image_obs renderer (work in progress)
#include "gc_device.cuh"
#include "gc_host.h"

extern "C" {

int gc_render(const gc_level*, int, const uint8_t*, const uint32_t*, const uint8_t*, uint8_t*, int64_t, int, void*) {
  return gc_fail(GC_E_ARG, "gc_render: not built yet");
}

}  // extern "C"
