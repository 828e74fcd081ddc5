// gc_host.h - host-side helpers shared by the .cu translation units of libgymcook.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gymcook.h"

struct GcLevelsDev;

// record `msg` (printf-style) as the thread's last error and return `code`
int gc_fail(int code, const char* fmt, ...);
// cudaGetLastError() after a launch -> GC_OK / GC_E_CUDA
int gc_check_launch(const char* what);
// GC_OK if a CUDA device is usable, else GC_E_CUDA ("no CPU fallback")
int gc_require_device();
// validate the host level array and convert it to the kernels' bitboard form
int gc_levels_to_dev(const gc_level* levels, int n_levels, int n_agents, GcLevelsDev* out, int* max_objs);
