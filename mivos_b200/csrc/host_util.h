// Host-side helpers shared by the C-ABI translation units: error codes, the driver entry point
// for cuTensorMapEncodeTiled (resolved at run time so the library links without libcuda), and the
// per-process device error flag the bounded mbarrier waits write to.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mivos_b200.h"

namespace mivos {

#define MIVOS_CUDA_OK(expr)                                                         \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return MIVOS_ERR_CUDA;                                                        \
    }                                                                               \
  } while (0)

#define MIVOS_REQUIRE(cond, ...)             \
  do {                                       \
    if (!(cond)) {                           \
      set_last_error(__VA_ARGS__);           \
      return MIVOS_ERR_INVALID;              \
    }                                        \
  } while (0)

void set_last_error(const char* fmt, ...);

// Encode a 2D row-major fp32 tensor [rows, cols] (row pitch `pitch_elems`) with a
// {box_cols x box_rows} box and 128-byte swizzle. Returns MIVOS_OK or an error code.
// `elem_bytes` 4 = fp32 (default), 2 = fp16; cols / pitch / box_cols are in elements.
int encode_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                   uint64_t pitch_elems, uint32_t box_cols, uint32_t box_rows, int elem_bytes = 4);

// Device int (one per process/device) that bounded waits write a non-zero code into.
int* device_error_flag();

// SM count of the current device (cached).
int num_sms();

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace mivos
