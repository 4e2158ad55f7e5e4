// Shared helpers for the hsv kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "hsv.h"

#ifdef __CUDA_ARCH__
#if __CUDA_ARCH__ < 1000
#error "hsv kernels are written for sm_100a (Blackwell B200) only"
#endif
#endif

namespace hsv {

void set_error(const char *fmt, ...);

inline int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return HSV_ERR_CUDA;
  }
  return HSV_OK;
}

#define HSV_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      hsv::set_error(__VA_ARGS__);    \
      return HSV_ERR_ARG;             \
    }                                 \
  } while (0)

// 12-tap kaiser-sinc filter, cutoff 0.25, half-width 0.3 (alias_free_torch/filter.py:28-57);
// fp32 values as stored in both bundled checkpoints.  Symmetric: f[j] == f[11-j].
#define HSV_F0 0.0020289647f
#define HSV_F1 0.0093894657f
#define HSV_F2 (-0.0255434588f)
#define HSV_F3 (-0.0576573834f)
#define HSV_F4 0.1285725832f
#define HSV_F5 0.4432097971f

__host__ __device__ inline int64_t blk16_rows(int64_t L) {
  return 2 * (int64_t)HSV_BLK_PAD + ((L + HSV_BLK_ROUND - 1) / HSV_BLK_ROUND) * HSV_BLK_ROUND;
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (PDL): a kernel launched with the attribute may start while its
// predecessor in the stream is still running; it must execute pdl_wait() before touching any memory
// the predecessor reads or writes.  Work that only depends on static data (weights, parameters,
// barrier/TMEM set-up) goes before the wait and overlaps the predecessor's tail.
extern int g_pdl;  // 0 = plain stream order (bring-up switch), 1 = PDL on the act/conv kernels
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace hsv
