// Shared helpers for the monosdf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/monosdf_b200.h"

#define MSDF_OK 0
#define MSDF_ERR_ARG 1
#define MSDF_ERR_CUDA 2
#define MSDF_ERR_UNSUPPORTED 3

// Thread-local last error text, read through msdf_last_error() (api.cu).
void msdf_set_error(const char* fmt, ...);

#define MSDF_CHECK_ARG(cond, ...)                         \
    do {                                                  \
        if (!(cond)) {                                    \
            msdf_set_error(__VA_ARGS__);                  \
            return MSDF_ERR_ARG;                          \
        }                                                 \
    } while (0)

#define MSDF_CHECK_LAUNCH(name)                                                          \
    do {                                                                                 \
        cudaError_t e_ = cudaGetLastError();                                             \
        if (e_ != cudaSuccess) {                                                         \
            msdf_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e_));  \
            return MSDF_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

#define MSDF_CUDA_CALL(expr)                                                              \
    do {                                                                                  \
        cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            msdf_set_error("%s failed: %s", #expr, cudaGetErrorString(e_));               \
            return MSDF_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

static inline int64_t msdf_div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t msdf_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// kernel launch counter (bench.py reports gpu_launches from it)
extern unsigned long long g_msdf_launches;
#define MSDF_COUNT_LAUNCH() (++g_msdf_launches)

// Optional per-launch device timing (CUDA events on the launching stream) of the dominant kernels, read by
// bench.py for the roofline line.  Disabled by default: msdf_prof_begin returns -1 and records nothing.
enum { MSDF_PROF_GEMM_F32 = 0, MSDF_PROF_GEMM_TC = 1, MSDF_PROF_HASH = 2, MSDF_PROF_SAMPLER = 3, MSDF_PROF_RENDER = 4, MSDF_PROF_CLASSES = 5 };
// sub-classes of the tcgen05 class (class id | sub << 8): msdf_profile_read(1, ..) sums all of them, msdf_profile_read(1 | k << 8, ..) one
enum { MSDF_PROF_TC_FUSED_SDF = MSDF_PROF_GEMM_TC | (1 << 8), MSDF_PROF_TC_FUSED_TRAIN = MSDF_PROF_GEMM_TC | (2 << 8),
       MSDF_PROF_TC_CHAIN = MSDF_PROF_GEMM_TC | (3 << 8), MSDF_PROF_TC_STREAM = MSDF_PROF_GEMM_TC | (4 << 8),
       MSDF_PROF_TC_GEMM = MSDF_PROF_GEMM_TC | (5 << 8), MSDF_PROF_TC_WGRAD = MSDF_PROF_GEMM_TC | (6 << 8) };
int msdf_prof_begin(int cls, double work, cudaStream_t st, double bytes = 0.0);   // work: FLOPs; bytes: algorithmic bytes
void msdf_prof_end(int slot, cudaStream_t st);
