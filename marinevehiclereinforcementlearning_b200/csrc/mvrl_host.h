// Host-side helpers shared by the translation units of libmvrl.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mvrl.h"

int mvrl_fail(int code, const char* fmt, ...);

#define MVRL_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return mvrl_fail(MVRL_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static inline unsigned mvrl_grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

static inline int mvrl_check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return mvrl_fail(MVRL_ECUDA, "%s launch failed: %s", what, cudaGetErrorString(e)); }
    return MVRL_OK;
}

// shared by every create(): validates the device ordinal, fails without a GPU (no CPU path)
int mvrl_require_device(int device);
bool mvrl_invert_n(const double* a, double* out, int n);
