// Shared helpers for the ssrs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/ssrs_b200.h"

namespace ssrs {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int sm_count();

#define SSRS_CUDA_TRY(expr)                                                             \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) return ::ssrs::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define SSRS_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            ::ssrs::set_error(__VA_ARGS__); \
            return SSRS_ERR_INVALID;     \
        }                                \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ssrs
