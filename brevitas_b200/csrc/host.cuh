// brevitas_b200 :: host-side helpers shared by the C-ABI translation units (status, errors, dispatch)
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>

#include "../../include/brevitas_b200.h"

namespace bvb {

// thread-local last-error text (bvb_last_error)
char* err_buf();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

struct Tuning {
    int rows_threads, rows_stages, rows_ctas_per_sm, stream_threads, stream_ctas_per_sm;
};
const Tuning& tuning();      // all zero (= heuristics) in the product build; see lib.cu
unsigned stat_grid(int64_t blocks_wanted, int default_per_sm);   // grid of a read-only statistic kernel (256-thread CTAs), int_quant.cu

struct QParams;
QParams make_qparams(float zero_point, float qmin, float qmax, int dtype);   // int_quant.cu

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// round a host float to the tensor dtype exactly like a 0-dim operand cast to the common dtype
inline float round_to_dtype(float v, int dtype) {
    if (dtype == BVB_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    if (dtype == BVB_F16) return __half2float(__float2half_rn(v));
    return v;
}
inline int dtype_size(int dtype) { return dtype == BVB_F32 ? 4 : 2; }

#define BVB_DISPATCH_DTYPE(dtype, ...)                                              \
    switch (dtype) {                                                                \
        case BVB_F32:  { using T = float;          __VA_ARGS__; break; }            \
        case BVB_BF16: { using T = __nv_bfloat16;  __VA_ARGS__; break; }            \
        case BVB_F16:  { using T = __half;         __VA_ARGS__; break; }            \
        default: return ::bvb::fail(BVB_EINVAL, "unknown dtype tag %d", (int)(dtype)); \
    }

#define BVB_DISPATCH_ROUND(rm, ...)                                                 \
    switch (rm) {                                                                   \
        case BVB_ROUND:         { constexpr int RM = 0; __VA_ARGS__; break; }       \
        case BVB_FLOOR:         { constexpr int RM = 1; __VA_ARGS__; break; }       \
        case BVB_CEIL:          { constexpr int RM = 2; __VA_ARGS__; break; }       \
        case BVB_ROUND_TO_ZERO: { constexpr int RM = 3; __VA_ARGS__; break; }       \
        case BVB_DPU_ROUND:     { constexpr int RM = 4; __VA_ARGS__; break; }       \
        default: return ::bvb::fail(BVB_EINVAL, "unknown round mode %d", (int)(rm)); \
    }

// forward kernels: specialise round-half-even with a zero zero-point (the default quantizers)
#define BVB_DISPATCH_MODE_FWD(rm, zp_is_zero, ...)                                  \
    if ((rm) == BVB_ROUND && (zp_is_zero)) { constexpr int RM = 0 | 8; __VA_ARGS__; } \
    else BVB_DISPATCH_ROUND(rm, __VA_ARGS__)

// backward kernels: additionally fix the clamp-gradient mode at compile time
#define BVB_DISPATCH_MODE_BWD(rm, zp_is_zero, masked, ...)                          \
    if ((rm) == BVB_ROUND && (zp_is_zero) && !(masked)) { constexpr int RM = 0 | 8 | 16; __VA_ARGS__; }      \
    else if ((rm) == BVB_ROUND && (zp_is_zero) && (masked)) { constexpr int RM = 0 | 8 | 16 | 32; __VA_ARGS__; } \
    else BVB_DISPATCH_ROUND(rm, __VA_ARGS__)

}  // namespace bvb
