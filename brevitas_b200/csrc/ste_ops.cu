// brevitas_b200 :: the 12 straight-through-estimator primitives of the reference's native plugin
// (src/brevitas/csrc/autograd_ste_ops.cpp:14-194, bound at :258-271; Python twin
// src/brevitas/ops/autograd_ste_ops.py).  Forward values only: every backward except
// abs_binary_sign_grad is the identity and is handled by the autograd wrapper without a kernel.
//
// All of them are pure streaming kernels (1 read + 1 write per element): 128-bit loads/stores,
// four independent vectors in flight per thread, grid = one block per 1024 vectors.
#include "common.cuh"
#include "host.cuh"

namespace bvb {

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;

__device__ __forceinline__ uint4 ldg_v4(const uint4* p) {       // coherent (output may alias input)
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

// y[i] = f(a[i])  over whole 16-byte vectors
template <typename T, typename F>
__global__ void __launch_bounds__(EW_THREADS) ew1_vec_kernel(const T* a, T* y, int64_t nvec, F f) {
    constexpr int V = DT<T>::VEC;
    const int64_t base = (int64_t)blockIdx.x * (EW_THREADS * EW_UNROLL) + threadIdx.x;
    const uint4* av = reinterpret_cast<const uint4*>(a);
    uint4* yv = reinterpret_cast<uint4*>(y);
    uint4 q[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        int64_t v = base + (int64_t)u * EW_THREADS;
        if (v < nvec) q[u] = ldg_v4(av + v);
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
        int64_t v = base + (int64_t)u * EW_THREADS;
        if (v < nvec) {
            float e[V];
            DT<T>::unpack(q[u], e);
#pragma unroll
            for (int i = 0; i < V; ++i) e[i] = f(e[i]);
            stg_stream(yv + v, DT<T>::pack(e));
        }
    }
}

// scalar fallback for [start, n): ragged tails and pointers that are not 16-byte aligned
template <typename T, typename F>
__global__ void ew1_scalar_kernel(const T* a, T* y, int64_t start, int64_t n, F f) {
    int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = DT<T>::from_f(f(DT<T>::to_f(a[i])));
}

template <typename T, typename F>
int launch_ew1(const void* x, void* y, int64_t n, F f, cudaStream_t st, const char* name) {
    if (n < 0) return fail(BVB_EINVAL, "%s: negative element count", name);
    if (n == 0) return BVB_OK;
    if (!x || !y) return fail(BVB_EINVAL, "%s: null pointer", name);
    constexpr int V = DT<T>::VEC;
    int64_t nvec = 0;
    if (aligned16(x) && aligned16(y)) nvec = n / V;
    if (nvec > 0) {
        int64_t blocks = (nvec + EW_THREADS * EW_UNROLL - 1) / (EW_THREADS * EW_UNROLL);
        ew1_vec_kernel<T, F><<<(unsigned)blocks, EW_THREADS, 0, st>>>((const T*)x, (T*)y, nvec, f);
    }
    int64_t done = nvec * V;
    if (done < n) {
        int64_t rem = n - done;
        int64_t blocks = (rem + 255) / 256;
        if (blocks > 4096) blocks = 4096;
        ew1_scalar_kernel<T, F><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, (T*)y, done, n, f);
    }
    return check_launch(name);
}

// ---- functors (operate on the fp32 image of one element; the result is rounded to T on store) --------
template <typename T> struct FRound       { __device__ float operator()(float v) const { return rintf(v); } };
template <typename T> struct FCeil        { __device__ float operator()(float v) const { return ceilf(v); } };
template <typename T> struct FFloor       { __device__ float operator()(float v) const { return floorf(v); } };
template <typename T> struct FBinarySign  { __device__ float operator()(float v) const { return binary_sign_f(v); } };
template <typename T> struct FTernarySign { __device__ float operator()(float v) const { return sign3(v); } };
template <typename T> struct FRoundToZero { __device__ float operator()(float v) const { return round_to_zero_T<T>(v); } };
template <typename T> struct FDpuRound    { __device__ float operator()(float v) const { return dpu_round_T<T>(v); } };
template <typename T> struct FAbs         { __device__ float operator()(float v) const { return fabsf(v); } };
// torch.clamp(x, min, max) with Scalars: min(max(x, lo), hi), NaN-propagating
template <typename T> struct FScalarClamp {
    float lo, hi;
    __device__ float operator()(float v) const {
        if (v != v) return v;
        float t = (v < lo) ? lo : v;
        return (t > hi) ? hi : t;
    }
};
template <typename T> struct FScalarClampMin {
    float lo;
    __device__ float operator()(float v) const { return clamp_min_nan(v, lo); }
};

// ---- abs_binary_sign_grad backward: gx = binary_sign(x) * gy ------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) absgrad_bwd_kernel(const T* x, const T* gy, T* gx, int64_t n, int vec_ok) {
    constexpr int V = DT<T>::VEC;
    const int64_t nvec = vec_ok ? n / V : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t v = tid; v < nvec; v += stride) {
        uint4 qx = ldg_v4(reinterpret_cast<const uint4*>(x) + v);
        uint4 qg = ldg_v4(reinterpret_cast<const uint4*>(gy) + v);
        float ex[V], eg[V];
        DT<T>::unpack(qx, ex);
        DT<T>::unpack(qg, eg);
#pragma unroll
        for (int i = 0; i < V; ++i) eg[i] = fmul(binary_sign_f(ex[i]), eg[i]);
        stg_stream(reinterpret_cast<uint4*>(gx) + v, DT<T>::pack(eg));
    }
    for (int64_t i = nvec * V + tid; i < n; i += stride)
        gx[i] = DT<T>::from_f(fmul(binary_sign_f(DT<T>::to_f(x[i])), DT<T>::to_f(gy[i])));
}

// ---- tensor clamp with broadcast min/max tensors ----------------------------------------------------------
// where-based (function/ops.py:98-99): NaN in x passes, NaN bounds never clamp.
// inplace_minmax (function/ops.py:109-110): torch.min(x, max) then torch.max(., min): NaN-propagating.
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) tensor_clamp_kernel(
        const T* x, const T* mn, const T* mx, T* y, int64_t n,
        int64_t mn_inner, int64_t mn_count, int64_t mx_inner, int64_t mx_count, int minmax) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = DT<T>::to_f(x[i]);
        float lo = DT<T>::to_f(mn[mn_count == 1 ? 0 : (i / mn_inner) % mn_count]);
        float hi = DT<T>::to_f(mx[mx_count == 1 ? 0 : (i / mx_inner) % mx_count]);
        float r;
        if (minmax) {
            float t = (v != v) ? v : ((hi != hi) ? hi : ((hi < v) ? hi : v));   // torch.min
            r = (t != t) ? t : ((lo != lo) ? lo : ((lo > t) ? lo : t));         // torch.max
        } else {
            r = where_clamp(v, lo, hi);
        }
        y[i] = DT<T>::from_f(r);
    }
}

// backward of the DIFFERENTIABLE tensor_clamp (function/ops.py:76-100: two torch.where under plain autograd):
//   out1 = where(x > max, max, x);  out = where(out1 < min, min, out1)
//   d x   = g where neither branch replaced the value;  d max = sum g over {x > max and not (max < min)};
//   d min = sum g over {out1 < min}
// 16-byte vectors when the bounds are scalars (the learned bit-width case: min / max are 0-dim), per-warp partial sums.
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) tensor_clamp_bwd_kernel(
        const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ mn, const T* __restrict__ mx,
        T* __restrict__ gx, float* gmin, float* gmax, int64_t n, int64_t mn_inner, int64_t mn_count, int64_t mx_inner,
        int64_t mx_count, int vec_ok) {
    constexpr int V = DT<T>::VEC;
    const bool scalar = mn_count == 1 && mx_count == 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float acc_lo = 0.f, acc_hi = 0.f;
    auto one = [&](float g, float v, int64_t i) -> float {
        const int64_t il = mn_count == 1 ? 0 : (i / mn_inner) % mn_count, ih = mx_count == 1 ? 0 : (i / mx_inner) % mx_count;
        const float lo = DT<T>::to_f(mn[il]), hi = DT<T>::to_f(mx[ih]);
        const bool over = v > hi;
        const float o1 = over ? hi : v;
        const bool under = o1 < lo;
        if (under) { if (scalar) acc_lo += g; else if (gmin) atomicAdd(gmin + il, g); }
        else if (over) { if (scalar) acc_hi += g; else if (gmax) atomicAdd(gmax + ih, g); }
        return (over || under) ? 0.f : g;
    };
    const int64_t nvec = (scalar && vec_ok) ? n / V : 0;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const uint4* gv = reinterpret_cast<const uint4*>(gy);
    uint4* ov = reinterpret_cast<uint4*>(gx);
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        float ex[V], eg[V];
        DT<T>::unpack(ldg_stream(xv + v), ex);
        DT<T>::unpack(ldg_stream(gv + v), eg);
#pragma unroll
        for (int i = 0; i < V; ++i) eg[i] = one(eg[i], ex[i], 0);
        stg_stream(ov + v, DT<T>::pack(eg));
    }
    for (int64_t i = nvec * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        gx[i] = DT<T>::from_f(one(DT<T>::to_f(gy[i]), DT<T>::to_f(x[i]), i));
    if (scalar) {
        acc_lo = warp_sum_f(acc_lo);
        acc_hi = warp_sum_f(acc_hi);
        if ((threadIdx.x & 31) == 0) {
            if (gmin && acc_lo != 0.f) atomicAdd(gmin, acc_lo);
            if (gmax && acc_hi != 0.f) atomicAdd(gmax, acc_hi);
        }
    }
}

// scalar-bounds fast path of the same op (the common case: 0-dim min_int / max_int tensors)
template <typename T> struct FWhereClampScalarPtr {
    const T* mn; const T* mx;
    __device__ float operator()(float v) const { return where_clamp(v, DT<T>::to_f(*mn), DT<T>::to_f(*mx)); }
};

static inline unsigned grid_for(int64_t n, int threads, int per_thread) {
    int64_t b = (n + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
    int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace bvb

using namespace bvb;

#define BVB_DEFINE_UNARY(cname, Functor)                                                        \
    extern "C" int cname(const void* x, void* y, int64_t n, int dtype, void* stream) {          \
        BVB_DISPATCH_DTYPE(dtype, return launch_ew1<T>(x, y, n, Functor<T>{}, (cudaStream_t)stream, #cname)); \
        return BVB_OK;                                                                          \
    }

BVB_DEFINE_UNARY(bvb_round_ste_impl, FRound)
BVB_DEFINE_UNARY(bvb_ceil_ste_impl, FCeil)
BVB_DEFINE_UNARY(bvb_floor_ste_impl, FFloor)
BVB_DEFINE_UNARY(bvb_binary_sign_ste_impl, FBinarySign)
BVB_DEFINE_UNARY(bvb_ternary_sign_ste_impl, FTernarySign)
BVB_DEFINE_UNARY(bvb_round_to_zero_ste_impl, FRoundToZero)
BVB_DEFINE_UNARY(bvb_dpu_round_ste_impl, FDpuRound)
BVB_DEFINE_UNARY(bvb_abs_binary_sign_grad_impl, FAbs)

extern "C" int bvb_scalar_clamp_ste_impl(const void* x, void* y, int64_t n, double min_val, double max_val,
                                         int dtype, void* stream) {
    // ATen casts the Scalar bounds to the tensor dtype before clamping
    float lo = round_to_dtype((float)min_val, dtype), hi = round_to_dtype((float)max_val, dtype);
    BVB_DISPATCH_DTYPE(dtype, return launch_ew1<T>(x, y, n, FScalarClamp<T>{lo, hi}, (cudaStream_t)stream,
                                                   "bvb_scalar_clamp_ste_impl"));
    return BVB_OK;
}

extern "C" int bvb_scalar_clamp_min_ste_impl(const void* x, void* y, int64_t n, double min_val, int dtype,
                                             void* stream) {
    float lo = round_to_dtype((float)min_val, dtype);
    BVB_DISPATCH_DTYPE(dtype, return launch_ew1<T>(x, y, n, FScalarClampMin<T>{lo}, (cudaStream_t)stream,
                                                   "bvb_scalar_clamp_min_ste_impl"));
    return BVB_OK;
}

extern "C" int bvb_abs_binary_sign_grad_bwd(const void* x, const void* gy, void* gx, int64_t n, int dtype,
                                            void* stream) {
    if (n < 0) return fail(BVB_EINVAL, "bvb_abs_binary_sign_grad_bwd: negative element count");
    if (n == 0) return BVB_OK;
    if (!x || !gy || !gx) return fail(BVB_EINVAL, "bvb_abs_binary_sign_grad_bwd: null pointer");
    int vec_ok = aligned16(x) && aligned16(gy) && aligned16(gx);
    BVB_DISPATCH_DTYPE(dtype, absgrad_bwd_kernel<T><<<grid_for(n, EW_THREADS, 16), EW_THREADS, 0, (cudaStream_t)stream>>>(
                                  (const T*)x, (const T*)gy, (T*)gx, n, vec_ok));
    return check_launch("bvb_abs_binary_sign_grad_bwd");
}

extern "C" int bvb_tensor_clamp_ste_impl(const void* x, const void* min_val, const void* max_val, void* y, int64_t n,
                                         int64_t min_inner, int64_t min_count, int64_t max_inner, int64_t max_count,
                                         int inplace_minmax, int dtype, void* stream) {
    if (n < 0) return fail(BVB_EINVAL, "bvb_tensor_clamp_ste_impl: negative element count");
    if (n == 0) return BVB_OK;
    if (!x || !y || !min_val || !max_val) return fail(BVB_EINVAL, "bvb_tensor_clamp_ste_impl: null pointer");
    if (min_inner < 1 || min_count < 1 || max_inner < 1 || max_count < 1)
        return fail(BVB_EINVAL, "bvb_tensor_clamp_ste_impl: broadcast pattern must have inner >= 1 and count >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (min_count == 1 && max_count == 1 && !inplace_minmax) {
        BVB_DISPATCH_DTYPE(dtype, return launch_ew1<T>(x, y, n, FWhereClampScalarPtr<T>{(const T*)min_val, (const T*)max_val},
                                                       st, "bvb_tensor_clamp_ste_impl"));
    }
    BVB_DISPATCH_DTYPE(dtype, tensor_clamp_kernel<T><<<grid_for(n, EW_THREADS, 4), EW_THREADS, 0, st>>>(
                                  (const T*)x, (const T*)min_val, (const T*)max_val, (T*)y, n,
                                  min_inner, min_count, max_inner, max_count, inplace_minmax));
    return check_launch("bvb_tensor_clamp_ste_impl");
}

extern "C" int bvb_tensor_clamp_bwd(const void* gy, const void* x, const void* min_val, const void* max_val, void* gx,
                                    float* gmin_out, float* gmax_out, int64_t n, int64_t min_inner, int64_t min_count,
                                    int64_t max_inner, int64_t max_count, int dtype, void* stream) {
    if (n < 0) return fail(BVB_EINVAL, "bvb_tensor_clamp_bwd: negative element count");
    if (min_inner < 1 || min_count < 1 || max_inner < 1 || max_count < 1)
        return fail(BVB_EINVAL, "bvb_tensor_clamp_bwd: broadcast pattern must have inner >= 1 and count >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;
    if (gmin_out) e = cudaMemsetAsync(gmin_out, 0, sizeof(float) * (size_t)min_count, st);
    if (e == cudaSuccess && gmax_out) e = cudaMemsetAsync(gmax_out, 0, sizeof(float) * (size_t)max_count, st);
    if (e != cudaSuccess) return fail(BVB_ECUDA, "bvb_tensor_clamp_bwd: memset: %s", cudaGetErrorString(e));
    if (n == 0) return BVB_OK;
    if (!gy || !x || !min_val || !max_val || !gx) return fail(BVB_EINVAL, "bvb_tensor_clamp_bwd: null pointer");
    const int vec_ok = aligned16(gy) && aligned16(x) && aligned16(gx);
    BVB_DISPATCH_DTYPE(dtype, tensor_clamp_bwd_kernel<T><<<grid_for(n, EW_THREADS, 8), EW_THREADS, 0, st>>>(
                                  (const T*)gy, (const T*)x, (const T*)min_val, (const T*)max_val, (T*)gx, gmin_out, gmax_out,
                                  n, min_inner, min_count, max_inner, max_count, vec_ok));
    return check_launch("bvb_tensor_clamp_bwd");
}
