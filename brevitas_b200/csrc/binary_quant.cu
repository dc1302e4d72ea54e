// brevitas_b200 :: binary fake-quantization, forward and STE backward.
//
// Replaces the ATen op chains behind
//   BinaryQuant.forward          src/brevitas/core/quant/binary.py:60-64    y = binary_sign_ste(x) * scale
//   ClampedBinaryQuant.forward   src/brevitas/core/quant/binary.py:120-125  y = binary_sign_ste(tensor_clamp(x, -scale, scale)) * scale
// and the autograd graph behind them: the clamp of the second form only shapes the
// gradient (TensorClamp, function/ops.py:98-99 => masked gradient + gradient to the scale).
//
// Both directions are pure streaming kernels (fwd 1R+1W, bwd 2R+1W) with 128-bit loads and stores.
#include "common.cuh"
#include "host.cuh"

namespace bvb {

constexpr int BQ_THREADS = 256;
constexpr int BQ_UNROLL = 4;

template <typename T>
__device__ __forceinline__ float load_scale0(const T* scale, int scale_f32) {   // see int_quant.cu
    return scale_f32 ? reinterpret_cast<const float*>(scale)[0] : DT<T>::to_f(scale[0]);
}

// forward value of one element.  The clamp never changes the sign of the result except through NaN
// bounds, so the clamped variant is evaluated literally: c = where(x > s, s, x); c = where(c < -s, -s, c).
template <typename T, bool CLAMPED>
__device__ __forceinline__ float binary_fwd_elem(float x, float s) {
    float c = x;
    if (CLAMPED) c = where_clamp(x, -s, s);
    return fmul(binary_sign_f(c), s);
}

template <typename T, bool CLAMPED>
__global__ void __launch_bounds__(BQ_THREADS) binary_quant_fwd_kernel(
        const T* __restrict__ x, const T* __restrict__ scale, T* __restrict__ y,
        int64_t nvec, int64_t inner_v, int64_t count, int smode, int scale_f32) {
    constexpr int V = DT<T>::VEC;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    uint4* yv = reinterpret_cast<uint4*>(y);
    const int64_t chunk = BQ_THREADS * BQ_UNROLL;
    const int64_t nchunks = (nvec + chunk - 1) / chunk;
    float s0 = 0.f;
    if (smode == 0) s0 = load_scale0<T>(scale, scale_f32);
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t base = c * chunk + threadIdx.x;
        uint4 q[BQ_UNROLL];
#pragma unroll
        for (int u = 0; u < BQ_UNROLL; ++u) {
            int64_t v = base + (int64_t)u * BQ_THREADS;
            if (v < nvec) q[u] = ldg_stream(xv + v);
        }
#pragma unroll
        for (int u = 0; u < BQ_UNROLL; ++u) {
            int64_t v = base + (int64_t)u * BQ_THREADS;
            if (v < nvec) {
                float s = s0;
                if (smode != 0) s = DT<T>::to_f(scale[(v / inner_v) % count]);
                float e[V];
                DT<T>::unpack(q[u], e);
#pragma unroll
                for (int i = 0; i < V; ++i) e[i] = binary_fwd_elem<T, CLAMPED>(e[i], s);
                stg_stream(yv + v, DT<T>::pack(e));
            }
        }
    }
}

template <typename T, bool CLAMPED>
__global__ void binary_quant_fwd_scalar_kernel(const T* x, const T* scale, T* y, int64_t start, int64_t n,
                                               int64_t inner, int64_t count, int scale_f32) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float s = count == 1 ? load_scale0<T>(scale, scale_f32) : DT<T>::to_f(scale[(i / inner) % count]);
        y[i] = DT<T>::from_f(binary_fwd_elem<T, CLAMPED>(DT<T>::to_f(x[i]), s));
    }
}

// backward of one element; returns gx, accumulates d(loss)/d(scale)
//   y = sign_b(c) * s            =>  d c = 0 through the STE?  No: binary_sign_ste is a straight-through
//   estimator, so d c = d y * s (MulBackward wrt sign_b(c)) passes unchanged through the sign.
//   d s += g * sign_b(c)                                   (MulBackward wrt scale)
//   clamped: c = clamp(x, -s, s): d x = d c * [not clamped]; d s += d c * [x > s] - d c * [c1 < -s]
template <typename T, bool CLAMPED>
__device__ __forceinline__ float binary_bwd_elem(float g, float x, float s, bool want_gs, float& gs_acc) {
    float d = DT<T>::rnd(fmul(g, s));
    float c = x;
    bool hi = false, lo = false;
    if (CLAMPED) {
        hi = x > s;
        float c1 = hi ? s : x;
        lo = c1 < -s;
        c = lo ? -s : c1;
    }
    if (want_gs) {
        gs_acc += g * binary_sign_f(c);
        if (CLAMPED) {
            if (hi) gs_acc += d;
            if (lo) gs_acc -= d;
        }
    }
    if (CLAMPED && (hi || lo)) return 0.f;
    return d;
}

// One 16-byte vector of the backward in packed-pair arithmetic (bf16 / fp16, ONE positive finite scale stored in T --
// the learned / constant scalar scales of the binary quantizers): grad*scale is one HMUL2 (the fp32 product of two T
// values is exact, hence a single rounding like rnd_T(g*s)), the clamp masks are two packed compares (for s > 0 the
// low test "clamp_hi(x) < -s" is "x < -s"), clamped lanes become +0 by masking, and d(scale) adds
// g*sign_b(x) (sign flipped where x < 0, times 0 where x is NaN: sign_b(NaN) = 0) + d[x > s] - d[x < -s] in fp32.
// Element-wise results are bit-identical to binary_bwd_elem; the d(scale) sum only changes its summation order.
template <typename T, bool CLAMPED>
__device__ __forceinline__ uint4 binary_bwd_vec_packed(const uint4& qg, const uint4& qx, uint32_t s2, uint32_t ns2,
                                                       bool want_gs, float& acc) {
    const uint32_t g[4] = {qg.x, qg.y, qg.z, qg.w};
    const uint32_t x[4] = {qx.x, qx.y, qx.z, qx.w};
    uint32_t o[4];
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t d = DT<T>::p_mul(g[j], s2);
        uint32_t hi = 0u, lo = 0u;
        if (CLAMPED) {
            hi = DT<T>::p_gt_mask(x[j], s2);
            lo = DT<T>::p_lt_mask(x[j], ns2);
        }
        o[j] = d & ~(hi | lo);
        if (want_gs) {
            const uint32_t neg = DT<T>::p_lt_mask(x[j], 0u);
            const uint32_t nan = DT<T>::p_gtu_mask(x[j], x[j]);
            float t0, t1;
            DT<T>::p_unpack(g[j] ^ (neg & 0x80008000u), t0, t1);
            if (nan & 0xffffu) t0 = fmul(t0, 0.f);                  // g * sign_b(NaN) = g * 0 (NaN for an infinite g)
            if (nan >> 16) t1 = fmul(t1, 0.f);
            a0 += t0; a1 += t1;
            if (CLAMPED) {
                DT<T>::p_unpack((d & (hi | lo)) ^ (lo & 0x80008000u), t0, t1);
                a0 += t0; a1 += t1;
            }
        }
    }
    if (want_gs) acc += a0 + a1;
    return make_uint4(o[0], o[1], o[2], o[3]);
}

template <typename T, bool CLAMPED>
__global__ void __launch_bounds__(BQ_THREADS) binary_quant_bwd_kernel(
        const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ scale, T* __restrict__ gx,
        float* gscale_out, int64_t nvec, int64_t inner_v, int64_t count, int smode, int scale_f32) {
    constexpr int V = DT<T>::VEC;
    __shared__ float red[32];
    const uint4* gv = reinterpret_cast<const uint4*>(gy);
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    uint4* ov = reinterpret_cast<uint4*>(gx);
    const bool want_gs = gscale_out != nullptr;
    const int64_t chunk = BQ_THREADS * BQ_UNROLL;
    const int64_t nchunks = (nvec + chunk - 1) / chunk;
    float s0 = 0.f;
    if (smode == 0) s0 = load_scale0<T>(scale, scale_f32);
    float acc = 0.f;
    int64_t acc_idx = -1;
    bool packed = false;
    uint32_t s2 = 0u, ns2 = 0u;
    if constexpr (DT<T>::LOWP) {
        packed = smode == 0 && !scale_f32 && s0 > 0.f && s0 < __int_as_float(0x7f800000);
        s2 = DT<T>::pack2(s0, s0);
        ns2 = DT<T>::pack2(-s0, -s0);
    }
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t base = c * chunk + threadIdx.x;
        uint4 qg[BQ_UNROLL], qx[BQ_UNROLL];
#pragma unroll
        for (int u = 0; u < BQ_UNROLL; ++u) {
            int64_t v = base + (int64_t)u * BQ_THREADS;
            if (v < nvec) { qg[u] = ldg_stream(gv + v); qx[u] = ldg_stream(xv + v); }
        }
        if constexpr (DT<T>::LOWP) {
            if (packed) {
#pragma unroll
                for (int u = 0; u < BQ_UNROLL; ++u) {
                    const int64_t v = base + (int64_t)u * BQ_THREADS;
                    if (v < nvec) stg_stream(ov + v, binary_bwd_vec_packed<T, CLAMPED>(qg[u], qx[u], s2, ns2, want_gs, acc));
                }
                continue;
            }
        }
#pragma unroll
        for (int u = 0; u < BQ_UNROLL; ++u) {
            int64_t v = base + (int64_t)u * BQ_THREADS;
            if (v < nvec) {
                float s = s0;
                if (smode != 0) {
                    int64_t idx = (v / inner_v) % count;
                    s = DT<T>::to_f(scale[idx]);
                    if (want_gs && idx != acc_idx) {
                        if (acc_idx >= 0) atomicAdd(gscale_out + acc_idx, acc);
                        acc = 0.f;
                        acc_idx = idx;
                    }
                }
                float eg[V], ex[V];
                DT<T>::unpack(qg[u], eg);
                DT<T>::unpack(qx[u], ex);
#pragma unroll
                for (int i = 0; i < V; ++i) eg[i] = binary_bwd_elem<T, CLAMPED>(eg[i], ex[i], s, want_gs, acc);
                stg_stream(ov + v, DT<T>::pack(eg));
            }
        }
    }
    if (want_gs) {
        if (smode == 0) {
            float t = block_sum_f(acc, red);
            if (threadIdx.x == 0) atomicAdd(gscale_out, t);
        } else if (acc_idx >= 0) {
            atomicAdd(gscale_out + acc_idx, acc);
        }
    }
}

template <typename T, bool CLAMPED>
__global__ void binary_quant_bwd_scalar_kernel(const T* gy, const T* x, const T* scale, T* gx, float* gscale_out,
                                               int64_t start, int64_t n, int64_t inner, int64_t count, int scale_f32) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool want_gs = gscale_out != nullptr;
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t idx = count == 1 ? 0 : (i / inner) % count;
        float s = count == 1 ? load_scale0<T>(scale, scale_f32) : DT<T>::to_f(scale[idx]);
        float acc = 0.f;
        gx[i] = DT<T>::from_f(binary_bwd_elem<T, CLAMPED>(DT<T>::to_f(gy[i]), DT<T>::to_f(x[i]), s, want_gs, acc));
        if (want_gs) atomicAdd(gscale_out + idx, acc);
    }
}

static inline unsigned bq_grid(int64_t work, int per_block) {
    int64_t b = (work + per_block - 1) / per_block;
    int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

template <typename T, bool CLAMPED>
static int launch_binary_fwd(const void* x, const void* scale, void* y, int64_t n, int64_t inner, int64_t count,
                             int scale_f32, cudaStream_t st) {
    constexpr int V = DT<T>::VEC;
    const int smode = (count == 1) ? 0 : 1;
    bool vec_ok = aligned16(x) && aligned16(y);
    if (smode == 1 && (inner % V) != 0) vec_ok = false;
    const int64_t nvec = vec_ok ? n / V : 0;
    if (nvec > 0)
        binary_quant_fwd_kernel<T, CLAMPED><<<bq_grid(nvec, BQ_THREADS * BQ_UNROLL), BQ_THREADS, 0, st>>>(
            (const T*)x, (const T*)scale, (T*)y, nvec, smode ? inner / V : 1, count, smode, scale_f32);
    const int64_t done = nvec * V;
    if (done < n)
        binary_quant_fwd_scalar_kernel<T, CLAMPED><<<bq_grid(n - done, 256), 256, 0, st>>>(
            (const T*)x, (const T*)scale, (T*)y, done, n, inner, count, scale_f32);
    return check_launch("bvb_binary_quant_fwd");
}

template <typename T, bool CLAMPED>
static int launch_binary_bwd(const void* gy, const void* x, const void* scale, void* gx, float* gscale_out, int64_t n,
                             int64_t inner, int64_t count, int scale_f32, cudaStream_t st) {
    constexpr int V = DT<T>::VEC;
    const int smode = (count == 1) ? 0 : 1;
    bool vec_ok = aligned16(gy) && aligned16(x) && aligned16(gx);
    if (smode == 1 && (inner % V) != 0) vec_ok = false;
    const int64_t nvec = vec_ok ? n / V : 0;
    if (nvec > 0)
        binary_quant_bwd_kernel<T, CLAMPED><<<bq_grid(nvec, BQ_THREADS * BQ_UNROLL), BQ_THREADS, 0, st>>>(
            (const T*)gy, (const T*)x, (const T*)scale, (T*)gx, gscale_out, nvec, smode ? inner / V : 1, count, smode,
            scale_f32);
    const int64_t done = nvec * V;
    if (done < n)
        binary_quant_bwd_scalar_kernel<T, CLAMPED><<<bq_grid(n - done, 256), 256, 0, st>>>(
            (const T*)gy, (const T*)x, (const T*)scale, (T*)gx, gscale_out, done, n, inner, count, scale_f32);
    return check_launch("bvb_binary_quant_bwd");
}

}  // namespace bvb

using namespace bvb;

#define BVB_CHECK_SCALE_DTYPE(name)                                                         \
    const int scale_f32 = (scale_dtype == BVB_F32 && dtype != BVB_F32) ? 1 : 0;             \
    if (scale_dtype != dtype && !(scale_f32 && scale_count == 1))                           \
        return fail(BVB_EUNSUPPORTED, name ": scale dtype must equal the tensor dtype, or be fp32 with one element");

extern "C" int bvb_binary_quant_fwd(const void* x, const void* scale, void* y, int64_t n, int64_t scale_inner,
                                    int64_t scale_count, int scale_dtype, int clamped, int dtype, void* stream) {
    BVB_CHECK_SCALE_DTYPE("bvb_binary_quant_fwd")
    if (n < 0) return fail(BVB_EINVAL, "bvb_binary_quant_fwd: negative size");
    if (n == 0) return BVB_OK;
    if (!x || !scale || !y) return fail(BVB_EINVAL, "bvb_binary_quant_fwd: null pointer");
    if (scale_inner < 1 || scale_count < 1) return fail(BVB_EINVAL, "bvb_binary_quant_fwd: bad scale broadcast pattern");
    cudaStream_t st = (cudaStream_t)stream;
    if (clamped) {
        BVB_DISPATCH_DTYPE(dtype, return (launch_binary_fwd<T, true>(x, scale, y, n, scale_inner, scale_count, scale_f32, st)));
    } else {
        BVB_DISPATCH_DTYPE(dtype, return (launch_binary_fwd<T, false>(x, scale, y, n, scale_inner, scale_count, scale_f32, st)));
    }
    return BVB_OK;
}

extern "C" int bvb_binary_quant_bwd(const void* gy, const void* x, const void* scale, void* gx, float* gscale_out,
                                    int64_t n, int64_t scale_inner, int64_t scale_count, int scale_dtype, int clamped,
                                    int dtype, void* stream) {
    BVB_CHECK_SCALE_DTYPE("bvb_binary_quant_bwd")
    if (n < 0) return fail(BVB_EINVAL, "bvb_binary_quant_bwd: negative size");
    if (scale_inner < 1 || scale_count < 1) return fail(BVB_EINVAL, "bvb_binary_quant_bwd: bad scale broadcast pattern");
    cudaStream_t st = (cudaStream_t)stream;
    if (gscale_out) {
        cudaError_t e = cudaMemsetAsync(gscale_out, 0, sizeof(float) * (size_t)scale_count, st);
        if (e != cudaSuccess) return fail(BVB_ECUDA, "bvb_binary_quant_bwd: memset: %s", cudaGetErrorString(e));
    }
    if (n == 0) return BVB_OK;
    if (!gy || !x || !scale || !gx) return fail(BVB_EINVAL, "bvb_binary_quant_bwd: null pointer");
    if (clamped) {
        BVB_DISPATCH_DTYPE(dtype, return (launch_binary_bwd<T, true>(gy, x, scale, gx, gscale_out, n, scale_inner, scale_count, scale_f32, st)));
    } else {
        BVB_DISPATCH_DTYPE(dtype, return (launch_binary_bwd<T, false>(gy, x, scale, gx, gscale_out, n, scale_inner, scale_count, scale_f32, st)));
    }
    return BVB_OK;
}
