// brevitas_b200 :: the remaining quantizer flavours of SURVEY.md 8(f) rank 3, one kernel per direction each.
//
//   general integer quantizer     DecoupledIntQuant.forward   src/brevitas/core/quant/int_base.py:100-182
//                                 IntQuant.forward            src/brevitas/core/quant/int_base.py:64-97 when the range is a
//                                 DEVICE value (learned bit-width, core/bit_width/parameter.py:23-98) or the call must not
//                                 read anything back (direct calls, CUDA-graph capture)
//       c = clamp(float_to_int(x / pre_scale + pre_zero_point), min_int, max_int)      y = (c - zero_point) * scale
//     every range input (both zero-points, min_int, max_int) is read from device memory; the backward returns, besides dx,
//     the four reductions the reference's autograd graph produces: d(pre_scale), d(scale), d(min_int), d(max_int) -- the
//     last two are what trains a learned bit-width (torch.where backward of function/ops.py:98-99).
//
//   ternary quantizer             TernaryQuant.forward        src/brevitas/core/quant/ternary.py:58-72
//       y = float(|x| > threshold * scale) * sign(x) * scale
//
// Streaming kernels: fwd 1R + 1W, bwd 2R + 1W, 16-byte accesses, reductions as per-thread fp32 partial sums that meet in
// fp64 (block reduction + one fp64 atomic per CTA and output, so the result is reproducible to fp32 accuracy).
#include "common.cuh"
#include "host.cuh"

namespace bvb {

constexpr int QV_THREADS = 256;
constexpr int QV_UNROLL = 2;

struct GenQ {
    const void* x;
    const void* gy;
    const void* pre_scale;
    const void* scale;
    const float* pre_zero_point;   // 1 element each, fp32, device
    const float* zero_point;
    const float* min_int;
    const float* max_int;
    void* out;                     // y (forward) or gx (backward)
    double* sums;                  // backward: [d pre_scale (pre_count) | d scale (post_count) | d min_int | d max_int]
    int64_t n, inner, pre_count, post_count;
    int round_mode, masked, same_scale, want_sums;
};

template <typename T>
__device__ __forceinline__ float f2i_runtime(float v, int rm) {
    switch (rm) {
        case RM_ROUND: return rintf(v);
        case RM_FLOOR: return floorf(v);
        case RM_CEIL: return ceilf(v);
        case RM_ROUND_TO_ZERO: return round_to_zero_T<T>(v);
        default: return dpu_round_T<T>(v);
    }
}

// range inputs rounded to the tensor dtype, as the reference's `.type_as(x)` / type promotion does for 0-dim operands
template <typename T>
struct GenRange {
    float pre_zp, zp, lo, hi;
    __device__ __forceinline__ explicit GenRange(const GenQ& q) {
        pre_zp = q.pre_zero_point[0];
        zp = q.zero_point[0];
        lo = DT<T>::rnd(q.min_int[0]);
        hi = DT<T>::rnd(q.max_int[0]);
    }
};

template <typename T>
__device__ __forceinline__ void gen_codes(float x, const DivBy& dv, const GenRange<T>& r, int rm, float& t1, float& t3, float& t5) {
    t1 = DT<T>::rnd(dv(x));
    const float t2 = DT<T>::rnd(fadd(t1, r.pre_zp));        // -0.0 + 0.0 = +0.0 like the reference
    t3 = f2i_runtime<T>(t2, rm);
    t5 = where_clamp(t3, r.lo, r.hi);
}

template <typename T>
__device__ __forceinline__ float gen_fwd_elem(float x, const DivBy& dv, float s, const GenRange<T>& r, int rm) {
    float t1, t3, t5;
    gen_codes<T>(x, dv, r, rm, t1, t3, t5);
    const float t6 = DT<T>::rnd(fsub(t5, r.zp));
    return fmul(t6, s);
}

struct GenAcc {
    float pre = 0.f, post = 0.f, lo = 0.f, hi = 0.f;
};

template <typename T>
__device__ __forceinline__ float gen_bwd_elem(float g, float x, const DivBy& dv, float s, const GenRange<T>& r, const GenQ& q,
                                              GenAcc& a) {
    const float gc = DT<T>::rnd(fmul(g, s));                  // MulBackward: d(c - zp)
    float d = gc;
    if (q.masked || q.want_sums) {
        float t1, t3, t5;
        gen_codes<T>(x, dv, r, q.round_mode, t1, t3, t5);
        if (q.masked) {
            // where(c1 < min, min, c1) with c1 = where(t3 > max, max, t3), walked backwards
            const bool over = t3 > r.hi;
            const float c1 = over ? r.hi : t3;
            const bool under = c1 < r.lo;
            if (under) { a.lo += gc; d = 0.f; }
            else if (over) { a.hi += gc; d = 0.f; }
        }
        if (q.want_sums) {
            const float t6 = DT<T>::rnd(fsub(t5, r.zp));
            const float back = d * DT<T>::rnd(dv(t1));                      // DivBackward wrt the divisor: -d * ((x / s) / s)
            if (q.same_scale) a.post += fmaf(g, t6, -back);                  // the two nearly cancel: difference per element
            else { a.post = fmaf(g, t6, a.post); a.pre -= back; }
        }
    }
    return dv(d);
}

// one 16-byte vector at a time: the divisions of a vector share ONE slow-path test (DivBy::div_n), the d(pre_scale) term
// (a tolerance-bound sum) multiplies by the reciprocal instead of dividing again, and RMC >= 0 fixes the rounding mode at
// compile time (round-half-even, the default of every named quantizer)
template <typename T, int RMC>
__device__ __forceinline__ float f2i_sel(float v, int rm) {
    if constexpr (RMC == RM_ROUND) return rintf(v);
    else return f2i_runtime<T>(v, rm);
}

template <typename T, int RMC, int N>
__device__ __forceinline__ void gen_fwd_n(float (&e)[N], const DivBy& dv, float s, const GenRange<T>& r, int rm) {
    float t1[N];
    dv.div_n<N>(e, t1);
    DT<T>::template rnd_n<N>(t1);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float t2 = DT<T>::rnd(fadd(t1[i], r.pre_zp));
        const float t5 = where_clamp(f2i_sel<T, RMC>(t2, rm), r.lo, r.hi);
        e[i] = fmul(DT<T>::rnd(fsub(t5, r.zp)), s);
    }
}

template <typename T, int RMC, int N>
__device__ __forceinline__ void gen_bwd_n(float (&eg)[N], const float (&ex)[N], const DivBy& dv, float inv_ps, float s,
                                          const GenRange<T>& r, const GenQ& q, GenAcc& a) {
    float d[N];
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = fmul(eg[i], s);
    DT<T>::template rnd_n<N>(d);
    if (q.masked || q.want_sums) {
        float t1[N];
        dv.div_n<N>(ex, t1);
        DT<T>::template rnd_n<N>(t1);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float t2 = DT<T>::rnd(fadd(t1[i], r.pre_zp));
            const float t3 = f2i_sel<T, RMC>(t2, q.round_mode);
            const bool over = t3 > r.hi;
            const float c1 = over ? r.hi : t3;
            const bool under = c1 < r.lo;
            const float t5 = under ? r.lo : c1;
            if (q.masked) {                                   // branch-free: selects, no divergence inside a vector
                a.lo += under ? d[i] : 0.f;
                a.hi += (over && !under) ? d[i] : 0.f;
                d[i] = (over || under) ? 0.f : d[i];
            }
            if (q.want_sums) {
                const float t6 = DT<T>::rnd(fsub(t5, r.zp));
                const float back = d[i] * (t1[i] * inv_ps);
                if (q.same_scale) a.post += fmaf(eg[i], t6, -back);
                else { a.post = fmaf(eg[i], t6, a.post); a.pre -= back; }
            }
        }
    }
    dv.div_n<N>(d, eg);
}

template <typename T>
__device__ __forceinline__ void gen_flush(const GenQ& q, GenAcc& a, int64_t idx) {
    if (!q.want_sums || idx < 0) return;
    const int64_t ip = q.pre_count == 1 ? 0 : idx, is = q.post_count == 1 ? 0 : idx;
    if (!q.same_scale && a.pre != 0.f) atomicAdd(q.sums + ip, (double)a.pre);
    if (a.post != 0.f) atomicAdd(q.sums + q.pre_count + is, (double)a.post);
    a.pre = a.post = 0.f;
}

__device__ __forceinline__ double block_sum_d(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = (lane < nw) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    __syncthreads();
    return r;
}

// VECTOR: 16-byte vectors (inner divisible by the vector length), else one element per access
template <typename T, bool BWD, bool VECTOR>
__global__ void __launch_bounds__(QV_THREADS) general_int_quant_kernel(GenQ q) {
    constexpr int V = VECTOR ? DT<T>::VEC : 1;
    __shared__ double red[32];
    const GenRange<T> r(q);
    const T* ps_p = reinterpret_cast<const T*>(q.pre_scale);
    const T* s_p = reinterpret_cast<const T*>(q.scale);
    const int64_t count = q.pre_count > q.post_count ? q.pre_count : q.post_count;
    const int64_t units = q.n / V, inner_u = q.inner / V;
    float ps0 = DT<T>::to_f(ps_p[0]), s0 = DT<T>::to_f(s_p[0]);
    GenAcc acc;
    int64_t acc_idx = -1;
    const int64_t chunk = (int64_t)QV_THREADS * QV_UNROLL;
    for (int64_t base = (int64_t)blockIdx.x * chunk + threadIdx.x; base < units; base += (int64_t)gridDim.x * chunk) {
#pragma unroll
        for (int u = 0; u < QV_UNROLL; ++u) {
            const int64_t v = base + (int64_t)u * QV_THREADS;
            if (v >= units) break;
            float ps = ps0, s = s0;
            if (count > 1) {
                const int64_t idx = (v / inner_u) % count;
                if (q.pre_count > 1) ps = DT<T>::to_f(ps_p[idx]);
                if (q.post_count > 1) s = DT<T>::to_f(s_p[idx]);
                if (BWD && idx != acc_idx) { gen_flush<T>(q, acc, acc_idx); acc_idx = idx; }
            }
            const DivBy dv(ps, DT<T>::MUL_DIV_EXACT);
            if constexpr (VECTOR) {
                float ex[V], eg[V];
                DT<T>::unpack(ldg_stream(reinterpret_cast<const uint4*>(q.x) + v), ex);
                if constexpr (BWD) {
                    DT<T>::unpack(ldg_stream(reinterpret_cast<const uint4*>(q.gy) + v), eg);
#pragma unroll
                    for (int i = 0; i < V; ++i) ex[i] = gen_bwd_elem<T>(eg[i], ex[i], dv, s, r, q, acc);
                } else {
#pragma unroll
                    for (int i = 0; i < V; ++i) ex[i] = gen_fwd_elem<T>(ex[i], dv, s, r, q.round_mode);
                }
                stg_stream(reinterpret_cast<uint4*>(q.out) + v, DT<T>::pack(ex));
            } else {
                const float xv = DT<T>::to_f(reinterpret_cast<const T*>(q.x)[v]);
                float o;
                if constexpr (BWD) o = gen_bwd_elem<T>(DT<T>::to_f(reinterpret_cast<const T*>(q.gy)[v]), xv, dv, s, r, q, acc);
                else o = gen_fwd_elem<T>(xv, dv, s, r, q.round_mode);
                reinterpret_cast<T*>(q.out)[v] = DT<T>::from_f(o);
            }
        }
    }
    if constexpr (BWD) {
        if (q.want_sums) {
            if (count > 1) {
                gen_flush<T>(q, acc, acc_idx);
            } else {
                const double pre = block_sum_d((double)acc.pre, red), post = block_sum_d((double)acc.post, red);
                if (threadIdx.x == 0) {
                    if (!q.same_scale && pre != 0.0) atomicAdd(q.sums, pre);
                    if (post != 0.0) atomicAdd(q.sums + q.pre_count, post);
                }
            }
            if (q.masked) {
                const double lo = block_sum_d((double)acc.lo, red), hi = block_sum_d((double)acc.hi, red);
                if (threadIdx.x == 0) {
                    if (lo != 0.0) atomicAdd(q.sums + q.pre_count + q.post_count, lo);
                    if (hi != 0.0) atomicAdd(q.sums + q.pre_count + q.post_count + 1, hi);
                }
            }
        }
    }
}

// Tiled flavour for long runs of one scale (one scale for the tensor, or a scale per row / plane with at least a tile of
// 16-byte vectors per run): a CTA owns a CONTIGUOUS range of tiles that never straddle a run, so the scale index, the
// divisor set-up and the scale loads are per tile instead of per vector, 4 independent vectors are in flight per thread,
// and the per-scale sums are reduced by the CTA once per run (two fp64 atomics per run and CTA instead of per thread).
constexpr int QT_UNROLL = 4;
constexpr int QT_TILE = QV_THREADS * QT_UNROLL;

template <typename T, bool BWD, int RMC>
__global__ void __launch_bounds__(QV_THREADS, BWD ? 3 : 4) general_int_quant_tiled_kernel(GenQ q, int64_t inner_u, int64_t tiles_per_run,
                                                                             int64_t total_tiles) {
    constexpr int V = DT<T>::VEC;
    __shared__ double red[32];
    const GenRange<T> r(q);
    const T* ps_p = reinterpret_cast<const T*>(q.pre_scale);
    const T* s_p = reinterpret_cast<const T*>(q.scale);
    const int64_t count = q.pre_count > q.post_count ? q.pre_count : q.post_count;
    const uint4* xv = reinterpret_cast<const uint4*>(q.x);
    const uint4* gv = reinterpret_cast<const uint4*>(q.gy);
    uint4* ov = reinterpret_cast<uint4*>(q.out);
    const int64_t t0 = total_tiles * blockIdx.x / gridDim.x, t1 = total_tiles * (blockIdx.x + 1) / gridDim.x;
    GenAcc acc;
    int64_t cur = -1;
    float ps = 1.f, s = 1.f, inv_ps = 1.f;
    DivBy dv;
    auto flush = [&]() {              // uniform over the CTA
        if (!BWD || !q.want_sums || cur < 0) return;
        const double pre = q.same_scale ? 0.0 : block_sum_d((double)acc.pre, red);
        const double post = block_sum_d((double)acc.post, red);
        if (threadIdx.x == 0) {
            if (!q.same_scale && pre != 0.0) atomicAdd(q.sums + (q.pre_count == 1 ? 0 : cur), pre);
            if (post != 0.0) atomicAdd(q.sums + q.pre_count + (q.post_count == 1 ? 0 : cur), post);
        }
        acc.pre = acc.post = 0.f;
    };
    for (int64_t t = t0; t < t1; ++t) {
        const int64_t run = t / tiles_per_run;
        const int64_t off = (t - run * tiles_per_run) * QT_TILE;
        const int64_t idx = count > 1 ? run % count : 0;
        if (idx != cur) {
            flush();
            cur = idx;
            ps = DT<T>::to_f(ps_p[q.pre_count > 1 ? idx : 0]);
            s = DT<T>::to_f(s_p[q.post_count > 1 ? idx : 0]);
            dv = DivBy(ps, DT<T>::MUL_DIV_EXACT);
            inv_ps = dv.approx_recip();
        }
        const int64_t base = run * inner_u + off;
        const int64_t len = inner_u - off;                       // vectors left in this run (the tile takes up to QT_TILE)
        uint4 qx[QT_UNROLL], qg[QT_UNROLL];
#pragma unroll
        for (int u = 0; u < QT_UNROLL; ++u) {
            const int64_t o = threadIdx.x + (int64_t)u * QV_THREADS;
            if (o < len) {
                qx[u] = ldg_stream(xv + base + o);
                if constexpr (BWD) qg[u] = ldg_stream(gv + base + o);
            }
        }
#pragma unroll
        for (int u = 0; u < QT_UNROLL; ++u) {
            const int64_t o = threadIdx.x + (int64_t)u * QV_THREADS;
            if (o < len) {
                float ex[V], eg[V];
                DT<T>::unpack(qx[u], ex);
                if constexpr (BWD) {
                    DT<T>::unpack(qg[u], eg);
                    gen_bwd_n<T, RMC, V>(eg, ex, dv, inv_ps, s, r, q, acc);
                    stg_stream(ov + base + o, DT<T>::pack(eg));
                } else {
                    gen_fwd_n<T, RMC, V>(ex, dv, s, r, q.round_mode);
                    stg_stream(ov + base + o, DT<T>::pack(ex));
                }
            }
        }
    }
    if constexpr (BWD) {
        flush();
        if (q.want_sums && q.masked) {
            const double lo = block_sum_d((double)acc.lo, red), hi = block_sum_d((double)acc.hi, red);
            if (threadIdx.x == 0) {
                if (lo != 0.0) atomicAdd(q.sums + q.pre_count + q.post_count, lo);
                if (hi != 0.0) atomicAdd(q.sums + q.pre_count + q.post_count + 1, hi);
            }
        }
    }
}

static inline unsigned qv_grid(int64_t units) {
    int64_t b = (units + (int64_t)QV_THREADS * QV_UNROLL - 1) / ((int64_t)QV_THREADS * QV_UNROLL);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

template <typename T, bool BWD>
static int launch_general(const GenQ& q, cudaStream_t st, const char* what) {
    constexpr int V = DT<T>::VEC;
    const int64_t count = q.pre_count > q.post_count ? q.pre_count : q.post_count;
    bool vec = aligned16(q.x) && aligned16(q.out) && (!BWD || aligned16(q.gy)) && q.n % V == 0;
    if (count > 1 && q.inner % V != 0) vec = false;
    if (count == 1 && q.n % V != 0 && q.n >= (int64_t)V * QT_TILE && aligned16(q.x) && aligned16(q.out) &&
        (!BWD || aligned16(q.gy))) {
        // one scale, ragged length: 16-byte vectors for the bulk, the last < V elements element-wise (the sums meet in
        // the same fp64 words)
        GenQ main = q, tail = q;
        main.n = q.n / V * V;
        const size_t skip = (size_t)main.n * sizeof(T);
        tail.n = q.n - main.n;
        tail.x = (const char*)q.x + skip;
        tail.out = (char*)q.out + skip;
        if (BWD) tail.gy = (const char*)q.gy + skip;
        const int rc = launch_general<T, BWD>(main, st, what);
        if (rc != BVB_OK) return rc;
        general_int_quant_kernel<T, BWD, false><<<1, QV_THREADS, 0, st>>>(tail);
        return check_launch(what);
    }
    const int64_t inner_u = count > 1 ? q.inner / V : q.n / V;
    if (vec && inner_u >= QT_TILE / 2) {
        const int64_t runs = (q.n / V) / inner_u;
        const int64_t tiles_per_run = (inner_u + QT_TILE - 1) / QT_TILE;
        const int64_t total = runs * tiles_per_run;
        // ONE resident wave: the tiles are dealt out in equal contiguous ranges, so a second wave would only add a tail
        auto launch = [&](auto kernel) {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, QV_THREADS, 0) != cudaSuccess || per_sm < 1)
                per_sm = 2;
            // (the forward, with nothing to reduce, measured best with 8 CTAs' worth of ranges per SM: 62 vs 74-81 us)
            int64_t grid = (int64_t)sm_count() * (BWD ? per_sm : 8);
            if (grid > total) grid = total;
            kernel<<<(unsigned)grid, QV_THREADS, 0, st>>>(q, inner_u, tiles_per_run, total);
        };
        if (q.round_mode == RM_ROUND) launch(general_int_quant_tiled_kernel<T, BWD, RM_ROUND>);
        else launch(general_int_quant_tiled_kernel<T, BWD, -1>);
    } else if (vec) general_int_quant_kernel<T, BWD, true><<<qv_grid(q.n / V), QV_THREADS, 0, st>>>(q);
    else general_int_quant_kernel<T, BWD, false><<<qv_grid(q.n), QV_THREADS, 0, st>>>(q);
    return check_launch(what);
}

// ---- ternary -------------------------------------------------------------------------------------------------------------
// fp32 only: the reference's `mask.float() * ternary_sign_ste(x)` produces fp32 whatever x is, so 16-bit inputs stay on the
// literal op sequence (their output dtype differs from their input dtype).
template <bool BWD>
__global__ void __launch_bounds__(QV_THREADS) ternary_quant_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                   const float* __restrict__ scale, float* __restrict__ out,
                                                                   double* gscale, int64_t n, int64_t nvec, float threshold) {
    __shared__ double red[32];
    const float s = scale[0];
    const float ts = fmul(s, threshold);                       // `self.threshold * scale`
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        float e[4], g[4];
        DT<float>::unpack(ldg_stream(reinterpret_cast<const uint4*>(x) + v), e);
        if (BWD) DT<float>::unpack(ldg_stream(reinterpret_cast<const uint4*>(gy) + v), g);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float m = fabsf(e[i]) > ts ? 1.f : 0.f;
            const float a = fmul(m, sign3(e[i]));
            if (BWD) {
                acc = fmaf(g[i], a, acc);
                e[i] = fmul(fmul(g[i], s), m);
            } else {
                e[i] = fmul(a, s);
            }
        }
        stg_stream(reinterpret_cast<uint4*>(out) + v, DT<float>::pack(e));
    }
    for (int64_t i = nvec * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float m = fabsf(x[i]) > ts ? 1.f : 0.f;
        const float a = fmul(m, sign3(x[i]));
        if (BWD) {
            acc = fmaf(gy[i], a, acc);
            out[i] = fmul(fmul(gy[i], s), m);
        } else {
            out[i] = fmul(a, s);
        }
    }
    if (BWD && gscale != nullptr) {
        const double t = block_sum_d((double)acc, red);
        if (threadIdx.x == 0 && t != 0.0) atomicAdd(gscale, t);
    }
}

}  // namespace bvb

using namespace bvb;

static int check_general(const char* name, int64_t n, int64_t inner, int64_t pre_count, int64_t post_count, int round_mode) {
    if (n < 0) return fail(BVB_EINVAL, "%s: negative size", name);
    if (inner < 1 || pre_count < 1 || post_count < 1) return fail(BVB_EINVAL, "%s: bad scale broadcast pattern", name);
    if (pre_count != 1 && post_count != 1 && pre_count != post_count)
        return fail(BVB_EINVAL, "%s: the two scales must share one broadcast pattern (or hold one element)", name);
    if (round_mode < BVB_ROUND || round_mode > BVB_DPU_ROUND) return fail(BVB_EINVAL, "%s: unknown round mode %d", name, round_mode);
    return BVB_OK;
}

extern "C" int bvb_general_int_quant_fwd(const void* x, const void* pre_scale, const void* scale, const float* pre_zero_point,
                                         const float* zero_point, const float* min_int, const float* max_int, void* y,
                                         int64_t n, int64_t scale_inner, int64_t pre_scale_count, int64_t scale_count,
                                         int round_mode, int dtype, void* stream) {
    const char* name = "bvb_general_int_quant_fwd";
    int rc = check_general(name, n, scale_inner, pre_scale_count, scale_count, round_mode);
    if (rc != BVB_OK) return rc;
    if (n == 0) return BVB_OK;
    if (!x || !pre_scale || !scale || !pre_zero_point || !zero_point || !min_int || !max_int || !y)
        return fail(BVB_EINVAL, "%s: null pointer", name);
    GenQ q{};
    q.x = x; q.pre_scale = pre_scale; q.scale = scale; q.pre_zero_point = pre_zero_point; q.zero_point = zero_point;
    q.min_int = min_int; q.max_int = max_int; q.out = y; q.n = n; q.inner = scale_inner; q.pre_count = pre_scale_count;
    q.post_count = scale_count; q.round_mode = round_mode;
    BVB_DISPATCH_DTYPE(dtype, return (launch_general<T, false>(q, (cudaStream_t)stream, name)));
    return BVB_OK;
}

extern "C" int64_t bvb_general_int_quant_sums(int64_t pre_scale_count, int64_t scale_count) {
    return pre_scale_count + scale_count + 2;
}

extern "C" int bvb_general_int_quant_bwd(const void* gy, const void* x, const void* pre_scale, const void* scale,
                                         const float* pre_zero_point, const float* zero_point, const float* min_int,
                                         const float* max_int, void* gx, double* sums, int64_t n, int64_t scale_inner,
                                         int64_t pre_scale_count, int64_t scale_count, int round_mode, int clamp_mode,
                                         int same_scale, int dtype, void* stream) {
    const char* name = "bvb_general_int_quant_bwd";
    int rc = check_general(name, n, scale_inner, pre_scale_count, scale_count, round_mode);
    if (rc != BVB_OK) return rc;
    if (same_scale && pre_scale_count != scale_count) return fail(BVB_EINVAL, "%s: same_scale needs equal scale counts", name);
    cudaStream_t st = (cudaStream_t)stream;
    if (sums) {
        cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)(pre_scale_count + scale_count + 2), st);
        if (e != cudaSuccess) return fail(BVB_ECUDA, "%s: memset: %s", name, cudaGetErrorString(e));
    }
    if (n == 0) return BVB_OK;
    if (!gy || !x || !pre_scale || !scale || !pre_zero_point || !zero_point || !min_int || !max_int || !gx)
        return fail(BVB_EINVAL, "%s: null pointer", name);
    GenQ q{};
    q.x = x; q.gy = gy; q.pre_scale = pre_scale; q.scale = scale; q.pre_zero_point = pre_zero_point; q.zero_point = zero_point;
    q.min_int = min_int; q.max_int = max_int; q.out = gx; q.sums = sums; q.n = n; q.inner = scale_inner;
    q.pre_count = pre_scale_count; q.post_count = scale_count; q.round_mode = round_mode;
    q.masked = clamp_mode == BVB_CLAMP_MASKED ? 1 : 0; q.same_scale = same_scale ? 1 : 0; q.want_sums = sums ? 1 : 0;
    BVB_DISPATCH_DTYPE(dtype, return (launch_general<T, true>(q, st, name)));
    return BVB_OK;
}

extern "C" int bvb_ternary_quant_fwd(const void* x, const void* scale, void* y, int64_t n, float threshold, int dtype,
                                     void* stream) {
    if (dtype != BVB_F32) return fail(BVB_EUNSUPPORTED, "bvb_ternary_quant_fwd: fp32 only (the reference's result is fp32)");
    if (n < 0) return fail(BVB_EINVAL, "bvb_ternary_quant_fwd: negative size");
    if (n == 0) return BVB_OK;
    if (!x || !scale || !y) return fail(BVB_EINVAL, "bvb_ternary_quant_fwd: null pointer");
    ternary_quant_kernel<false><<<qv_grid((n + 3) / 4), QV_THREADS, 0, (cudaStream_t)stream>>>(
        (const float*)x, nullptr, (const float*)scale, (float*)y, nullptr, n, (aligned16(x) && aligned16(y)) ? n / 4 : 0, threshold);
    return check_launch("bvb_ternary_quant_fwd");
}

extern "C" int bvb_ternary_quant_bwd(const void* gy, const void* x, const void* scale, void* gx, double* gscale, int64_t n,
                                     float threshold, int dtype, void* stream) {
    if (dtype != BVB_F32) return fail(BVB_EUNSUPPORTED, "bvb_ternary_quant_bwd: fp32 only (the reference's result is fp32)");
    if (n < 0) return fail(BVB_EINVAL, "bvb_ternary_quant_bwd: negative size");
    cudaStream_t st = (cudaStream_t)stream;
    if (gscale) {
        cudaError_t e = cudaMemsetAsync(gscale, 0, sizeof(double), st);
        if (e != cudaSuccess) return fail(BVB_ECUDA, "bvb_ternary_quant_bwd: memset: %s", cudaGetErrorString(e));
    }
    if (n == 0) return BVB_OK;
    if (!gy || !x || !scale || !gx) return fail(BVB_EINVAL, "bvb_ternary_quant_bwd: null pointer");
    ternary_quant_kernel<true><<<qv_grid((n + 3) / 4), QV_THREADS, 0, st>>>(
        (const float*)x, (const float*)gy, (const float*)scale, (float*)gx, gscale, n,
        (aligned16(x) && aligned16(gy) && aligned16(gx)) ? n / 4 : 0, threshold);
    return check_launch("bvb_ternary_quant_bwd");
}
