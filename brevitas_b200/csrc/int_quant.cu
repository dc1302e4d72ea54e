// brevitas_b200 :: integer fake-quantization, forward and STE backward.
//
// Replaces the ATen op chains issued by
//   IntQuant.to_int / IntQuant.forward          src/brevitas/core/quant/int_base.py:64-97
//   RescalingIntQuant.forward                   src/brevitas/core/quant/int.py:156-163
//   StatsFromParameterScaling / RuntimeStatsScaling + AbsMax
//                                               src/brevitas/core/scaling/runtime.py:19-102,
//                                               src/brevitas/core/stats/stats_op.py:129-141
// and the autograd graph PyTorch builds behind them.
//
// Kernels
//   int_quant_fwd_kernel          provided scale, pure streaming, 1R+1W
//   int_quant_bwd_kernel          provided scale, 2R+1W, optional d(scale) reduction
//   rows_fwd_tma_kernel           per-row abs-max + quant-dequant in ONE HBM pass: each row is staged
//                                 in shared memory by a TMA bulk copy (ring of mbarrier-tracked stages),
//                                 reduced with redux/shuffle, then quantised from shared memory
//   rows_fwd_generic_kernel       same contract for ragged / unaligned / oversized rows (two reads)
//   rows_bwd_kernel               per-row streaming backward incl. gradient through the abs-max
//   absmax_tensor_kernel (+finalise), tensor_bwd_fixup_kernel: whole-tensor statistic
#include "common.cuh"
#include "host.cuh"

namespace bvb {

constexpr int ST_THREADS = 256;   // streaming kernels
constexpr int ST_UNROLL = 4;

// A one-element scale may be an fp32 0-dim tensor while x is bf16/fp16 (fp32 quantizer modules fed
// low-precision activations): ATen's mul/div kernels then use the fp32 value in opmath, un-rounded.
template <typename T>
__device__ __forceinline__ float load_scale0(const T* scale, int scale_f32) {
    return scale_f32 ? reinterpret_cast<const float*>(scale)[0] : DT<T>::to_f(scale[0]);
}

// ------------------------------------------------------------------------------------------------------
// provided-scale forward
// smode 0: one scale for the whole tensor; 1: scale index constant within a 16-byte vector
// ------------------------------------------------------------------------------------------------------
template <typename T, int RM>
__global__ void __launch_bounds__(ST_THREADS) int_quant_fwd_kernel(
        const T* __restrict__ x, const T* __restrict__ scale, T* __restrict__ y, T* __restrict__ codes,
        int64_t nvec, int64_t inner_v, int64_t count, int smode, int scale_f32, int reverse, QParams p) {
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    uint4* yv = reinterpret_cast<uint4*>(y);
    uint4* cv = reinterpret_cast<uint4*>(codes);
    const int64_t chunk = ST_THREADS * ST_UNROLL;
    const int64_t nchunks = (nvec + chunk - 1) / chunk;
    float s0 = 1.f;
    if (smode == 0) s0 = load_scale0<T>(scale, scale_f32);
    const ScaleCtx<T> cx0(s0, !scale_f32, p, smode == 0 && PackedPath<T, RM>::value);
    with_mode<T, RM, true>(cx0.mode, [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            const int64_t cc = reverse ? (nchunks - 1 - c) : c;
            const int64_t base = cc * chunk + threadIdx.x;
            uint4 q[ST_UNROLL];
#pragma unroll
            for (int u = 0; u < ST_UNROLL; ++u) {
                int64_t v = base + (int64_t)u * ST_THREADS;
                if (v < nvec) q[u] = ldg_stream(xv + v);
            }
#pragma unroll
            for (int u = 0; u < ST_UNROLL; ++u) {
                int64_t v = base + (int64_t)u * ST_THREADS;
                if (v < nvec) {
                    uint4 kq, yq;
                    if (p.pre_relu) q[u] = relu_vec<T>(q[u]);
                    if (smode != 0) {      // scale index constant within a vector (literal formulation)
                        const ScaleCtx<T> cxv(DT<T>::to_f(scale[(v / inner_v) % count]), true, p, false);
                        yq = qdq_vec<T, RM, VM_LITERAL, !DT<T>::LOWP>(q[u], cxv, p, codes ? &kq : nullptr);
                    } else {
                        // (fp32 activations: zeros -- half of a post-ReLU tensor -- stay on the fast division path;
                        // the 16-bit formulations multiply by an exact reciprocal and have no slow path to avoid)
                        yq = qdq_vec<T, RM, MODE, !DT<T>::LOWP>(q[u], cx0, p, codes ? &kq : nullptr);
                    }
                    stg_stream(yv + v, yq);
                    if (codes) stg_stream(cv + v, kq);
                }
            }
        }
    });
}

// element-wise fallback for [start, n): any broadcast pattern, any alignment
template <typename T, int RM>
__global__ void int_quant_fwd_scalar_kernel(const T* x, const T* scale, T* y, T* codes, int64_t start, int64_t n,
                                            int64_t inner, int64_t count, int scale_f32, QParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float s = count == 1 ? load_scale0<T>(scale, scale_f32) : DT<T>::to_f(scale[(i / inner) % count]);
        float t1, t3, t5;
        float xe = DT<T>::to_f(x[i]);
        if (p.pre_relu) xe = relu_f(xe);
        to_int_chain<T, RM>(xe, DivBy(s), p, t1, t3, t5);
        float t6 = fsub(t5, p.zp);
        if (DT<T>::LOWP && p.zp_nonzero) t6 = DT<T>::rnd(t6);
        y[i] = DT<T>::from_f(fmul(t6, s));
        if (codes) codes[i] = DT<T>::from_f(t5);
    }
}

// ------------------------------------------------------------------------------------------------------
// one element of the backward
// ------------------------------------------------------------------------------------------------------
// bwd_elem<T, RM>: one element of the backward, in common.cuh (shared with bn_act_quant.cu)

// one 16-byte vector of the backward: eg[] holds the incoming gradient on entry, gx on exit
template <typename T, int RM, int N>
__device__ __forceinline__ void bwd_n(float (&eg)[N], const float (&ex)[N], const DivBy& dv, float inv_s,
                                      const QParams& p, int masked_rt, bool want_gs, float& gs_acc) {
    const bool masked = (RM & RM_MASK_KNOWN) ? ((RM & RM_MASKED) != 0) : (masked_rt != 0);
    float d[N];
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = fmul(eg[i], dv.b);                      // grad * scale
    DT<T>::template rnd_n<N>(d);
    if (masked || want_gs) {
        float t1[N];
        // (the zero-tolerant division of the forward was tried here too: the extra select per element costs more than the
        // slow path it avoids -- ReLU-folded learned-scale backward 91.3 -> 99.0 us)
        dv.div_n<N>(ex, t1);
        DT<T>::template rnd_n<N>(t1);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float t1r = t1[i];
            float t2;
            if (RM & RM_ZP0) {
                t2 = fadd(t1r, 0.f);
            } else {
                t2 = fadd(t1r, p.zp);
                if (DT<T>::LOWP && p.zp_nonzero) t2 = DT<T>::rnd(t2);
            }
            const float t3 = float_to_int<T, RM>(t2);
            // the clamped code only feeds the d(scale) sum here (a zero's sign is invisible), and the clamp mask
            // "not (t3 > qmax) and not (t3 < qmin)" is "the clamp left t3 unchanged, or t3 is NaN": one ordered
            // not-equal compare instead of two compares on the bounds
            const float t5 = minmax_clamp(t3, p.qmin, p.qmax);
            if (masked) d[i] = (t3 < t5 || t3 > t5) ? 0.f : d[i];
            if (want_gs) {
                const float t6 = (RM & RM_ZP0) ? t5 : fsub(t5, p.zp);
                gs_acc += fmaf(eg[i], t6, -(d[i] * (t1r * inv_s)));      // per-element difference first (see bwd_elem)
            }
        }
    }
    dv.div_n<N>(d, eg);                                                        // grad / scale (rounded at store)
}

// One 16-byte vector of the backward; returns gx.  bf16 / fp16 with the default modes use packed-pair arithmetic
// (see qdq_vec in common.cuh): grad*scale is one HMUL2 (the fp32 product of two T values is exact, hence a single
// rounding), the clamp mask is two packed compares of t2 = rnd_T(x/s)+0 against thresholds the host derived from
// round()'s monotonicity (round(v) > qmax <=> v > thr_hi), and the integer code for d(scale) is
// round(clamp(t2)) by magic-number adds.  Element-wise results are bit-identical to the literal sequence; the
// d(scale) partial sum only changes its summation order (tolerance-bound by contract).
template <typename T, int RM, int MODE>
__device__ __forceinline__ uint4 bwd_vec(const uint4& qg, const uint4& qx, const ScaleCtx<T>& cx, const QParams& p,
                                         int masked_rt, bool want_gs, float& gs_acc) {
    constexpr int V = DT<T>::VEC;
    if constexpr (MODE != VM_LITERAL && PackedPath<T, RM>::value) {
        const bool masked = (RM & RM_MASK_KNOWN) ? ((RM & RM_MASKED) != 0) : (masked_rt != 0);
        const uint32_t qgw[4] = {qg.x, qg.y, qg.z, qg.w};
        uint32_t d[V / 2];
#pragma unroll
        for (int j = 0; j < V / 2; ++j) d[j] = DT<T>::p_mul(qgw[j], cx.s2);           // rnd_T(grad * scale)
        float dm[V];
        if (masked || want_gs) {
            float ex[V], t1[V];
            DT<T>::unpack(qx, ex);
            cx.dv.template div_n<V>(ex, t1);
            float s_gt = 0.f, s_dt = 0.f;
#pragma unroll
            for (int j = 0; j < V / 2; ++j) {
                const uint32_t t1p = DT<T>::pack2(t1[2 * j], t1[2 * j + 1]);
                const uint32_t t2 = DT<T>::p_add(t1p, 0u);
                if (masked) d[j] &= ~(DT<T>::p_gt_mask(t2, p.pk_thr_hi) | DT<T>::p_lt_mask(t2, p.pk_thr_lo));
                DT<T>::p_unpack(d[j], dm[2 * j], dm[2 * j + 1]);
                if (want_gs) {
                    const uint32_t c = DT<T>::p_max_nan(DT<T>::p_min_nan(t2, p.pk_hi), p.pk_lo);
                    float k0, k1, g0, g1, a0, a1;
                    DT<T>::p_rint_f(c, k0, k1);                                        // t6 (zero-point is 0)
                    DT<T>::p_unpack(qgw[j], g0, g1);
                    DT<T>::p_unpack(t1p, a0, a1);
                    s_gt = fmaf(g0, k0, s_gt); s_gt = fmaf(g1, k1, s_gt);
                    s_dt = fmaf(dm[2 * j], a0, s_dt); s_dt = fmaf(dm[2 * j + 1], a1, s_dt);
                }
            }
            // d(scale) += sum g*t6 - sum d*((x/s)/s)
            if (want_gs) gs_acc += fmaf(-cx.inv_s, s_dt, s_gt);
        } else {
#pragma unroll
            for (int j = 0; j < V / 2; ++j) DT<T>::p_unpack(d[j], dm[2 * j], dm[2 * j + 1]);
        }
        float gx[V];
        cx.dv.template div_n<V>(dm, gx);                                               // grad / scale
        return DT<T>::pack(gx);
    } else {
        float eg[V], ex[V];
        DT<T>::unpack(qg, eg);
        DT<T>::unpack(qx, ex);
        bwd_n<T, RM, V>(eg, ex, cx.dv, cx.inv_s, p, masked_rt, want_gs, gs_acc);
        return DT<T>::pack(eg);
    }
}

// provided-scale backward; gscale_out (nullable) accumulated with float atomics
template <typename T, int RM>
__global__ void __launch_bounds__(ST_THREADS) int_quant_bwd_kernel(
        const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ scale, T* __restrict__ gx,
        float* gscale_out, int64_t nvec, int64_t inner_v, int64_t count, int smode, int scale_f32, int masked, QParams p) {
    __shared__ float red[32];
    const uint4* gv = reinterpret_cast<const uint4*>(gy);
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    uint4* ov = reinterpret_cast<uint4*>(gx);
    const bool want_gs = gscale_out != nullptr;
    const int64_t chunk = ST_THREADS * ST_UNROLL;
    const int64_t nchunks = (nvec + chunk - 1) / chunk;
    float s0 = 1.f;
    if (smode == 0) s0 = load_scale0<T>(scale, scale_f32);
    const ScaleCtx<T> cx0(s0, !scale_f32, p, smode == 0 && PackedPath<T, RM>::value);
    float acc = 0.f;          // smode 0: block-wide; smode 1: run of equal scale indices
    int64_t acc_idx = -1;
    with_mode<T, RM, false>(cx0.mode, [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            const int64_t base = c * chunk + threadIdx.x;
            uint4 qg[ST_UNROLL], qx[ST_UNROLL];
#pragma unroll
            for (int u = 0; u < ST_UNROLL; ++u) {
                int64_t v = base + (int64_t)u * ST_THREADS;
                if (v < nvec) { qg[u] = ldg_stream(gv + v); qx[u] = ldg_stream(xv + v); }
            }
#pragma unroll
            for (int u = 0; u < ST_UNROLL; ++u) {
                int64_t v = base + (int64_t)u * ST_THREADS;
                if (v < nvec) {
                    if (smode != 0) {
                        const int64_t idx = (v / inner_v) % count;
                        const ScaleCtx<T> cxv(DT<T>::to_f(scale[idx]), true, p, false);
                        if (want_gs && idx != acc_idx) {
                            if (acc_idx >= 0) atomicAdd(gscale_out + acc_idx, acc);
                            acc = 0.f;
                            acc_idx = idx;
                        }
                        uint4 o = bwd_vec<T, RM, VM_LITERAL>(qg[u], p.pre_relu ? relu_vec<T>(qx[u]) : qx[u], cxv, p, masked,
                                                             want_gs, acc);
                        stg_stream(ov + v, p.pre_relu ? relu_grad_vec<T>(o, qx[u]) : o);
                    } else {
                        uint4 o = bwd_vec<T, RM, MODE>(qg[u], p.pre_relu ? relu_vec<T>(qx[u]) : qx[u], cx0, p, masked, want_gs,
                                                       acc);
                        stg_stream(ov + v, p.pre_relu ? relu_grad_vec<T>(o, qx[u]) : o);
                    }
                }
            }
        }
    });
    if (want_gs) {
        if (smode == 0) {
            float t = block_sum_f(acc, red);
            if (threadIdx.x == 0) atomicAdd(gscale_out, t);
        } else if (acc_idx >= 0) {
            atomicAdd(gscale_out + acc_idx, acc);
        }
    }
}

template <typename T, int RM>
__global__ void int_quant_bwd_scalar_kernel(const T* gy, const T* x, const T* scale, T* gx, float* gscale_out,
                                            int64_t start, int64_t n, int64_t inner, int64_t count, int scale_f32,
                                            int masked, QParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool want_gs = gscale_out != nullptr;
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t idx = count == 1 ? 0 : (i / inner) % count;
        float s = count == 1 ? load_scale0<T>(scale, scale_f32) : DT<T>::to_f(scale[idx]);
        float acc = 0.f;
        const DivBy dv(s);
        const float x0 = DT<T>::to_f(x[i]);
        float r = bwd_elem<T, RM>(DT<T>::to_f(gy[i]), p.pre_relu ? relu_f(x0) : x0, dv, dv.approx_recip(), p, masked, want_gs,
                                  acc);
        if (p.pre_relu && x0 <= 0.f) r = 0.f;
        gx[i] = DT<T>::from_f(r);
        if (want_gs) atomicAdd(gscale_out + idx, acc);
    }
}

// ------------------------------------------------------------------------------------------------------
// provided scale with a channel / row broadcast: element i uses scale[(i / inner) % count], i.e. the tensor is a
// sequence of contiguous "planes" of `inner` elements sharing one scale ([O,1,..] rows, [1,C,1,1] NCHW channels,
// [B,T,1] tokens).  A group of G threads (a warp for small planes, the CTA for large ones) owns a plane: the
// divisor set-up is hoisted per plane and d(scale) is reduced inside the group, then ONE atomic per plane.
// ------------------------------------------------------------------------------------------------------
constexpr int PL_THREADS = 256;
constexpr int PL_UNROLL = 4;

template <typename T, int RM, bool BWD>
__global__ void __launch_bounds__(PL_THREADS, BWD ? 2 : 4) int_quant_planes_kernel(
        const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ scale, T* __restrict__ out,
        T* __restrict__ codes, float* gscale_out, int64_t nplanes, int64_t inner, int64_t count, int group,
        int vec_ok, int masked, QParams p) {
    constexpr int V = DT<T>::VEC;
    __shared__ float red[32];
    const int lane = threadIdx.x & 31;
    const int gid = (group == 32) ? (threadIdx.x >> 5) : 0;               // group index inside the CTA
    const int gtid = (group == 32) ? lane : threadIdx.x;                    // thread index inside the group
    const int groups_per_cta = PL_THREADS / group;
    const bool want_gs = BWD && gscale_out != nullptr;
    for (int64_t pl = (int64_t)blockIdx.x * groups_per_cta + gid; pl < nplanes;
         pl += (int64_t)gridDim.x * groups_per_cta) {
        const int64_t sidx = pl % count;
        const ScaleCtx<T> cx(DT<T>::to_f(scale[sidx]), true, p, PackedPath<T, RM>::value);
        const int64_t base = pl * inner;
        float acc = 0.f;
        if (vec_ok) {
            const int64_t nv = inner / V;
            const uint4* xv = reinterpret_cast<const uint4*>(x + base);
            const uint4* gv = reinterpret_cast<const uint4*>(BWD ? gy + base : x + base);
            uint4* ov = reinterpret_cast<uint4*>(out + base);
            uint4* cv = codes ? reinterpret_cast<uint4*>(codes + base) : nullptr;
            // PL_UNROLL independent 16-byte loads in flight per thread before any arithmetic (ncu r01b: one load at
            // a time left this kernel latency-bound at 42 % issue utilisation and 4.6 TB/s)
            with_mode<T, RM, !BWD>(cx.mode, [&](auto mode_tag) {
                constexpr int MODE = decltype(mode_tag)::value;
                for (int64_t v0 = gtid; v0 < nv; v0 += (int64_t)group * PL_UNROLL) {
                    uint4 qx[PL_UNROLL], qg[PL_UNROLL];
#pragma unroll
                    for (int u = 0; u < PL_UNROLL; ++u) {
                        const int64_t v = v0 + (int64_t)u * group;
                        if (v < nv) {
                            qx[u] = ldg_stream(xv + v);
                            if (BWD) qg[u] = ldg_stream(gv + v);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PL_UNROLL; ++u) {
                        const int64_t v = v0 + (int64_t)u * group;
                        if (v < nv) {
                            if (BWD) {
                                uint4 o = bwd_vec<T, RM, MODE>(qg[u], p.pre_relu ? relu_vec<T>(qx[u]) : qx[u], cx, p, masked,
                                                               want_gs, acc);
                                stg_stream(ov + v, p.pre_relu ? relu_grad_vec<T>(o, qx[u]) : o);
                            } else {
                                uint4 kq;
                                if (p.pre_relu) qx[u] = relu_vec<T>(qx[u]);
                                const uint4 yq = qdq_vec<T, RM, MODE, !DT<T>::LOWP>(qx[u], cx, p, cv ? &kq : nullptr);
                                stg_stream(ov + v, yq);
                                if (cv) stg_stream(cv + v, kq);
                            }
                        }
                    }
                }
            });
        } else {
            for (int64_t j = gtid; j < inner; j += group) {
                const float x0 = DT<T>::to_f(x[base + j]);
                float ex[1] = {p.pre_relu ? relu_f(x0) : x0};
                if (BWD) {
                    float eg[1] = {DT<T>::to_f(gy[base + j])};
                    bwd_n<T, RM, 1>(eg, ex, cx.dv, cx.inv_s, p, masked, want_gs, acc);
                    if (p.pre_relu && x0 <= 0.f) eg[0] = 0.f;
                    out[base + j] = DT<T>::from_f(eg[0]);
                } else {
                    float k[1];
                    quant_dequant_n<T, RM, 1>(ex, cx.dv, p, codes ? k : nullptr);
                    out[base + j] = DT<T>::from_f(ex[0]);
                    if (codes) codes[base + j] = DT<T>::from_f(k[0]);
                }
            }
        }
        if (want_gs) {
            if (group == 32) {
                const float t = warp_sum_f(acc);
                if (lane == 0) atomicAdd(gscale_out + sidx, t);
            } else {
                const float t = block_sum_f(acc, red);
                if (threadIdx.x == 0) atomicAdd(gscale_out + sidx, t);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// provided scale per CHANNEL of a channels-last tensor (NHWC activations with a [1,C,1,1] scale: the learned
// per-channel activation scales of MobileNetV1): element i uses scale[i % C].  Every 16-byte vector spans V
// consecutive channels, and with C | blockDim * V a thread meets the SAME V channels at every grid-stride step: the
// V divisor set-ups are hoisted out of the loop like a row's single one, and d(scale) is accumulated in V registers,
// reduced through shared memory, then C atomics per CTA.
// ------------------------------------------------------------------------------------------------------
constexpr int CL_THREADS = 256;
constexpr int CL_UNROLL = 4;
constexpr int CL_MAX_C = 2048;

// one vector of the channels-last kernels through the literal op sequence with the divisors rebuilt on the spot: the
// never-taken branch of the packed variant, kept out of line so that it costs the hot loop no registers
template <typename T> struct ClSlow { uint4 o, k; float dacc[DT<T>::VEC]; };
template <typename T, int RM, bool BWD>
__device__ __noinline__ ClSlow<T> chanlast_slow_vec(uint4 qx, uint4 qg, uint4 qs, QParams p, int masked_rt, bool want_gs) {
    constexpr int V = DT<T>::VEC;
    ClSlow<T> r;
    float ex[V], eg[V], sv[V], eo[V], ek[V];
    DT<T>::unpack(qx, ex);
    DT<T>::unpack(qg, eg);
    DT<T>::unpack(qs, sv);
#pragma unroll 1
    for (int i = 0; i < V; ++i) {
        const DivBy d(sv[i], DT<T>::MUL_DIV_EXACT);
        float a = 0.f;
        if (BWD) {
            eo[i] = bwd_elem<T, RM>(eg[i], ex[i], d, d.approx_recip(), p, masked_rt, want_gs, a);
            ek[i] = 0.f;
        } else {
            float t1, t3, t5;
            to_int_chain<T, RM>(ex[i], d, p, t1, t3, t5);
            float t6 = fsub(t5, p.zp);
            if (DT<T>::LOWP && p.zp_nonzero) t6 = DT<T>::rnd(t6);
            eo[i] = fmul(t6, d.b);
            ek[i] = t5;
        }
        r.dacc[i] = a;
    }
    r.o = DT<T>::pack(eo);
    r.k = DT<T>::pack(ek);
    return r;
}

// PK: packed-pair formulation (qdq_vec / bwd_vec) with one scale per LANE, for bf16 with the default quantizers when
// the host knows the bounds qualify (QParams::pk_ok): the thread keeps V reciprocals and V/2 packed scale pairs instead
// of V full divisor set-ups.  A thread whose scales leave the exact multiply-by-reciprocal window (|s| outside
// [2^-40, 2^40)) rebuilds the literal divisor per element instead -- correct, slow, and never seen in practice.
template <typename T, int RM, bool BWD, bool PK = false>
__global__ void __launch_bounds__(CL_THREADS) int_quant_chanlast_kernel(
        const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ scale, T* __restrict__ out,
        T* __restrict__ codes, float* gscale_out, int64_t nvec, int C, int masked_rt, QParams p) {
    constexpr int V = DT<T>::VEC;
    constexpr int NDV = PK ? 1 : V;
    __shared__ float sacc[BWD ? CL_MAX_C : 1];
    const bool want_gs = BWD && gscale_out != nullptr;
    const bool masked = (RM & RM_MASK_KNOWN) ? ((RM & RM_MASKED) != 0) : (masked_rt != 0);
    (void)masked;
    const int c0 = (int)(((int64_t)threadIdx.x * V) % C);
    if (want_gs) {
        for (int c = threadIdx.x; c < C; c += CL_THREADS) sacc[c] = 0.f;
        __syncthreads();
    }
    const uint4 qs = *reinterpret_cast<const uint4*>(scale + c0);          // c0 is a multiple of V: 16-byte aligned
    float acc[V], inv_s[V];                         // PK: inv_s doubles as the exact reciprocal
    DivBy dvs[NDV];                                 // !PK: the thread's V divisor set-ups, built once
    bool lanes_ok = true;
    {
        float sv[V];
        DT<T>::unpack(qs, sv);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            acc[i] = 0.f;
            const DivBy d(sv[i], DT<T>::MUL_DIV_EXACT);
            inv_s[i] = d.approx_recip();
            lanes_ok = lanes_ok && d.mul_only != 0u;
            if constexpr (!PK) dvs[i] = d;
        }
    }
    const uint32_t s2[4] = {qs.x, qs.y, qs.z, qs.w};                        // PK: the scales as packed pairs of T
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const uint4* gv = reinterpret_cast<const uint4*>(BWD ? gy : x);
    uint4* ov = reinterpret_cast<uint4*>(out);
    uint4* cv = reinterpret_cast<uint4*>(codes);

    auto literal = [&](const uint4& qx, const uint4& qg, int64_t v) -> uint4 {
        if constexpr (PK) {
            const ClSlow<T> r = chanlast_slow_vec<T, RM, BWD>(qx, qg, qs, p, masked_rt, want_gs);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] += r.dacc[i];
            if (!BWD && codes) stg_stream(cv + v, r.k);
            return r.o;
        } else {
            float ex[V], eo[V], ek[V];
            DT<T>::unpack(qx, ex);
            if (BWD) {
                float eg[V];
                DT<T>::unpack(qg, eg);
#pragma unroll
                for (int i = 0; i < V; ++i)
                    eo[i] = bwd_elem<T, RM>(eg[i], ex[i], dvs[i], inv_s[i], p, masked_rt, want_gs, acc[i]);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    float t1, t3, t5;
                    to_int_chain<T, RM>(ex[i], dvs[i], p, t1, t3, t5);
                    float t6 = fsub(t5, p.zp);
                    if (DT<T>::LOWP && p.zp_nonzero) t6 = DT<T>::rnd(t6);
                    eo[i] = fmul(t6, dvs[i].b);
                    ek[i] = t5;
                }
                if (codes) stg_stream(cv + v, DT<T>::pack(ek));
            }
            return DT<T>::pack(eo);
        }
    };

    auto body = [&](uint4 qx, const uint4& qg, int64_t v) {
        const uint4 qx0 = qx;
        if (p.pre_relu) qx = relu_vec<T>(qx);
        uint4 o;
        if constexpr (PK) {
            if (lanes_ok) {
                float ex[V], t1[V];
                DT<T>::unpack(qx, ex);
#pragma unroll
                for (int i = 0; i < V; ++i) t1[i] = fmul(ex[i], inv_s[i]);             // x / s (exact after the rounding)
                if constexpr (!BWD) {
                    uint32_t w[V / 2], k[V / 2];
#pragma unroll
                    for (int j = 0; j < V / 2; ++j) {
                        const uint32_t t2 = DT<T>::p_add(DT<T>::pack2(t1[2 * j], t1[2 * j + 1]), 0u);
                        const uint32_t c = DT<T>::p_max_nan(DT<T>::p_min_nan(t2, p.pk_hi), p.pk_lo_pre);
                        uint32_t r = DT<T>::p_rint(c);
                        if (p.pk_lo_zero) r &= ~DT<T>::p_lt_mask(r, 0u);
                        k[j] = r;
                        w[j] = DT<T>::p_mul(r, s2[j]);
                    }
                    if (codes) stg_stream(cv + v, make_uint4(k[0], k[1], k[2], k[3]));
                    o = make_uint4(w[0], w[1], w[2], w[3]);
                } else {
                    const uint32_t qgw[4] = {qg.x, qg.y, qg.z, qg.w};
                    float gxe[V];
#pragma unroll
                    for (int j = 0; j < V / 2; ++j) {
                        uint32_t d = DT<T>::p_mul(qgw[j], s2[j]);                          // rnd_T(grad * scale)
                        const uint32_t t1p = DT<T>::pack2(t1[2 * j], t1[2 * j + 1]);
                        const uint32_t t2 = DT<T>::p_add(t1p, 0u);
                        if (masked) d &= ~(DT<T>::p_gt_mask(t2, p.pk_thr_hi) | DT<T>::p_lt_mask(t2, p.pk_thr_lo));
                        float d0, d1;
                        DT<T>::p_unpack(d, d0, d1);
                        if (want_gs) {
                            const uint32_t c = DT<T>::p_max_nan(DT<T>::p_min_nan(t2, p.pk_hi), p.pk_lo);
                            float k0, k1, g0, g1, a0, a1;
                            DT<T>::p_rint_f(c, k0, k1);                                    // t6 (zero-point is 0)
                            DT<T>::p_unpack(qgw[j], g0, g1);
                            DT<T>::p_unpack(t1p, a0, a1);
                            acc[2 * j] = fmaf(g0, k0, acc[2 * j]);
                            acc[2 * j] = fmaf(-d0, fmul(a0, inv_s[2 * j]), acc[2 * j]);
                            acc[2 * j + 1] = fmaf(g1, k1, acc[2 * j + 1]);
                            acc[2 * j + 1] = fmaf(-d1, fmul(a1, inv_s[2 * j + 1]), acc[2 * j + 1]);
                        }
                        gxe[2 * j] = fmul(d0, inv_s[2 * j]);                               // grad / scale
                        gxe[2 * j + 1] = fmul(d1, inv_s[2 * j + 1]);
                    }
                    o = DT<T>::pack(gxe);
                }
            } else {
                o = literal(qx, qg, v);
            }
        } else {
            o = literal(qx, qg, v);
        }
        if (BWD && p.pre_relu) o = relu_grad_vec<T>(o, qx0);
        stg_stream(ov + v, o);
    };

    // grid-stride over vectors; stride * V is a multiple of C, so a thread keeps its V channels.  Full chunks issue
    // all their loads before the first use (no predication), the ragged end goes one vector at a time.
    constexpr int U = BWD ? CL_UNROLL / 2 : CL_UNROLL;       // the backward loads two vectors per step
    const int64_t stride = (int64_t)gridDim.x * CL_THREADS;
    int64_t v0 = (int64_t)blockIdx.x * CL_THREADS + threadIdx.x;
    for (; v0 - threadIdx.x + (int64_t)(U - 1) * stride + CL_THREADS <= nvec; v0 += stride * U) {
        uint4 qx[U], qg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            qx[u] = ldg_stream(xv + v0 + (int64_t)u * stride);
            qg[u] = BWD ? ldg_stream(gv + v0 + (int64_t)u * stride) : qx[u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) body(qx[u], qg[u], v0 + (int64_t)u * stride);
    }
    for (; v0 < nvec; v0 += stride) {
        const uint4 qx1 = ldg_stream(xv + v0);
        const uint4 qg1 = BWD ? ldg_stream(gv + v0) : qx1;
        body(qx1, qg1, v0);
    }
    if (want_gs) {
#pragma unroll
        for (int i = 0; i < V; ++i) atomicAdd(&sacc[c0 + i], acc[i]);
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += CL_THREADS) atomicAdd(gscale_out + c, sacc[c]);
    }
}

// ------------------------------------------------------------------------------------------------------
// provided scale AND a tensor-valued zero-point (asymmetric quantizers: StatsFromParameterZeroPoint,
// ParameterFromRuntimeZeroPoint, ParameterZeroPoint -- the ShiftedUint8* quantizers): element i uses
// scale[(i / inner) % count] and zero_point[(i / inner) % count].  A group of threads owns a plane of `inner`
// elements (for one scale the tensor is cut into planes of ZP_CHUNK elements); the backward also reduces
//   d(scale)      = sum g*(q - zp) - d*((x/s)/s)
//   d(zero_point) = sum (d - g*s)            (the clamp mask removes d, the "- zp" of the dequantization always counts)
// per plane.  Literal op sequence (the zero-point steps need their own roundings in 16-bit dtypes).
// ------------------------------------------------------------------------------------------------------
constexpr int ZP_CHUNK = 16384;

template <typename T, int RM, bool BWD>
__global__ void __launch_bounds__(PL_THREADS) int_quant_zpt_kernel(
        const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ scale, const T* __restrict__ zero_point,
        T* __restrict__ out, float* gscale_out, float* gzp_out, int64_t n, int64_t inner, int64_t count, int scale_f32,
        int zp_f32, int masked, QParams p0) {
    constexpr int V = DT<T>::VEC;
    __shared__ float red[32];
    const int64_t nplanes = (n + inner - 1) / inner;
    const bool want_g = BWD && gscale_out != nullptr;
    for (int64_t pl = blockIdx.x; pl < nplanes; pl += gridDim.x) {
        const int64_t sidx = pl % count;
        const float s = count == 1 ? load_scale0<T>(scale, scale_f32) : DT<T>::to_f(scale[sidx]);
        QParams p = p0;
        // a 0-dim fp32 zero-point next to a 16-bit tensor is rounded to the tensor dtype by ATen's add / sub
        p.zp = DT<T>::rnd(count == 1 ? load_scale0<T>(zero_point, zp_f32) : DT<T>::to_f(zero_point[sidx]));
        p.zp_nonzero = 1;
        const DivBy dv(s);
        const float inv_s = dv.approx_recip();
        const int64_t base = pl * inner;
        const int64_t len = (n - base < inner) ? (n - base) : inner;
        float acc_s = 0.f, acc_z = 0.f;
        const bool vec = ((reinterpret_cast<uintptr_t>(x + base) | reinterpret_cast<uintptr_t>(out + base) |
                           (BWD ? reinterpret_cast<uintptr_t>(gy + base) : 0)) & 15u) == 0;
        const int64_t nv = vec ? len / V : 0;
        const uint4* xv = reinterpret_cast<const uint4*>(x + base);
        const uint4* gv = reinterpret_cast<const uint4*>(BWD ? gy + base : x + base);
        uint4* ov = reinterpret_cast<uint4*>(out + base);
        auto one = [&](float g, float xe) -> float {
            if (!BWD) return quant_dequant<T, RM>(xe, dv, p);
            const float gsv = DT<T>::rnd(fmul(g, dv.b));
            float t1, t3, t5;
            to_int_chain<T, RM>(xe, dv, p, t1, t3, t5);
            float d = gsv;
            if (masked) d = (!(t3 > p.qmax) && !(t3 < p.qmin)) ? gsv : 0.f;
            if (want_g) {
                const float t6 = DT<T>::rnd(fsub(t5, p.zp));
                acc_s = fmaf(g, t6, acc_s);
                acc_s = fmaf(-d, t1 * inv_s, acc_s);
                acc_z += d - gsv;
            }
            return dv(d);
        };
        for (int64_t v0 = threadIdx.x; v0 < nv; v0 += (int64_t)PL_THREADS * PL_UNROLL) {
            uint4 qx[PL_UNROLL], qg[PL_UNROLL];
#pragma unroll
            for (int u = 0; u < PL_UNROLL; ++u) {
                const int64_t v = v0 + (int64_t)u * PL_THREADS;
                if (v < nv) {
                    qx[u] = ldg_stream(xv + v);
                    if (BWD) qg[u] = ldg_stream(gv + v);
                }
            }
#pragma unroll
            for (int u = 0; u < PL_UNROLL; ++u) {
                const int64_t v = v0 + (int64_t)u * PL_THREADS;
                if (v < nv) {
                    float ex[V], eg[V];
                    DT<T>::unpack(qx[u], ex);
                    if (BWD) DT<T>::unpack(qg[u], eg);
#pragma unroll
                    for (int i = 0; i < V; ++i) ex[i] = one(BWD ? eg[i] : 0.f, ex[i]);
                    stg_stream(ov + v, DT<T>::pack(ex));
                }
            }
        }
        for (int64_t j = nv * V + threadIdx.x; j < len; j += PL_THREADS)
            out[base + j] = DT<T>::from_f(one(BWD ? DT<T>::to_f(gy[base + j]) : 0.f, DT<T>::to_f(x[base + j])));
        if (want_g) {
            const float ts = block_sum_f(acc_s, red);
            const float tz = block_sum_f(acc_z, red);
            if (threadIdx.x == 0) {
                atomicAdd(gscale_out + sidx, ts);
                atomicAdd(gzp_out + sidx, tz);
            }
        }
    }
}

template <typename T, int RM, bool BWD>
static int launch_int_quant_zpt(const void* gy, const void* x, const void* scale, const void* zp, void* out,
                                float* gscale_out, float* gzp_out, int64_t n, int64_t inner, int64_t count, int scale_f32,
                                int zp_f32, int masked, const QParams& p, cudaStream_t st) {
    if (BWD && gscale_out) {
        cudaError_t e = cudaMemsetAsync(gscale_out, 0, sizeof(float) * (size_t)count, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(gzp_out, 0, sizeof(float) * (size_t)count, st);
        if (e != cudaSuccess) return fail(BVB_ECUDA, "bvb_int_quant_zpt_bwd: memset: %s", cudaGetErrorString(e));
    }
    const int64_t plane = count == 1 ? (int64_t)ZP_CHUNK : inner;
    const int64_t nplanes = (n + plane - 1) / plane;
    int64_t grid = nplanes;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    int_quant_zpt_kernel<T, RM, BWD><<<(unsigned)grid, PL_THREADS, 0, st>>>(
        (const T*)gy, (const T*)x, (const T*)scale, (const T*)zp, (T*)out, gscale_out, gzp_out, n, plane, count, scale_f32,
        zp_f32, masked, p);
    return check_launch(BWD ? "bvb_int_quant_zpt_bwd" : "bvb_int_quant_zpt_fwd");
}

// ------------------------------------------------------------------------------------------------------
// integer export: clamp(round(x / scale + zero_point), qmin, qmax) stored in a real integer dtype
// (IntQuant.to_int + the cast of QuantTensor.int(), quant_tensor/__init__.py:174-187).  1 read of T, 1 write of
// 1 or 4 bytes per element.  OUT: int8_t / uint8_t / int32_t.  The codes are integer-valued floats inside the
// output range by construction, so the conversion is exact; NaN codes are stored as 0 (cvt.rni semantics).
// ------------------------------------------------------------------------------------------------------
template <typename OUT> struct IntPack;
template <> struct IntPack<int32_t> {
    template <int N> __device__ __forceinline__ static void store(int32_t* dst, const int (&c)[N]) {
#pragma unroll
        for (int i = 0; i < N; i += 4) *reinterpret_cast<int4*>(dst + i) = make_int4(c[i], c[i + 1], c[i + 2], c[i + 3]);
    }
};
template <typename B> struct IntPack8 {
    template <int N> __device__ __forceinline__ static void store(B* dst, const int (&c)[N]) {
        uint32_t w[N / 4];
#pragma unroll
        for (int i = 0; i < N / 4; ++i)
            w[i] = (uint32_t)(c[4 * i] & 0xff) | ((uint32_t)(c[4 * i + 1] & 0xff) << 8) |
                   ((uint32_t)(c[4 * i + 2] & 0xff) << 16) | ((uint32_t)(c[4 * i + 3] & 0xff) << 24);
        if (N == 4) *reinterpret_cast<uint32_t*>(dst) = w[0];
        else *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[N / 4 - 1]);
    }
};
template <> struct IntPack<int8_t> : IntPack8<int8_t> {};
template <> struct IntPack<uint8_t> : IntPack8<uint8_t> {};

template <typename T, int RM, typename OUT>
__global__ void __launch_bounds__(ST_THREADS) to_int_kernel(const T* __restrict__ x, const T* __restrict__ scale,
                                                            OUT* __restrict__ out, int64_t n, int64_t nvec, int64_t inner,
                                                            int64_t count, int scale_f32, QParams p) {
    constexpr int V = DT<T>::VEC;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const int64_t stride = (int64_t)gridDim.x * ST_THREADS;
    float s_one = 1.f;
    if (count == 1) s_one = load_scale0<T>(scale, scale_f32);
    const DivBy dv_one(s_one);
    for (int64_t v0 = (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; v0 < nvec; v0 += stride * ST_UNROLL) {
        uint4 q[ST_UNROLL];
#pragma unroll
        for (int u = 0; u < ST_UNROLL; ++u) {
            const int64_t v = v0 + (int64_t)u * stride;
            if (v < nvec) q[u] = ldg_stream(xv + v);
        }
#pragma unroll
        for (int u = 0; u < ST_UNROLL; ++u) {
            const int64_t v = v0 + (int64_t)u * stride;
            if (v < nvec) {
                // the host only takes this path when a vector never straddles two scales (inner % V == 0 or one scale)
                const DivBy dv = count == 1 ? dv_one : DivBy(DT<T>::to_f(scale[((v * V) / inner) % count]));
                float e[V];
                int c[V];
                DT<T>::unpack(q[u], e);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    float t1, t3, t5;
                    to_int_chain<T, RM>(e[i], dv, p, t1, t3, t5);
                    c[i] = __float2int_rn(t5);
                }
                IntPack<OUT>::template store<V>(out + v * V, c);
            }
        }
    }
    // ragged tail / generic broadcast: element-wise
    for (int64_t i = nvec * V + (int64_t)blockIdx.x * ST_THREADS + threadIdx.x; i < n; i += stride) {
        const float s = count == 1 ? s_one : DT<T>::to_f(scale[(i / inner) % count]);
        float t1, t3, t5;
        to_int_chain<T, RM>(DT<T>::to_f(x[i]), DivBy(s), p, t1, t3, t5);
        out[i] = (OUT)__float2int_rn(t5);
    }
}

// ------------------------------------------------------------------------------------------------------
// fused per-row abs-max + quant-dequant, TMA-staged (the C2 / C3 headline kernel)
// dynamic smem: [0,64) mbarriers | [64,192) reduction scratch | [256, ...) `stages` row buffers
// ------------------------------------------------------------------------------------------------------
constexpr int ROWS_MAX_STAGES = 8;
constexpr int ROWS_SMEM_HEADER = 256;

template <typename T>
__device__ __forceinline__ float finalize_scale(float amax, float min_val, int has_min, float int_thr, int scale_f32 = 0) {
    // _StatsScaling: ScalarClampMinSte(scaling_min_val) (core/restrict_val.py:22-42), then
    // RescalingIntQuant: scale = threshold / int_threshold (core/quant/int.py:160), both rounded to T
    // (0-dim T threshold / 0-dim fp32 int_threshold promotes to an fp32 scale: scale_f32)
    float thr = has_min ? clamp_min_nan(amax, min_val) : amax;
    float s = fdiv(thr, int_thr);
    return scale_f32 ? s : DT<T>::rnd(s);
}

template <typename T, int RM>
__global__ void rows_fwd_tma_kernel(const T* __restrict__ x, T* __restrict__ y, T* __restrict__ scale_out,
                                    T* __restrict__ absmax_out, int rows, int cols, int stages, uint32_t stage_stride,
                                    float min_val, int has_min, float int_thr, QParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint32_t* red = reinterpret_cast<uint32_t*>(smem + 64);
    unsigned char* bufs = smem + ROWS_SMEM_HEADER;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const uint32_t row_bytes = (uint32_t)cols * (uint32_t)sizeof(T);
    const int nvec = (int)(row_bytes >> 4);
    const int first = blockIdx.x, step = gridDim.x;
    const int my_rows = (rows - first + step - 1) / step;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    // Start-up: every CTA of the grid asks for its rows at the same moment.  If each one requested all its stages at once,
    // the FIRST rows -- the ones the whole chip is waiting for before it can store anything -- would share the read
    // bandwidth with everybody's prefetch (592 CTAs x 2 x 22 KB = 26 MB = 3.7 us of read-only time on C2 bf16).  So only
    // stage 0 is requested up front; the remaining stages are requested as soon as it has landed.
    const int pre = my_rows < stages ? my_rows : stages;
    if (tid == 0 && pre > 0) {
        mbar_arrive_expect_tx(&bars[0], row_bytes);
        bulk_g2s(bufs, x + (size_t)first * cols, row_bytes, &bars[0]);
    }

    int s = 0;
    uint32_t parity = 0;
    for (int it = 0; it < my_rows; ++it) {
        const int row = first + it * step;
        mbar_wait(&bars[s], parity);
        if (it == 0 && tid == 0) {
            for (int q = 1; q < pre; ++q) {
                mbar_arrive_expect_tx(&bars[q], row_bytes);
                bulk_g2s(bufs + (size_t)q * stage_stride, x + (size_t)(first + q * step) * cols, row_bytes, &bars[q]);
            }
        }
        const uint4* buf = reinterpret_cast<const uint4*>(bufs + (size_t)s * stage_stride);

        // pass 1 (shared memory): max |x| on raw bit patterns, 128-bit loads
        AbsMaxAcc<T> am;
#pragma unroll 4
        for (int v = tid; v < nvec; v += blockDim.x) am.add(lds128(buf + v));
        uint32_t m = warp_max_u32(am.result());
        if (lane == 0) red[warp] = m;
        __syncthreads();
        m = warp_max_u32(lane < nw ? red[lane] : 0u);

        const float amax = DT<T>::bits_to_f(m);
        const float sc = finalize_scale<T>(amax, min_val, has_min, int_thr);
        const ScaleCtx<T> cx(sc, true, p, PackedPath<T, RM>::value);
        if (tid == 0) {
            scale_out[row] = DT<T>::from_f(sc);
            if (absmax_out) absmax_out[row] = DT<T>::from_f(amax);
        }

        // pass 2 (shared memory -> HBM): quant-dequant, 128-bit stores
        uint4* yrow = reinterpret_cast<uint4*>(y + (size_t)row * cols);
        with_mode<T, RM, true>(cx.mode, [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll 2
            for (int v = tid; v < nvec; v += blockDim.x)
                stg_stream(yrow + v, qdq_vec<T, RM, MODE>(lds128(buf + v), cx, p));
        });
        __syncthreads();      // everyone is done with buf[s] and red[]
        if (tid == 0 && it + stages < my_rows) {
            mbar_arrive_expect_tx(&bars[s], row_bytes);
            bulk_g2s(bufs + (size_t)s * stage_stride, x + (size_t)(row + stages * step) * cols, row_bytes, &bars[s]);
        }
        if (++s == stages) { s = 0; parity ^= 1u; }
    }
}

// Same contract, output through the TMA engine as well: the row is quantized IN PLACE in its shared-memory stage and
// leaves with ONE bulk copy (cp.async.bulk.global.shared::cta) instead of a 128-bit store per thread and vector.  While
// row r is computed, row r-1 is still being read out of its stage; that stage is refilled (row r+stages-1) in the middle
// of row r, after the abs-max pass.  stages = 2 therefore prefetches only during the quantization pass of the row before,
// stages = 3 a whole row ahead.
template <typename T, int RM>
__global__ void rows_fwd_tma_store_kernel(const T* __restrict__ x, T* __restrict__ y, T* __restrict__ scale_out,
                                          T* __restrict__ absmax_out, int rows, int cols, int stages, uint32_t stage_stride,
                                          float min_val, int has_min, float int_thr, QParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint32_t* red = reinterpret_cast<uint32_t*>(smem + 64);
    unsigned char* bufs = smem + ROWS_SMEM_HEADER;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const uint32_t row_bytes = (uint32_t)cols * (uint32_t)sizeof(T);
    const int nvec = (int)(row_bytes >> 4);
    const int first = blockIdx.x, step = gridDim.x;
    const int my_rows = (rows - first + step - 1) / step;
    const int ahead = stages - 1;                     // loads in flight, the row being computed included

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int pre = my_rows < ahead ? my_rows : ahead;
    if (tid == 0 && pre > 0) {                        // staggered start-up, see rows_fwd_tma_kernel
        mbar_arrive_expect_tx(&bars[0], row_bytes);
        bulk_g2s(bufs, x + (size_t)first * cols, row_bytes, &bars[0]);
    }

    int s = 0;
    uint32_t parity = 0;
    for (int it = 0; it < my_rows; ++it) {
        const int row = first + it * step;
        mbar_wait(&bars[s], parity);
        if (it == 0 && tid == 0) {
            for (int q = 1; q < pre; ++q) {
                mbar_arrive_expect_tx(&bars[q], row_bytes);
                bulk_g2s(bufs + (size_t)q * stage_stride, x + (size_t)(first + q * step) * cols, row_bytes, &bars[q]);
            }
        }
        unsigned char* stage = bufs + (size_t)s * stage_stride;
        const uint4* buf = reinterpret_cast<const uint4*>(stage);

        AbsMaxAcc<T> am;
#pragma unroll 4
        for (int v = tid; v < nvec; v += blockDim.x) am.add(lds128(buf + v));
        uint32_t m = warp_max_u32(am.result());
        if (lane == 0) red[warp] = m;
        __syncthreads();
        m = warp_max_u32(lane < nw ? red[lane] : 0u);

        // refill the stage whose row left one iteration ago (its read-out has had the whole abs-max pass to finish;
        // the only bulk group that can still be pending is that one store)
        if (tid == 0 && it + ahead < my_rows) {
            bulk_wait_read<0>();
            const int sn = (s == 0) ? stages - 1 : s - 1;
            mbar_arrive_expect_tx(&bars[sn], row_bytes);
            bulk_g2s(bufs + (size_t)sn * stage_stride, x + (size_t)(first + (it + ahead) * step) * cols, row_bytes, &bars[sn]);
        }

        const float amax = DT<T>::bits_to_f(m);
        const float sc = finalize_scale<T>(amax, min_val, has_min, int_thr);
        const ScaleCtx<T> cx(sc, true, p, PackedPath<T, RM>::value);
        if (tid == 0) {
            scale_out[row] = DT<T>::from_f(sc);
            if (absmax_out) absmax_out[row] = DT<T>::from_f(amax);
        }

        with_mode<T, RM, true>(cx.mode, [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll 2
            for (int v = tid; v < nvec; v += blockDim.x)
                sts128(stage + (size_t)v * 16, qdq_vec<T, RM, MODE>(lds128(buf + v), cx, p));
        });
        fence_proxy_async_smem();                     // the generic-proxy writes above -> visible to the bulk copy
        __syncthreads();                              // everyone is done with stage s and red[]
        if (tid == 0) {
            bulk_s2g(y + (size_t)row * cols, stage, row_bytes);
            bulk_commit();
        }
        if (++s == stages) { s = 0; parity ^= 1u; }
    }
    if (tid == 0) bulk_wait_all<0>();                 // the stores must have left shared memory before the CTA exits
}

// generic rows forward: one CTA per row, element loads, any cols / alignment (second read hits L1/L2)
template <typename T, int RM>
__global__ void rows_fwd_generic_kernel(const T* __restrict__ x, T* __restrict__ y, T* __restrict__ scale_out,
                                        T* __restrict__ absmax_out, int64_t rows, int64_t cols,
                                        float min_val, int has_min, float int_thr, int quantize, QParams p) {
    __shared__ uint32_t red[32];
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const T* xr = x + row * cols;
        uint32_t m = 0;
        for (int64_t j = threadIdx.x; j < cols; j += blockDim.x) {
            uint32_t b = __float_as_uint(DT<T>::to_f(xr[j])) & 0x7fffffffu;
            m = max(m, b);
        }
        m = block_max_u32(m, red);
        const float amax = __uint_as_float(m);       // exact: came from a T value widened to fp32
        const float sc = finalize_scale<T>(amax, min_val, has_min, int_thr);
        if (threadIdx.x == 0) {
            if (scale_out) scale_out[row] = DT<T>::from_f(sc);
            if (absmax_out) absmax_out[row] = DT<T>::from_f(amax);
        }
        if (quantize) {
            T* yr = y + row * cols;
            const DivBy dv(sc, DT<T>::MUL_DIV_EXACT);
            for (int64_t j = threadIdx.x; j < cols; j += blockDim.x)
                yr[j] = DT<T>::from_f(quant_dequant<T, RM>(DT<T>::to_f(xr[j]), dv, p));
        }
    }
}

// per-row abs-max alone (AbsMax(dim) outside a fused quantizer: AbsMaxAve / AbsMaxL2, statistics over several
// tracked parameters): 128-bit loads, 4 in flight per thread, sign-split integer maxima; a warp per short row, the
// CTA per long row.  (The element-wise generic kernel above ran at 1.9 TB/s.)
constexpr int AMR_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(AMR_THREADS) absmax_rows_vec_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                                       int64_t rows, int64_t row_vecs, int group) {
    __shared__ uint32_t red[32];
    const int lane = threadIdx.x & 31;
    const int gid = (group == 32) ? (threadIdx.x >> 5) : 0;
    const int gtid = (group == 32) ? lane : threadIdx.x;
    const int groups_per_cta = AMR_THREADS / group;
    for (int64_t row = (int64_t)blockIdx.x * groups_per_cta + gid; row < rows; row += (int64_t)gridDim.x * groups_per_cta) {
        const uint4* xv = reinterpret_cast<const uint4*>(x) + row * row_vecs;
        AbsMaxAcc<T> am;
        int64_t v0 = gtid;
        for (; v0 + (int64_t)3 * group < row_vecs; v0 += (int64_t)group * 4) {      // full chunks: no predication
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = ldg_stream(xv + v0 + (int64_t)u * group);
#pragma unroll
            for (int u = 0; u < 4; ++u) am.add(q[u]);
        }
        for (; v0 < row_vecs; v0 += group) am.add(ldg_stream(xv + v0));
        uint32_t m = am.result();
        if (group == 32) {
            m = warp_max_u32(m);
        } else {
            m = block_max_u32(m, red);
        }
        if (gtid == 0) out[row] = DT<T>::from_f(DT<T>::bits_to_f(m));
    }
}

// ------------------------------------------------------------------------------------------------------
// per-row backward with gradient through the abs-max (streaming: 2R + 1W, then a one-element fix-up)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t canon_abs_bits(float v) {
    uint32_t b = __float_as_uint(v) & 0x7fffffffu;
    return b > 0x7f800000u ? 0x7f800001u : b;        // all NaNs tie, like torch.max
}

// Arg-max bookkeeping costs ~1.5 instructions per element: each thread keeps the max |x| bit pattern (in T's
// own bit domain) of the 16-byte vectors it visits and the index of the FIRST vector attaining it; the
// element inside the winning vector is located once per row by thread 0.
template <typename T, int RM, bool VECTOR>
__global__ void rows_bwd_kernel(const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ scale,
                                const T* __restrict__ gscale, T* __restrict__ gx, int64_t rows, int64_t cols,
                                float int_thr, int masked, QParams p) {
    constexpr int V = DT<T>::VEC;
    __shared__ uint32_t red_u[32];
    __shared__ float red_f[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const ScaleCtx<T> cx(DT<T>::to_f(scale[row]), true, p, false);       // generic path: literal formulation
        const DivBy& dv = cx.dv;
        const float inv_s = cx.inv_s;
        const T* gr = gy + row * cols;
        const T* xr = x + row * cols;
        T* outr = gx + row * cols;
        float acc = 0.f;
        uint32_t best = 0, best_pos = 0xffffffffu;
        if (VECTOR) {
            const int nvec = (int)(cols / V);
            const uint4* gv = reinterpret_cast<const uint4*>(gr);
            const uint4* xv = reinterpret_cast<const uint4*>(xr);
            uint4* ov = reinterpret_cast<uint4*>(outr);
            if (tid < nvec) best_pos = (uint32_t)tid;
            for (int v0 = tid; v0 < nvec; v0 += blockDim.x * 2) {
                const int v1 = v0 + blockDim.x;
                uint4 qg0 = ldg_stream(gv + v0), qx0 = ldg_stream(xv + v0);
                uint4 qg1 = make_uint4(0, 0, 0, 0), qx1 = make_uint4(0, 0, 0, 0);
                if (v1 < nvec) { qg1 = ldg_stream(gv + v1); qx1 = ldg_stream(xv + v1); }
                {
                    const uint32_t mv = DT<T>::absmax_fold(DT<T>::absmax_acc(0u, qx0));
                    if (mv > best) { best = mv; best_pos = (uint32_t)v0; }
                    stg_stream(ov + v0, bwd_vec<T, RM, VM_LITERAL>(qg0, qx0, cx, p, masked, true, acc));
                }
                if (v1 < nvec) {
                    const uint32_t mv = DT<T>::absmax_fold(DT<T>::absmax_acc(0u, qx1));
                    if (mv > best) { best = mv; best_pos = (uint32_t)v1; }
                    stg_stream(ov + v1, bwd_vec<T, RM, VM_LITERAL>(qg1, qx1, cx, p, masked, true, acc));
                }
            }
        } else {
            if (tid < cols) best_pos = (uint32_t)tid;
            for (int64_t j = tid; j < cols; j += blockDim.x) {
                const float xe = DT<T>::to_f(xr[j]);
                const uint32_t b = DT<T>::abs_bits_s(xe);
                if (b > best) { best = b; best_pos = (uint32_t)j; }
                outr[j] = DT<T>::from_f(bwd_elem<T, RM>(DT<T>::to_f(gr[j]), xe, dv, inv_s, p, masked, true, acc));
            }
        }
        // row reductions: max bits, sum, then the smallest position attaining the max
        uint32_t wm = warp_max_u32(best);
        float ws = warp_sum_f(acc);
        if (lane == 0) { red_u[warp] = wm; red_f[warp] = ws; }
        __syncthreads();
        const uint32_t rmax = warp_max_u32(lane < nw ? red_u[lane] : 0u);
        const float rsum = warp_sum_f(lane < nw ? red_f[lane] : 0.f);
        __syncthreads();
        uint32_t cand = (best == rmax) ? best_pos : 0xffffffffu;
        cand = warp_min_u32(cand);
        if (lane == 0) red_u[warp] = cand;
        __syncthreads();      // also orders this block's gx stores before the fix-up read below
        if (tid == 0) {
            uint32_t amin = 0xffffffffu;
            for (int w = 0; w < nw; ++w) amin = min(amin, red_u[w]);
            if (amin != 0xffffffffu) {     // cols > 0
                int64_t idx = amin;
                if (VECTOR) {              // first element of the winning vector that attains the row max
                    idx = (int64_t)amin * V;
                    for (int i = 0; i < V; ++i)
                        if (DT<T>::abs_bits_s(DT<T>::to_f(xr[idx + i])) == rmax) { idx += i; break; }
                }
                float gsc = rsum + (gscale ? DT<T>::to_f(gscale[row]) : 0.f);
                // scale = thr / int_thr  =>  d thr = d scale / int_thr; clamp_min_ste and view are identity;
                // max(dim) routes it to the arg-max, abs multiplies by sgn(x)
                float dthr = DT<T>::rnd(fdiv(DT<T>::rnd(gsc), int_thr));
                float xe = DT<T>::to_f(xr[idx]);
                float contrib = fmul(dthr, sign3(xe));
                float cur = DT<T>::to_f(outr[idx]);
                outr[idx] = DT<T>::from_f(fadd(cur, contrib));
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------
// per-row backward, TMA-pipelined: a producer warp streams tiles of the gradient row and the input row
// into a shared-memory ring with bulk copies (mbarrier full/empty pairs), consumer warps compute from
// shared memory and store gx with 128-bit streaming stores.  Loads are decoupled from the (heavy)
// arithmetic, so the HBM queue stays full regardless of register pressure / occupancy.
// dynamic smem: [0,64) full barriers | [64,128) empty barriers | [128,512) reduction scratch |
//               [1024, ...) stages x (gradient tile | input tile)
// ------------------------------------------------------------------------------------------------------
constexpr int BWD_MAX_STAGES = 8;
constexpr int BWD_SMEM_HEADER = 1024;

template <typename T, int RM>
__global__ void rows_bwd_tma_kernel(const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ scale,
                                    const T* __restrict__ gscale, T* __restrict__ gx, int rows, int cols,
                                    int tile_vecs, int stages, float int_thr, int masked, QParams p) {
    constexpr int V = DT<T>::VEC;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = reinterpret_cast<uint64_t*>(smem + 64);
    uint32_t* red_u = reinterpret_cast<uint32_t*>(smem + 128);      // 2 x 64 words
    float* red_f = reinterpret_cast<float*>(smem + 640);            // 2 x 32 floats
    unsigned char* ring = smem + BWD_SMEM_HEADER;

    const int tid = threadIdx.x;
    const int ncw = (blockDim.x >> 5) - 1;                // consumer warps (warp 0 is the producer)
    const int nct = ncw * 32;
    const uint32_t tile_bytes = (uint32_t)tile_vecs * 16u;
    const uint32_t row_bytes = (uint32_t)cols * (uint32_t)sizeof(T);
    const int row_vecs = (int)(row_bytes >> 4);
    const int tiles_per_row = (row_vecs + tile_vecs - 1) / tile_vecs;
    const int first = blockIdx.x, step = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (uint32_t)ncw); }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid < 32) {
        // ---------------- producer: one elected lane issues the bulk copies ----------------
        if (tid != 0) return;
        int it = 0;
        for (int row = first; row < rows; row += step) {
            const unsigned char* grow = reinterpret_cast<const unsigned char*>(gy + (size_t)row * cols);
            const unsigned char* xrow = reinterpret_cast<const unsigned char*>(x + (size_t)row * cols);
            for (int t = 0; t < tiles_per_row; ++t, ++it) {
                const int s = it % stages;
                const int use = it / stages;
                if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
                const uint32_t off = (uint32_t)t * tile_bytes;
                const uint32_t bytes = min(tile_bytes, row_bytes - off);
                unsigned char* gbuf = ring + (size_t)s * 2u * tile_bytes;
                mbar_arrive_expect_tx(&full[s], 2u * bytes);
                bulk_g2s(gbuf, grow + off, bytes, &full[s]);
                bulk_g2s(gbuf + tile_bytes, xrow + off, bytes, &full[s]);
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int ctid = tid - 32, lane = tid & 31, cw = (tid >> 5) - 1;
    int it = 0;
    int par = 0;                                               // reduction scratch is double-buffered by row parity
    // The one-element fix-up of a row (gradient through the abs-max) is software-pipelined ACROSS rows: the thread
    // that owns the arg-max vector re-loads that vector of x and of its own output at the end of row r and applies
    // the fix-up at the end of row r+1, so the loads' latency never stalls a consumer warp (and with it the ring).
    bool pend = false;
    uint4 pend_x = make_uint4(0, 0, 0, 0), pend_o = make_uint4(0, 0, 0, 0);
    uint32_t pend_rmax = 0;
    float pend_dthr = 0.f;
    T* pend_ptr = nullptr;
    auto apply_pending = [&]() {
        float fx[V], fo[V];
        DT<T>::unpack(pend_x, fx);
        DT<T>::unpack(pend_o, fo);
        int sel = 0;
        bool found = false;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const bool hit = !found && (DT<T>::abs_bits_s(fx[i]) == pend_rmax);
            if (hit) sel = i;
            found = found || hit;
        }
        float xe = fx[0], cur = fo[0];
#pragma unroll
        for (int i = 1; i < V; ++i) { if (sel == i) { xe = fx[i]; cur = fo[i]; } }
        pend_ptr[sel] = DT<T>::from_f(fadd(cur, fmul(pend_dthr, sign3(xe))));
    };
    float s_next = (first < rows) ? DT<T>::to_f(scale[first]) : 1.f;
    for (int row = first; row < rows; row += step, par ^= 1) {
        const ScaleCtx<T> cx(s_next, true, p, PackedPath<T, RM>::value);
        if (row + step < rows) s_next = DT<T>::to_f(scale[row + step]);     // prefetch: hides the load latency
        T* outr = gx + (size_t)row * cols;
        uint4* ov = reinterpret_cast<uint4*>(outr);
        const uint4* xv = reinterpret_cast<const uint4*>(x + (size_t)row * cols);
        float acc = 0.f;
        // arg-max bookkeeping: bit pattern of the largest |x| this thread has seen and the first vector attaining it
        uint32_t best = 0, best_pos = (ctid < row_vecs) ? (uint32_t)ctid : 0xffffffffu;
        with_mode<T, RM, false>(cx.mode, [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            for (int t = 0; t < tiles_per_row; ++t, ++it) {
                const int s = it % stages;
                mbar_wait(&full[s], (uint32_t)((it / stages) & 1));
                const int v_base = t * tile_vecs;
                const int nv = min(tile_vecs, row_vecs - v_base);
                const uint4* gbuf = reinterpret_cast<const uint4*>(ring + (size_t)s * 2u * tile_bytes);
                const uint4* xbuf = gbuf + tile_vecs;
                for (int v = ctid; v < nv; v += nct) {
                    const uint4 qg = lds128(gbuf + v);
                    const uint4 qx = lds128(xbuf + v);
                    const uint32_t mv = DT<T>::absmax_fold(DT<T>::absmax_acc(0u, qx));
                    stg_stream(ov + v_base + v, bwd_vec<T, RM, MODE>(qg, qx, cx, p, masked, true, acc));
                    if (mv > best) { best = mv; best_pos = (uint32_t)(v_base + v); }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);            // this warp no longer reads stage s
            }
        });
        // ---- row epilogue (ONE named barrier): per-warp (max bits, first position, sum) -> shared -> every warp
        const uint32_t wmax = warp_max_u32(best);
        const uint32_t wpos = warp_min_u32(best == wmax ? best_pos : 0xffffffffu);
        const float wsum = warp_sum_f(acc);
        uint32_t* ru = red_u + par * 64;
        float* rf = red_f + par * 32;
        if (lane == 0) { ru[cw] = wmax; ru[32 + cw] = wpos; rf[cw] = wsum; }
        named_bar_sync(1, nct);
        const uint32_t m_l = lane < ncw ? ru[lane] : 0u;
        const uint32_t p_l = lane < ncw ? ru[32 + lane] : 0xffffffffu;
        const uint32_t rmax = warp_max_u32(m_l);
        const uint32_t amin = warp_min_u32(m_l == rmax ? p_l : 0xffffffffu);
        const float rsum = warp_sum_f(lane < ncw ? rf[lane] : 0.f);
        if (pend) { apply_pending(); pend = false; }           // previous row's fix-up: its loads landed long ago
        if (best == rmax && best_pos == amin && amin != 0xffffffffu) {
            // exactly one thread: it owns the first vector attaining the row maximum
            const float gsc = rsum + (gscale ? DT<T>::to_f(gscale[row]) : 0.f);
            // scale = thr / int_thr  =>  d thr = d scale / int_thr; clamp_min_ste and the view are identity;
            // max(dim) routes it to the FIRST arg-max element, abs multiplies by sgn(x)
            pend_dthr = DT<T>::rnd(fdiv(DT<T>::rnd(gsc), int_thr));
            pend_rmax = rmax;
            pend_x = ldg_stream(xv + amin);
            pend_o = ldg_coherent(ov + amin);                 // this thread's own earlier store: same-thread RAW order
            pend_ptr = outr + (size_t)amin * V;
            pend = true;
        }
    }
    if (pend) apply_pending();
}

// ------------------------------------------------------------------------------------------------------
// provided-scale backward, TMA-pipelined (learned / constant / running-statistic scales: ParameterScaling,
// ParameterFromRuntimeStatsScaling, ConstScaling).  Same producer / consumer ring as rows_bwd_tma_kernel, minus the
// statistic: the tensor is a sequence of `nrows` chunks of `row_vecs` 16-byte vectors (the last one may be shorter);
// chunk r uses scale[r % count].  count == 1 (one scale for the tensor): d(scale) is accumulated over the whole CTA
// and added once; count > 1 (a scale per output channel / NCHW channel / token): one atomic per warp and chunk.
// ncu r01b: the register-staged streaming kernel it replaces ran at 24 % occupancy, loads not overlapped with the
// arithmetic (4.5 TB/s bf16, 5.3 TB/s fp32).
// dynamic smem: [0,64) full barriers | [64,128) empty barriers | [128,256) reduction scratch | [1024, ...) ring
// ------------------------------------------------------------------------------------------------------
template <typename T, int RM>
__global__ void scaled_bwd_tma_kernel(const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ scale,
                                      T* __restrict__ gx, float* gscale_out, long long n_vecs, int row_vecs,
                                      long long nrows, long long count, int tile_vecs, int stages, int scale_f32,
                                      int masked, QParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = reinterpret_cast<uint64_t*>(smem + 64);
    float* red_f = reinterpret_cast<float*>(smem + 128);
    unsigned char* ring = smem + BWD_SMEM_HEADER;

    const int tid = threadIdx.x;
    const int ncw = (blockDim.x >> 5) - 1;                // consumer warps (warp 0 is the producer)
    const int nct = ncw * 32;
    const uint32_t tile_bytes = (uint32_t)tile_vecs * 16u;
    const long long first = blockIdx.x, step = gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (uint32_t)ncw); }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid < 32) {
        if (tid != 0) return;
        const unsigned char* gb = reinterpret_cast<const unsigned char*>(gy);
        const unsigned char* xb = reinterpret_cast<const unsigned char*>(x);
        int it = 0;
        for (long long row = first; row < nrows; row += step) {
            const long long v0 = row * row_vecs;
            const int rv = (int)min((long long)row_vecs, n_vecs - v0);
            for (int t = 0; t * tile_vecs < rv; ++t, ++it) {
                const int s = it % stages;
                const int use = it / stages;
                if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
                const uint32_t bytes = (uint32_t)min(tile_vecs, rv - t * tile_vecs) * 16u;
                const size_t off = ((size_t)v0 + (size_t)t * tile_vecs) * 16u;
                unsigned char* gbuf = ring + (size_t)s * 2u * tile_bytes;
                mbar_arrive_expect_tx(&full[s], 2u * bytes);
                bulk_g2s(gbuf, gb + off, bytes, &full[s]);
                bulk_g2s(gbuf + tile_bytes, xb + off, bytes, &full[s]);
            }
        }
        return;
    }

    const int ctid = tid - 32, lane = tid & 31, cw = (tid >> 5) - 1;
    const bool want_gs = gscale_out != nullptr;
    const bool one_scale = count == 1;
    uint4* ov = reinterpret_cast<uint4*>(gx);
    int it = 0;
    float acc = 0.f;
    float s_next = 1.f;
    if (first < nrows) s_next = one_scale ? load_scale0<T>(scale, scale_f32) : DT<T>::to_f(scale[first % count]);
    for (long long row = first; row < nrows; row += step) {
        const ScaleCtx<T> cx(s_next, !scale_f32, p, PackedPath<T, RM>::value);
        if (!one_scale && row + step < nrows) s_next = DT<T>::to_f(scale[(row + step) % count]);   // prefetch
        const long long v0 = row * row_vecs;
        const int rv = (int)min((long long)row_vecs, n_vecs - v0);
        with_mode<T, RM, false>(cx.mode, [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
            for (int t = 0; t * tile_vecs < rv; ++t, ++it) {
                const int s = it % stages;
                mbar_wait(&full[s], (uint32_t)((it / stages) & 1));
                const int v_base = t * tile_vecs;
                const int nv = min(tile_vecs, rv - v_base);
                const uint4* gbuf = reinterpret_cast<const uint4*>(ring + (size_t)s * 2u * tile_bytes);
                const uint4* xbuf = gbuf + tile_vecs;
                for (int v = ctid; v < nv; v += nct) {
                    const uint4 qx = lds128(xbuf + v);
                    uint4 o = bwd_vec<T, RM, MODE>(lds128(gbuf + v), p.pre_relu ? relu_vec<T>(qx) : qx, cx, p, masked,
                                                   want_gs, acc);
                    stg_stream(ov + v0 + v_base + v, p.pre_relu ? relu_grad_vec<T>(o, qx) : o);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
        });
        if (want_gs && !one_scale) {
            const float wsum = warp_sum_f(acc);
            if (lane == 0) atomicAdd(gscale_out + (row % count), wsum);
            acc = 0.f;
        }
    }
    if (want_gs && one_scale) {
        const float wsum = warp_sum_f(acc);
        if (lane == 0) red_f[cw] = wsum;
        named_bar_sync(1, nct);
        if (cw == 0) {
            const float tsum = warp_sum_f(lane < ncw ? red_f[lane] : 0.f);
            if (lane == 0) atomicAdd(gscale_out, tsum);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// whole-tensor abs-max: phase 1 block maxima -> atomicMax, last block finalises the scale
// workspace words: [0] max bits  [1] ticket  [2] Gs (float)  [3] tie count  [4..] tie indices (int64)
// ------------------------------------------------------------------------------------------------------
constexpr int WS_MAXBITS = 0, WS_TICKET = 1, WS_GS = 2, WS_TIES = 3, WS_LIST = 4;
constexpr int TIE_CAP = 4096;

template <typename T>
__global__ void __launch_bounds__(ST_THREADS) absmax_tensor_kernel(
        const T* __restrict__ x, int64_t n, int vec_ok, uint32_t* ws, T* scale_out, T* absmax_out,
        float min_val, int has_min, float int_thr, int scale_f32) {
    constexpr int V = DT<T>::VEC;
    __shared__ uint32_t red[32];
    const int64_t nvec = vec_ok ? n / V : 0;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    // every CTA reduces ONE contiguous, equally sized slice (a grid-stride loop over 16 KB chunks left some CTAs with
    // 5 chunks and others with 4: a 20 % tail on a read-only kernel), 8 independent 16-byte loads in flight per thread
    const int64_t per_cta = (nvec + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per_cta;
    const int64_t hi = lo + per_cta < nvec ? lo + per_cta : nvec;
    AbsMaxAcc<T> am;
    int64_t b = lo + threadIdx.x;
    for (; b + (int64_t)7 * ST_THREADS < hi; b += (int64_t)ST_THREADS * 8) {      // full chunks: no predication
        uint4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = ldg_stream(xv + b + (int64_t)u * ST_THREADS);
#pragma unroll
        for (int u = 0; u < 8; ++u) am.add(q[u]);
    }
    for (; b < hi; b += ST_THREADS) am.add(ldg_stream(xv + b));
    uint32_t m = am.result();
    // widen to fp32 bit ordering so that all dtypes share the same workspace encoding
    uint32_t mf = __float_as_uint(DT<T>::bits_to_f(m));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = nvec * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        mf = max(mf, __float_as_uint(DT<T>::to_f(x[i])) & 0x7fffffffu);
    mf = block_max_u32(mf, red);
    if (threadIdx.x == 0) {
        atomicMax(ws + WS_MAXBITS, mf);
        __threadfence();
        uint32_t ticket = atomicAdd(ws + WS_TICKET, 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            uint32_t all = atomicMax(ws + WS_MAXBITS, 0u);
            float amax = __uint_as_float(all);
            if (absmax_out) absmax_out[0] = DT<T>::from_f(amax);
            if (scale_out) {
                float sc = finalize_scale<T>(amax, min_val, has_min, int_thr, scale_f32);
                if (scale_f32) reinterpret_cast<float*>(scale_out)[0] = sc;
                else scale_out[0] = DT<T>::from_f(sc);
            }
        }
    }
}

// provided-scale backward for the whole-tensor statistic: additionally collects the positions of the tied maxima
template <typename T, int RM>
__global__ void __launch_bounds__(ST_THREADS) tensor_bwd_kernel(
        const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ scale, const T* __restrict__ absmax,
        T* __restrict__ gx, int64_t n, int vec_ok, uint32_t* ws, int scale_f32, int masked, QParams p) {
    constexpr int V = DT<T>::VEC;
    __shared__ float red[32];
    const ScaleCtx<T> cx(load_scale0<T>(scale, scale_f32), !scale_f32, p, PackedPath<T, RM>::value);
    const DivBy& dv = cx.dv;
    const float inv_s = cx.inv_s;
    const uint32_t mbits = canon_abs_bits(DT<T>::to_f(absmax[0]));
    long long* list = reinterpret_cast<long long*>(ws + WS_LIST);
    float acc = 0.f;
    const int64_t nvec = vec_ok ? n / V : 0;
    const uint4* gv = reinterpret_cast<const uint4*>(gy);
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    uint4* ov = reinterpret_cast<uint4*>(gx);
    const int64_t chunk = ST_THREADS * ST_UNROLL;
    const int64_t nchunks = (nvec + chunk - 1) / chunk;
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t base = c * chunk + threadIdx.x;
        uint4 qg[ST_UNROLL], qx[ST_UNROLL];
#pragma unroll
        for (int u = 0; u < ST_UNROLL; ++u) {
            int64_t v = base + (int64_t)u * ST_THREADS;
            if (v < nvec) { qg[u] = ldg_stream(gv + v); qx[u] = ldg_stream(xv + v); }
        }
#pragma unroll
        for (int u = 0; u < ST_UNROLL; ++u) {
            int64_t v = base + (int64_t)u * ST_THREADS;
            if (v < nvec) {
                float eg[V], ex[V];
                DT<T>::unpack(qg[u], eg);
                DT<T>::unpack(qx[u], ex);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    if (canon_abs_bits(ex[i]) == mbits) {
                        uint32_t slot = atomicAdd(ws + WS_TIES, 1u);
                        if (slot < TIE_CAP) list[slot] = v * V + i;
                    }
                }
                if (cx.mode == VM_LITERAL) stg_stream(ov + v, bwd_vec<T, RM, VM_LITERAL>(qg[u], qx[u], cx, p, masked, true, acc));
                else stg_stream(ov + v, bwd_vec<T, RM, VM_PACKED>(qg[u], qx[u], cx, p, masked, true, acc));
            }
        }
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = nvec * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float xe = DT<T>::to_f(x[i]);
        if (canon_abs_bits(xe) == mbits) {
            uint32_t slot = atomicAdd(ws + WS_TIES, 1u);
            if (slot < TIE_CAP) list[slot] = i;
        }
        gx[i] = DT<T>::from_f(bwd_elem<T, RM>(DT<T>::to_f(gy[i]), xe, dv, inv_s, p, masked, true, acc));
    }
    float t = block_sum_f(acc, red);
    if (threadIdx.x == 0) atomicAdd(reinterpret_cast<float*>(ws + WS_GS), t);
}

// torch.max() backward: grad / (#ties) to every tied maximum (evenly_distribute_backward), times sgn(x)
template <typename T>
__global__ void tensor_bwd_fixup_kernel(const T* __restrict__ x, const T* __restrict__ absmax, const T* __restrict__ gscale,
                                        T* __restrict__ gx, int64_t n, const uint32_t* ws, float int_thr, int scale_f32) {
    const uint32_t ties = ws[WS_TIES];
    if (ties == 0) return;
    const float gsc = *reinterpret_cast<const float*>(ws + WS_GS) + (gscale ? load_scale0<T>(gscale, scale_f32) : 0.f);
    const float dthr = DT<T>::rnd(fdiv(DT<T>::rnd(gsc), int_thr));
    const float share = DT<T>::rnd(fdiv(dthr, (float)ties));
    const long long* list = reinterpret_cast<const long long*>(ws + WS_LIST);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (ties <= TIE_CAP) {
        for (int64_t k = tid; k < ties; k += stride) {
            int64_t i = list[k];
            float xe = DT<T>::to_f(x[i]);
            gx[i] = DT<T>::from_f(fadd(DT<T>::to_f(gx[i]), fmul(share, sign3(xe))));
        }
    } else {       // pathological tie count: rescan
        const uint32_t mbits = canon_abs_bits(DT<T>::to_f(absmax[0]));
        for (int64_t i = tid; i < n; i += stride) {
            float xe = DT<T>::to_f(x[i]);
            if (canon_abs_bits(xe) == mbits)
                gx[i] = DT<T>::from_f(fadd(DT<T>::to_f(gx[i]), fmul(share, sign3(xe))));
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host-side launch logic
// ------------------------------------------------------------------------------------------------------
// ---- host-side helpers for the packed low-precision path: walk T's value grid through 16-bit patterns ----
static inline uint16_t t_bits(float v, int dtype) {
    if (dtype == BVB_BF16) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
    __half h = __float2half_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
}
static inline float t_value(uint16_t b, int dtype) {
    if (dtype == BVB_BF16) { __nv_bfloat16 h; *reinterpret_cast<uint16_t*>(&h) = b; return __bfloat162float(h); }
    __half h;
    *reinterpret_cast<uint16_t*>(&h) = b;
    return __half2float(h);
}
static inline uint16_t t_step(uint16_t b, int dir) {        // next representable value above (dir > 0) / below
    if ((b & 0x7fffu) == 0) return dir > 0 ? (uint16_t)0x0001u : (uint16_t)0x8001u;
    const bool neg = (b & 0x8000u) != 0;
    return (uint16_t)((neg == (dir < 0)) ? b + 1 : b - 1);
}

QParams make_qparams(float zero_point, float qmin, float qmax, int dtype) {
    QParams p;
    p.qmin = round_to_dtype(qmin, dtype);
    p.qmax = round_to_dtype(qmax, dtype);
    p.zp = round_to_dtype(zero_point, dtype);
    p.zp_nonzero = (zero_point != 0.f) ? 1 : 0;
    p.pre_relu = 0;
    p.pk_ok = p.pk_lo_zero = p.pk_lo = p.pk_hi = p.pk_lo_pre = p.pk_thr_lo = p.pk_thr_hi = 0;
    if (dtype == BVB_F32 || zero_point != 0.f) return p;
    // magic-number rounding range: fp16 adds 1536 in fp16 ([-512, 511] keeps the sum in the ulp-1 binade), bf16 adds
    // 1.5*2^23 in fp32; the bounds must be the integers the caller asked for (a 16-bit range is not representable)
    const float lim = dtype == BVB_F16 ? 511.f : 2097152.f;
    const float lo = p.qmin, hi = p.qmax, lo_pre = (lo == 0.f) ? -1.f : lo;
    if (!(lo == qmin && hi == qmax && lo == rintf(lo) && hi == rintf(hi) && lo_pre >= -lim - 1.f && hi <= lim && lo <= hi))
        return p;
    // thr_hi: the largest T value v with round(v) <= hi;  thr_lo: the smallest with round(v) >= lo  (round-half-even,
    // monotone), found by walking T's grid from a nearby starting point
    uint16_t bh = t_bits(hi + 1.f, dtype), bl = t_bits(lo - 1.f, dtype);
    int guard = 0;
    while (rintf(t_value(bh, dtype)) > hi && ++guard < 4096) bh = t_step(bh, -1);
    while (rintf(t_value(t_step(bh, +1), dtype)) <= hi && ++guard < 4096) bh = t_step(bh, +1);
    while (rintf(t_value(bl, dtype)) < lo && ++guard < 4096) bl = t_step(bl, +1);
    while (rintf(t_value(t_step(bl, -1), dtype)) >= lo && ++guard < 4096) bl = t_step(bl, -1);
    if (guard >= 4096) return p;
    auto dup = [&](float v) { const uint32_t b = t_bits(v, dtype); return b | (b << 16); };
    p.pk_lo = dup(lo);
    p.pk_hi = dup(hi);
    p.pk_lo_pre = dup(lo_pre);
    p.pk_lo_zero = (lo == 0.f) ? 1u : 0u;
    p.pk_thr_hi = (uint32_t)bh | ((uint32_t)bh << 16);
    p.pk_thr_lo = (uint32_t)bl | ((uint32_t)bl << 16);
    p.pk_ok = 1;
    return p;
}

static inline unsigned stream_grid(int64_t nvec_or_n, int per_block) {
    int64_t b = (nvec_or_n + per_block - 1) / per_block;
    int per_sm = tuning().stream_ctas_per_sm > 0 ? tuning().stream_ctas_per_sm : 16;
    int64_t cap = (int64_t)sm_count() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// Read-only statistic kernels (abs-max, radix-select histograms).  tools/readbench.cu: a bare read pass over 180 MB
// takes 30.8 us with exactly 1024 resident threads per SM and 34.7-35.0 us with 512 or 2048, whatever the unroll depth
// (profiles/sweeps/r01w_readbench.log); with the statistic's own arithmetic on top the optimum moves to 5-6 CTAs of
// 256 threads per SM (tools/statbench.py, profiles/sweeps/r01w_statbench.log).  One wave, contiguous work per CTA.
unsigned stat_grid(int64_t blocks_wanted, int default_per_sm) {
    const int per_sm = tuning().stream_ctas_per_sm > 0 ? tuning().stream_ctas_per_sm : default_per_sm;
    int64_t cap = (int64_t)sm_count() * per_sm;
    int64_t b = blocks_wanted < cap ? blocks_wanted : cap;
    return (unsigned)(b < 1 ? 1 : b);
}

// channels-last per-channel kernels: ONE resident wave (every CTA pays V divisor set-ups and, in the backward, C atomics;
// tools/clbench.py: 63.5 us with the 3 CTAs/SM the backward's registers allow, 78 us with a grid of 4/SM = 1.33 waves)
template <typename K>
static unsigned chanlast_grid(K kernel, int64_t nvec) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, CL_THREADS, 0) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 2;
    }
    if (occ > 4) occ = 4;                                   // 1024 threads per SM
    if (tuning().stream_ctas_per_sm > 0) occ = tuning().stream_ctas_per_sm;
    int64_t b = (nvec + CL_THREADS - 1) / CL_THREADS;
    const int64_t cap = (int64_t)sm_count() * occ;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

template <typename T, int RM>
static int launch_int_quant_fwd(const void* x, const void* scale, void* y, void* codes, int64_t n,
                                int64_t inner, int64_t count, int scale_f32, const QParams& p, int reverse,
                                cudaStream_t st) {
    constexpr int V = DT<T>::VEC;
    int64_t nvec = 0;
    int smode = (count == 1) ? 0 : 1;
    bool vec_ok = aligned16(x) && aligned16(y) && (!codes || aligned16(codes));
    if (smode == 1 && inner > 1) {          // planes of `inner` elements sharing one scale
        const int64_t nplanes = n / inner;
        const int group = inner * (int64_t)sizeof(T) <= 4096 ? 32 : PL_THREADS;
        const int pv = (vec_ok && (inner % V) == 0) ? 1 : 0;
        int64_t grid = (nplanes + (PL_THREADS / group) - 1) / (PL_THREADS / group);
        const int64_t cap = (int64_t)sm_count() * 8;
        if (grid > cap) grid = cap;
        int_quant_planes_kernel<T, RM, false><<<(unsigned)grid, PL_THREADS, 0, st>>>(
            nullptr, (const T*)x, (const T*)scale, (T*)y, (T*)codes, nullptr, nplanes, inner, count, group, pv, 0, p);
        return check_launch("bvb_int_quant_fwd");
    }
    if (smode == 1 && inner == 1 && vec_ok && aligned16(scale) && count <= CL_MAX_C && (count % V) == 0 &&
        ((int64_t)CL_THREADS * V) % count == 0 && (n % count) == 0 && n >= (int64_t)CL_THREADS * V) {
        // channels-last tensor, one scale per channel
        const int64_t nv = n / V;
        if constexpr (PackedPath<T, RM>::value && DT<T>::MUL_DIV_EXACT) {
            if (p.pk_ok) {
                const unsigned grid = chanlast_grid(int_quant_chanlast_kernel<T, RM, false, true>, nv);
                int_quant_chanlast_kernel<T, RM, false, true><<<grid, CL_THREADS, 0, st>>>(
                    nullptr, (const T*)x, (const T*)scale, (T*)y, (T*)codes, nullptr, nv, (int)count, 0, p);
                return check_launch("bvb_int_quant_fwd");
            }
        }
        const unsigned grid = chanlast_grid(int_quant_chanlast_kernel<T, RM, false, false>, nv);
        int_quant_chanlast_kernel<T, RM, false, false><<<grid, CL_THREADS, 0, st>>>(
            nullptr, (const T*)x, (const T*)scale, (T*)y, (T*)codes, nullptr, nv, (int)count, 0, p);
        return check_launch("bvb_int_quant_fwd");
    }
    if (smode == 1 && (inner % V) != 0) vec_ok = false;
    if (vec_ok) nvec = n / V;
    if (nvec > 0) {
        unsigned grid = stream_grid(nvec, ST_THREADS * ST_UNROLL);
        int_quant_fwd_kernel<T, RM><<<grid, ST_THREADS, 0, st>>>((const T*)x, (const T*)scale, (T*)y, (T*)codes, nvec,
                                                                 smode ? inner / V : 1, count, smode, scale_f32, reverse, p);
    }
    int64_t done = nvec * V;
    if (done < n) {
        unsigned grid = stream_grid(n - done, 256);
        int_quant_fwd_scalar_kernel<T, RM><<<grid, 256, 0, st>>>((const T*)x, (const T*)scale, (T*)y, (T*)codes,
                                                                 done, n, inner, count, scale_f32, p);
    }
    return check_launch("bvb_int_quant_fwd");
}

struct BwdGeom { int threads, tile_vecs, stages, ctas_per_sm; size_t smem; bool ok; };

// persistent grids must fit in ONE resident wave: a grid of 4 CTAs/SM of which only 3 are resident runs a second,
// quarter-full wave (sweeps: 100.9 us against 90.5 us for the same kernel)
template <typename K>
static int resident_ctas(K kernel, int threads, size_t smem, int wanted) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        return wanted;
    }
    return wanted < occ ? wanted : occ;
}
static BwdGeom bwd_geometry(int64_t cols, int elem_size, bool provided_scale = false);

// TMA-pipelined provided-scale backward over n_vecs 16-byte vectors cut into chunks of row_vecs
template <typename T, int RM>
static int launch_scaled_bwd_tma(const void* gy, const void* x, const void* scale, void* gx, float* gscale_out,
                                 int64_t n_vecs, int64_t row_vecs, int64_t count, int scale_f32, const QParams& p,
                                 int masked, cudaStream_t st, bool* launched) {
    *launched = false;
    BwdGeom g = bwd_geometry(row_vecs * 16 / (int64_t)sizeof(T), (int)sizeof(T), true);
    if (!g.ok || row_vecs >= ((int64_t)1 << 27)) return BVB_OK;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(scaled_bwd_tma_kernel<T, RM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) return fail(BVB_ECUDA, "scaled_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int64_t nrows = (n_vecs + row_vecs - 1) / row_vecs;
    int64_t grid = (int64_t)resident_ctas(scaled_bwd_tma_kernel<T, RM>, g.threads, g.smem, g.ctas_per_sm) * sm_count();
    if (grid > nrows) grid = nrows;
    scaled_bwd_tma_kernel<T, RM><<<(unsigned)grid, g.threads, g.smem, st>>>(
        (const T*)gy, (const T*)x, (const T*)scale, (T*)gx, gscale_out, (long long)n_vecs, (int)row_vecs,
        (long long)nrows, (long long)count, g.tile_vecs, g.stages, scale_f32, masked, p);
    *launched = true;
    return check_launch("bvb_int_quant_bwd");
}

template <typename T, int RM>
static int launch_int_quant_bwd(const void* gy, const void* x, const void* scale, void* gx, float* gscale_out,
                                int64_t n, int64_t inner, int64_t count, int scale_f32, const QParams& p, int masked,
                                cudaStream_t st) {
    constexpr int V = DT<T>::VEC;
    if (gscale_out) {
        cudaError_t e = cudaMemsetAsync(gscale_out, 0, sizeof(float) * (size_t)count, st);
        if (e != cudaSuccess) return fail(BVB_ECUDA, "bvb_int_quant_bwd: memset: %s", cudaGetErrorString(e));
    }
    int64_t nvec = 0;
    int smode = (count == 1) ? 0 : 1;
    bool vec_ok = aligned16(gy) && aligned16(x) && aligned16(gx);
    if (smode == 1 && inner > 1) {
        const int64_t nplanes = n / inner;
        const int group = inner * (int64_t)sizeof(T) <= 4096 ? 32 : PL_THREADS;
        const int pv = (vec_ok && (inner % V) == 0) ? 1 : 0;
        if (pv && inner * (int64_t)sizeof(T) >= 4096 && n * (int64_t)sizeof(T) >= (1 << 20)) {
            bool launched = false;            // large planes: the TMA producer / consumer ring
            int rc = launch_scaled_bwd_tma<T, RM>(gy, x, scale, gx, gscale_out, n / V, inner / V, count, 0, p, masked, st,
                                                  &launched);
            if (rc != BVB_OK || launched) return rc;
        }
        int64_t grid = (nplanes + (PL_THREADS / group) - 1) / (PL_THREADS / group);
        const int64_t cap = (int64_t)sm_count() * 8;
        if (grid > cap) grid = cap;
        int_quant_planes_kernel<T, RM, true><<<(unsigned)grid, PL_THREADS, 0, st>>>(
            (const T*)gy, (const T*)x, (const T*)scale, (T*)gx, nullptr, gscale_out, nplanes, inner, count, group, pv,
            masked, p);
        return check_launch("bvb_int_quant_bwd");
    }
    if (smode == 1 && inner == 1 && vec_ok && aligned16(scale) && count <= CL_MAX_C && (count % V) == 0 &&
        ((int64_t)CL_THREADS * V) % count == 0 && (n % count) == 0 && n >= (int64_t)CL_THREADS * V) {
        const int64_t nv = n / V;
        if constexpr (PackedPath<T, RM>::value && DT<T>::MUL_DIV_EXACT) {
            if (p.pk_ok) {
                const unsigned grid = chanlast_grid(int_quant_chanlast_kernel<T, RM, true, true>, nv);
                int_quant_chanlast_kernel<T, RM, true, true><<<grid, CL_THREADS, 0, st>>>(
                    (const T*)gy, (const T*)x, (const T*)scale, (T*)gx, nullptr, gscale_out, nv, (int)count, masked, p);
                return check_launch("bvb_int_quant_bwd");
            }
        }
        const unsigned grid = chanlast_grid(int_quant_chanlast_kernel<T, RM, true, false>, nv);
        int_quant_chanlast_kernel<T, RM, true, false><<<grid, CL_THREADS, 0, st>>>(
            (const T*)gy, (const T*)x, (const T*)scale, (T*)gx, nullptr, gscale_out, nv, (int)count, masked, p);
        return check_launch("bvb_int_quant_bwd");
    }
    if (smode == 1 && (inner % V) != 0) vec_ok = false;
    if (vec_ok) nvec = n / V;
    bool tma_done = false;
    if (smode == 0 && nvec >= (1 << 16)) {    // one scale, >= 1 MiB: chunks of 64 KiB through the TMA ring
        int rc = launch_scaled_bwd_tma<T, RM>(gy, x, scale, gx, gscale_out, nvec, 4096, 1, scale_f32, p, masked, st,
                                              &tma_done);
        if (rc != BVB_OK) return rc;
    }
    if (nvec > 0 && !tma_done) {
        unsigned grid = stream_grid(nvec, ST_THREADS * ST_UNROLL);
        int_quant_bwd_kernel<T, RM><<<grid, ST_THREADS, 0, st>>>((const T*)gy, (const T*)x, (const T*)scale, (T*)gx,
                                                                 gscale_out, nvec, smode ? inner / V : 1, count, smode,
                                                                 scale_f32, masked, p);
    }
    int64_t done = nvec * V;
    if (done < n) {
        unsigned grid = stream_grid(n - done, 256);
        int_quant_bwd_scalar_kernel<T, RM><<<grid, 256, 0, st>>>((const T*)gy, (const T*)x, (const T*)scale, (T*)gx,
                                                                 gscale_out, done, n, inner, count, scale_f32, masked, p);
    }
    return check_launch("bvb_int_quant_bwd");
}

// geometry of the TMA rows kernel
struct RowsGeom { int threads, stages, ctas_per_sm; uint32_t stage_stride; size_t smem; bool ok; bool tma_store; };

static RowsGeom rows_geometry(int64_t cols, int elem_size) {
    RowsGeom g = {0, 0, 0, 0, 0, false, false};
    const int64_t row_bytes = cols * elem_size;
    if (row_bytes < 16 || (row_bytes & 15) != 0) return g;
    g.stage_stride = (uint32_t)((row_bytes + 127) & ~(int64_t)127);
    const int64_t stride = g.stage_stride;
    if (2 * stride + ROWS_SMEM_HEADER > 226 * 1024) return g;       // a row pair must fit in one SM's shared memory
    // Geometry from tools/kbench.py sweeps on a B200 (profiles/r01f_sweeps.md; 45 M elements, row lengths 2 Ki .. 14 Ki
    // elements).  Both dtypes want FEW threads per SM -- the arithmetic is cheap next to the HBM time, and more
    // concurrently storing warps made every shape slower -- and a bounded amount of prefetch:
    //   fp32          ~512 threads and ~64-130 KB of row buffers per SM: 4 CTAs x 128 threads for 8 KB rows,
    //                 2 x 256 for 16 KB, 1 x 512 with 3 stages for 32 KB, 1 x 512 with 2 stages beyond
    //   bf16 / fp16   (heavier per byte) 128-thread CTAs, 2 stages, as many CTAs as fit up to 6
    int threads, stages = 2, ctas;
    if (elem_size == 4) {
        ctas = (int)((32 * 1024 + stride / 2) / stride);
        if (ctas < 1) ctas = 1;
        if (ctas > 8) ctas = 8;
        threads = 512 / ctas;
        if (threads < 64) threads = 64;
        if (ctas == 1 && 3 * stride <= 100 * 1024) stages = 3;
    } else {
        // r02 sweep with 32 / 64 / 128-thread CTAs (profiles/r02_fwd_bf16_small_cta_sweep.md): two warps per row beat four
        // by 3 % on both north-star shapes (C2 bf16 35.3 -> 34.1 us, C3 bf16 48.3 -> 46.8 us); one warp per row (a
        // quarter of the per-row instructions) gains nothing more -- these kernels are not issue-bound
        threads = row_bytes >= 4096 ? 64 : 128;
        ctas = (int)((200 * 1024) / (2 * stride));
        if (ctas > 6) ctas = 6;
        if (ctas < 1) { ctas = 1; threads = 256; }
        // long 16-bit rows (C2: 22 KB): quantize in place and let the TMA engine write the row back
        // (rows_fwd_tma_store_kernel; 3 stages x 3 CTAs x 128 threads).  r02 sweep, profiles/r02_fwd_tma_store_sweep.md:
        // C2 bf16 33.8 -> 32.9 us, fp16 34.6 -> 33.0 us; 8 KB rows (C3) and fp32 rows gain nothing (<= 1 %) and keep
        // the per-thread stores
        if (row_bytes >= 16384 && 3 * (ROWS_SMEM_HEADER + 3 * stride + 1024) <= 227 * 1024) {
            g.tma_store = true;
            threads = 128; stages = 3; ctas = 3;
        }
    }
    const Tuning& t = tuning();
    if (t.rows_threads > 0) threads = t.rows_threads;
    if (t.rows_stages >= 100) { g.tma_store = true; stages = t.rows_stages - 100; }     // sweep build: 100 + stages
    else if (t.rows_stages > 0) { g.tma_store = false; stages = t.rows_stages; }
    if (t.rows_ctas_per_sm > 0) ctas = t.rows_ctas_per_sm;
    if (stages > ROWS_MAX_STAGES) stages = ROWS_MAX_STAGES;
    int max_ctas = 2048 / threads;
    if (max_ctas > 16) max_ctas = 16;
    if (ctas > max_ctas) ctas = max_ctas;
    // re-validate against the shared-memory budget
    while (ctas > 1 && (int64_t)ctas * (ROWS_SMEM_HEADER + (int64_t)stages * stride + 1024) > 227 * 1024) --ctas;
    while (stages > 1 && (ROWS_SMEM_HEADER + (int64_t)stages * stride) > 226 * 1024) --stages;
    if (stages < 2) return g;
    g.threads = threads;
    g.stages = stages;
    g.ctas_per_sm = ctas;
    g.smem = ROWS_SMEM_HEADER + (size_t)stages * g.stage_stride;
    g.ok = true;
    return g;
}

template <typename T, int RM>
static int launch_rows_fwd(const void* x, void* y, void* scale_out, void* absmax_out, int64_t rows, int64_t cols,
                           float min_val, int has_min, float int_thr, const QParams& p, cudaStream_t st) {
    RowsGeom g = rows_geometry(cols, (int)sizeof(T));
    const bool tma_ok = g.ok && aligned16(x) && aligned16(y) && rows < (int64_t)1 << 30 && cols < (int64_t)1 << 30;
    if (tma_ok) {
        static bool attr_set[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(rows_fwd_tma_kernel<T, RM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 227 * 1024);
            if (e != cudaSuccess) return fail(BVB_ECUDA, "rows_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        if constexpr (sizeof(T) == 2) if (g.tma_store) {      // 16-bit rows only (fp32 gains nothing: not instantiated)
            static bool attr_set_store[64] = {false};
            if (dev < 0 || dev >= 64 || !attr_set_store[dev]) {
                cudaError_t e = cudaFuncSetAttribute(rows_fwd_tma_store_kernel<T, RM>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (e != cudaSuccess) return fail(BVB_ECUDA, "rows_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                if (dev >= 0 && dev < 64) attr_set_store[dev] = true;
            }
            int64_t grid = (int64_t)resident_ctas(rows_fwd_tma_store_kernel<T, RM>, g.threads, g.smem, g.ctas_per_sm) * sm_count();
            if (grid > rows) grid = rows;
            rows_fwd_tma_store_kernel<T, RM><<<(unsigned)grid, g.threads, g.smem, st>>>(
                (const T*)x, (T*)y, (T*)scale_out, (T*)absmax_out, (int)rows, (int)cols, g.stages, g.stage_stride,
                min_val, has_min, int_thr, p);
            return check_launch("bvb_rows_absmax_int_quant_fwd");
        }
        int64_t grid = (int64_t)resident_ctas(rows_fwd_tma_kernel<T, RM>, g.threads, g.smem, g.ctas_per_sm) * sm_count();
        if (grid > rows) grid = rows;
        rows_fwd_tma_kernel<T, RM><<<(unsigned)grid, g.threads, g.smem, st>>>(
            (const T*)x, (T*)y, (T*)scale_out, (T*)absmax_out, (int)rows, (int)cols, g.stages, g.stage_stride,
            min_val, has_min, int_thr, p);
    } else {
        int threads = cols >= 4096 ? 512 : (cols >= 512 ? 256 : 64);
        int64_t grid = rows;
        int64_t cap = (int64_t)sm_count() * (2048 / threads);
        if (grid > cap) grid = cap;
        rows_fwd_generic_kernel<T, RM><<<(unsigned)grid, threads, 0, st>>>(
            (const T*)x, (T*)y, (T*)scale_out, (T*)absmax_out, rows, cols, min_val, has_min, int_thr, 1, p);
    }
    return check_launch("bvb_rows_absmax_int_quant_fwd");
}

// geometry of the TMA backward: consumer warps, tile size, ring depth, CTAs per SM

static BwdGeom bwd_geometry(int64_t cols, int elem_size, bool provided_scale) {
    BwdGeom g = {0, 0, 0, 0, 0, false};
    const int64_t row_bytes = cols * elem_size;
    if (row_bytes < 16 || (row_bytes & 15) != 0 || row_bytes >= ((int64_t)1 << 31)) return g;
    const int64_t row_vecs = row_bytes / 16;
    // defaults from tools/kbench.py sweeps on a B200 (C2 4096x11008 fp32/bf16, C3 16384x4096 bf16): 8 KB tiles,
    // 4 consumer warps, a 2-deep ring per CTA and 4 CTAs per SM won or tied every sweep; bigger tiles lose.
    // re-swept after the packed-arithmetic rewrite (profiles/r01f_sweeps.md): like the forward, ~500-770 threads per SM
    // win -- fp32: 3 consumer warps x 4 CTAs (85.5 vs 90.4 us on C2); 16-bit long rows: 4 x 4; 16-bit rows below
    // 16 KB (C3): 2 consumer warps x 8 CTAs with 4 KB tiles (67.9 vs 74.4 us)
    // The provided-scale kernel (masked clamp + d(scale), the literal fp32 chain is its heaviest case) wants 8 consumer
    // warps and exactly one resident wave: 3 CTAs/SM fp32 (90.5 vs 102 us), 2 CTAs/SM 16-bit (70.2 vs 72.3 us).
    int ncw, ctas_default;
    if (provided_scale) {
        ncw = row_vecs >= 1024 ? 8 : (row_vecs >= 256 ? 4 : 2);
        ctas_default = elem_size == 4 ? 3 : 2;
        if (ncw < 8) ctas_default = 4;
    } else if (elem_size == 4) {
        ncw = row_vecs >= 384 ? 3 : (row_vecs >= 128 ? 2 : 1);
        ctas_default = ncw >= 3 ? 4 : (ncw >= 2 ? 6 : 8);
    } else {
        ncw = row_vecs >= 1024 ? 4 : (row_vecs >= 128 ? 2 : 1);
        ctas_default = ncw >= 4 ? 4 : 8;
    }
    int per_thread = 4;                                     // vectors per consumer thread per tile
    const Tuning& t = tuning();
    if (t.stream_threads > 0) ncw = t.stream_threads / 32;
    if (t.rows_threads > 0) per_thread = t.rows_threads;     // (tuning sweeps reuse this knob)
    int64_t tile_vecs = (int64_t)ncw * 32 * per_thread;
    if (tile_vecs > row_vecs) tile_vecs = row_vecs;
    const int64_t stage_bytes = 2 * tile_vecs * 16;
    int ctas = t.stream_ctas_per_sm > 0 ? t.stream_ctas_per_sm : ctas_default;
    int stages = 2;                                          // deeper rings lost every sweep
    if (t.rows_stages > 0) stages = t.rows_stages;
    if (stages > BWD_MAX_STAGES) stages = BWD_MAX_STAGES;
    int max_ctas = 2048 / ((ncw + 1) * 32);
    if (ctas > max_ctas) ctas = max_ctas;
    g.threads = (ncw + 1) * 32;
    g.tile_vecs = (int)tile_vecs;
    g.stages = stages;
    g.ctas_per_sm = ctas;
    g.smem = BWD_SMEM_HEADER + (size_t)stages * (size_t)stage_bytes;
    g.ok = g.smem <= 227 * 1024;
    return g;
}

template <typename T, int RM>
static int launch_rows_bwd(const void* gy, const void* x, const void* scale, const void* gscale, void* gx,
                           int64_t rows, int64_t cols, float int_thr, const QParams& p, int masked, cudaStream_t st) {
    constexpr int V = DT<T>::VEC;
    const bool vec_ok = aligned16(gy) && aligned16(x) && aligned16(gx) && (cols % V) == 0 && cols < (int64_t)1 << 31;
    BwdGeom g = bwd_geometry(cols, (int)sizeof(T));
    if (vec_ok && g.ok && rows < ((int64_t)1 << 31)) {
        static bool attr_set[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(rows_bwd_tma_kernel<T, RM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 227 * 1024);
            if (e != cudaSuccess) return fail(BVB_ECUDA, "rows_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        int64_t grid = (int64_t)resident_ctas(rows_bwd_tma_kernel<T, RM>, g.threads, g.smem, g.ctas_per_sm) * sm_count();
        if (grid > rows) grid = rows;
        rows_bwd_tma_kernel<T, RM><<<(unsigned)grid, g.threads, g.smem, st>>>(
            (const T*)gy, (const T*)x, (const T*)scale, (const T*)gscale, (T*)gx, (int)rows, (int)cols, g.tile_vecs,
            g.stages, int_thr, masked, p);
        return check_launch("bvb_rows_absmax_int_quant_bwd");
    }
    const int64_t nvec = cols / V;
    int threads = 64;
    if (vec_ok) { while (threads < 512 && nvec > (int64_t)threads * 4) threads *= 2; }
    else        { while (threads < 512 && cols > (int64_t)threads * 8) threads *= 2; }
    int per_sm = 2048 / threads;
    if (per_sm > 16) per_sm = 16;
    int64_t grid = rows;
    int64_t cap = (int64_t)sm_count() * per_sm;
    if (grid > cap) grid = cap;
    if (vec_ok)
        rows_bwd_kernel<T, RM, true><<<(unsigned)grid, threads, 0, st>>>((const T*)gy, (const T*)x, (const T*)scale,
                                                                         (const T*)gscale, (T*)gx, rows, cols, int_thr, masked, p);
    else
        rows_bwd_kernel<T, RM, false><<<(unsigned)grid, threads, 0, st>>>((const T*)gy, (const T*)x, (const T*)scale,
                                                                          (const T*)gscale, (T*)gx, rows, cols, int_thr, masked, p);
    return check_launch("bvb_rows_absmax_int_quant_bwd");
}

template <typename T>
static int launch_absmax_tensor(const void* x, int64_t n, void* ws, void* scale_out, void* absmax_out,
                                float min_val, int has_min, float int_thr, int scale_f32, cudaStream_t st) {
    constexpr int V = DT<T>::VEC;
    cudaError_t e = cudaMemsetAsync(ws, 0, 16, st);
    if (e != cudaSuccess) return fail(BVB_ECUDA, "absmax_tensor: memset: %s", cudaGetErrorString(e));
    int vec_ok = aligned16(x) ? 1 : 0;
    int64_t work = vec_ok ? (n / V + 1) : n;
    unsigned grid = stat_grid((work + ST_THREADS * 8 - 1) / (ST_THREADS * 8), 5);
    absmax_tensor_kernel<T><<<grid, ST_THREADS, 0, st>>>((const T*)x, n, vec_ok, (uint32_t*)ws, (T*)scale_out,
                                                         (T*)absmax_out, min_val, has_min, int_thr, scale_f32);
    return check_launch("absmax_tensor");
}

}  // namespace bvb

using namespace bvb;

// scale_dtype may differ from dtype only as "fp32 one-element scale with low-precision x"
#define BVB_CHECK_SCALE_DTYPE(name)                                                         \
    const int scale_f32 = (scale_dtype == BVB_F32 && dtype != BVB_F32) ? 1 : 0;             \
    if (scale_dtype != dtype && !(scale_f32 && scale_count == 1))                           \
        return fail(BVB_EUNSUPPORTED, name ": scale dtype must equal the tensor dtype, or be fp32 with one element");

#define BVB_CHECK_COMMON(name, n)                                                           \
    if ((n) < 0) return fail(BVB_EINVAL, name ": negative size");                           \
    if (!(qmin <= qmax)) return fail(BVB_EINVAL, name ": qmin must be <= qmax");

extern "C" int bvb_debug_packed_constants(float zero_point, float qmin, float qmax, int dtype, uint32_t* out7) {
    if (!out7) return fail(BVB_EINVAL, "bvb_debug_packed_constants: null pointer");
    if (dtype != BVB_F32 && dtype != BVB_BF16 && dtype != BVB_F16) return fail(BVB_EINVAL, "unknown dtype tag %d", dtype);
    const QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    const uint32_t v[7] = {p.pk_ok, p.pk_lo_zero, p.pk_lo, p.pk_hi, p.pk_lo_pre, p.pk_thr_lo, p.pk_thr_hi};
    for (int i = 0; i < 7; ++i) out7[i] = v[i];
    return BVB_OK;
}

static int int_quant_fwd_entry(const char* name, const void* x, const void* scale, void* y, void* codes_out, int64_t n,
                               int64_t scale_inner, int64_t scale_count, int scale_dtype, float zero_point, float qmin,
                               float qmax, int round_mode, int dtype, int pre_relu, void* stream) {
    if (n < 0) return fail(BVB_EINVAL, "%s: negative size", name);
    if (!(qmin <= qmax)) return fail(BVB_EINVAL, "%s: qmin must be <= qmax", name);
    const int scale_f32 = (scale_dtype == BVB_F32 && dtype != BVB_F32) ? 1 : 0;
    if (scale_dtype != dtype && !(scale_f32 && scale_count == 1))
        return fail(BVB_EUNSUPPORTED, "%s: scale dtype must equal the tensor dtype, or be fp32 with one element", name);
    if (n == 0) return BVB_OK;
    if (!x || !scale || !y) return fail(BVB_EINVAL, "%s: null pointer", name);
    if (scale_inner < 1 || scale_count < 1) return fail(BVB_EINVAL, "%s: bad scale broadcast pattern", name);
    QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    p.pre_relu = pre_relu;
    BVB_DISPATCH_DTYPE(dtype, BVB_DISPATCH_MODE_FWD(round_mode, zero_point == 0.f, return (launch_int_quant_fwd<T, RM>(
        x, scale, y, codes_out, n, scale_inner, scale_count, scale_f32, p, 0, (cudaStream_t)stream))));
    return BVB_OK;
}

static int int_quant_bwd_entry(const char* name, const void* gy, const void* x, const void* scale, void* gx,
                               float* gscale_out, int64_t n, int64_t scale_inner, int64_t scale_count, int scale_dtype,
                               float zero_point, float qmin, float qmax, int round_mode, int clamp_mode, int dtype,
                               int pre_relu, void* stream) {
    if (n < 0) return fail(BVB_EINVAL, "%s: negative size", name);
    if (!(qmin <= qmax)) return fail(BVB_EINVAL, "%s: qmin must be <= qmax", name);
    const int scale_f32 = (scale_dtype == BVB_F32 && dtype != BVB_F32) ? 1 : 0;
    if (scale_dtype != dtype && !(scale_f32 && scale_count == 1))
        return fail(BVB_EUNSUPPORTED, "%s: scale dtype must equal the tensor dtype, or be fp32 with one element", name);
    if (scale_inner < 1 || scale_count < 1) return fail(BVB_EINVAL, "%s: bad scale broadcast pattern", name);
    if (n == 0) {
        if (gscale_out) cudaMemsetAsync(gscale_out, 0, sizeof(float) * (size_t)scale_count, (cudaStream_t)stream);
        return BVB_OK;
    }
    if (!gy || !x || !scale || !gx) return fail(BVB_EINVAL, "%s: null pointer", name);
    QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    p.pre_relu = pre_relu;
    const int masked = clamp_mode == BVB_CLAMP_MASKED;
    BVB_DISPATCH_DTYPE(dtype, BVB_DISPATCH_MODE_BWD(round_mode, zero_point == 0.f, masked, return (launch_int_quant_bwd<T, RM>(
        gy, x, scale, gx, gscale_out, n, scale_inner, scale_count, scale_f32, p, masked, (cudaStream_t)stream))));
    return BVB_OK;
}

extern "C" int bvb_int_quant_fwd(const void* x, const void* scale, void* y, void* codes_out, int64_t n,
                                 int64_t scale_inner, int64_t scale_count, int scale_dtype, float zero_point, float qmin,
                                 float qmax, int round_mode, int dtype, void* stream) {
    return int_quant_fwd_entry("bvb_int_quant_fwd", x, scale, y, codes_out, n, scale_inner, scale_count, scale_dtype,
                               zero_point, qmin, qmax, round_mode, dtype, 0, stream);
}

extern "C" int bvb_int_quant_bwd(const void* gy, const void* x, const void* scale, void* gx, float* gscale_out, int64_t n,
                                 int64_t scale_inner, int64_t scale_count, int scale_dtype, float zero_point, float qmin,
                                 float qmax, int round_mode, int clamp_mode, int dtype, void* stream) {
    return int_quant_bwd_entry("bvb_int_quant_bwd", gy, x, scale, gx, gscale_out, n, scale_inner, scale_count, scale_dtype,
                               zero_point, qmin, qmax, round_mode, clamp_mode, dtype, 0, stream);
}

extern "C" int bvb_relu_int_quant_fwd(const void* x, const void* scale, void* y, void* codes_out, int64_t n,
                                      int64_t scale_inner, int64_t scale_count, int scale_dtype, float zero_point,
                                      float qmin, float qmax, int round_mode, int dtype, void* stream) {
    return int_quant_fwd_entry("bvb_relu_int_quant_fwd", x, scale, y, codes_out, n, scale_inner, scale_count, scale_dtype,
                               zero_point, qmin, qmax, round_mode, dtype, 1, stream);
}

extern "C" int bvb_relu_int_quant_bwd(const void* gy, const void* x, const void* scale, void* gx, float* gscale_out,
                                      int64_t n, int64_t scale_inner, int64_t scale_count, int scale_dtype,
                                      float zero_point, float qmin, float qmax, int round_mode, int clamp_mode, int dtype,
                                      void* stream) {
    return int_quant_bwd_entry("bvb_relu_int_quant_bwd", gy, x, scale, gx, gscale_out, n, scale_inner, scale_count,
                               scale_dtype, zero_point, qmin, qmax, round_mode, clamp_mode, dtype, 1, stream);
}

static int zpt_check(const char* name, int64_t n, int64_t scale_inner, int64_t scale_count, int scale_dtype, int zp_dtype,
                     int dtype, float qmin, float qmax, int* scale_f32, int* zp_f32) {
    if (n < 0) return fail(BVB_EINVAL, "%s: negative size", name);
    if (!(qmin <= qmax)) return fail(BVB_EINVAL, "%s: qmin must be <= qmax", name);
    if (scale_inner < 1 || scale_count < 1) return fail(BVB_EINVAL, "%s: bad scale broadcast pattern", name);
    if (scale_count > 1 && n % scale_inner != 0) return fail(BVB_EINVAL, "%s: size is not a multiple of the plane size", name);
    *scale_f32 = (scale_dtype == BVB_F32 && dtype != BVB_F32) ? 1 : 0;
    *zp_f32 = (zp_dtype == BVB_F32 && dtype != BVB_F32) ? 1 : 0;
    if ((scale_dtype != dtype && !(*scale_f32 && scale_count == 1)) || (zp_dtype != dtype && !(*zp_f32 && scale_count == 1)))
        return fail(BVB_EUNSUPPORTED, "%s: scale / zero-point dtype must equal the tensor dtype, or be fp32 with one element", name);
    return BVB_OK;
}

extern "C" int bvb_int_quant_zpt_fwd(const void* x, const void* scale, const void* zero_point, void* y, int64_t n,
                                     int64_t scale_inner, int64_t scale_count, int scale_dtype, int zp_dtype, float qmin,
                                     float qmax, int round_mode, int dtype, void* stream) {
    int scale_f32 = 0, zp_f32 = 0;
    int rc = zpt_check("bvb_int_quant_zpt_fwd", n, scale_inner, scale_count, scale_dtype, zp_dtype, dtype, qmin, qmax,
                       &scale_f32, &zp_f32);
    if (rc != BVB_OK) return rc;
    if (n == 0) return BVB_OK;
    if (!x || !scale || !zero_point || !y) return fail(BVB_EINVAL, "bvb_int_quant_zpt_fwd: null pointer");
    QParams p = make_qparams(1.f, qmin, qmax, dtype);          // non-zero zero-point: literal formulation
    BVB_DISPATCH_DTYPE(dtype, BVB_DISPATCH_ROUND(round_mode, return (launch_int_quant_zpt<T, RM, false>(
        nullptr, x, scale, zero_point, y, nullptr, nullptr, n, scale_inner, scale_count, scale_f32, zp_f32, 0, p,
        (cudaStream_t)stream))));
    return BVB_OK;
}

extern "C" int bvb_int_quant_zpt_bwd(const void* gy, const void* x, const void* scale, const void* zero_point, void* gx,
                                     float* gscale_out, float* gzp_out, int64_t n, int64_t scale_inner,
                                     int64_t scale_count, int scale_dtype, int zp_dtype, float qmin, float qmax,
                                     int round_mode, int clamp_mode, int dtype, void* stream) {
    int scale_f32 = 0, zp_f32 = 0;
    int rc = zpt_check("bvb_int_quant_zpt_bwd", n, scale_inner, scale_count, scale_dtype, zp_dtype, dtype, qmin, qmax,
                       &scale_f32, &zp_f32);
    if (rc != BVB_OK) return rc;
    if ((gscale_out == nullptr) != (gzp_out == nullptr))
        return fail(BVB_EINVAL, "bvb_int_quant_zpt_bwd: pass both gradient outputs or neither");
    if (n == 0) {
        if (gscale_out) {
            cudaMemsetAsync(gscale_out, 0, sizeof(float) * (size_t)scale_count, (cudaStream_t)stream);
            cudaMemsetAsync(gzp_out, 0, sizeof(float) * (size_t)scale_count, (cudaStream_t)stream);
        }
        return BVB_OK;
    }
    if (!gy || !x || !scale || !zero_point || !gx) return fail(BVB_EINVAL, "bvb_int_quant_zpt_bwd: null pointer");
    QParams p = make_qparams(1.f, qmin, qmax, dtype);
    const int masked = clamp_mode == BVB_CLAMP_MASKED;
    BVB_DISPATCH_DTYPE(dtype, BVB_DISPATCH_ROUND(round_mode, return (launch_int_quant_zpt<T, RM, true>(
        gy, x, scale, zero_point, gx, gscale_out, gzp_out, n, scale_inner, scale_count, scale_f32, zp_f32, masked, p,
        (cudaStream_t)stream))));
    return BVB_OK;
}

extern "C" int bvb_int_quant_to_int(const void* x, const void* scale, void* out, int64_t n, int64_t scale_inner,
                                    int64_t scale_count, int scale_dtype, float zero_point, float qmin, float qmax,
                                    int round_mode, int out_kind, int dtype, void* stream) {
    if (n < 0) return fail(BVB_EINVAL, "bvb_int_quant_to_int: negative size");
    if (!(qmin <= qmax)) return fail(BVB_EINVAL, "bvb_int_quant_to_int: qmin must be <= qmax");
    if (out_kind < BVB_OUT_I8 || out_kind > BVB_OUT_I32) return fail(BVB_EINVAL, "bvb_int_quant_to_int: unknown output kind %d", out_kind);
    const float lo_rep = out_kind == BVB_OUT_I8 ? -128.f : (out_kind == BVB_OUT_U8 ? 0.f : -2147483648.f);
    const float hi_rep = out_kind == BVB_OUT_I8 ? 127.f : (out_kind == BVB_OUT_U8 ? 255.f : 2147483520.f);
    if (qmin < lo_rep || qmax > hi_rep)
        return fail(BVB_EINVAL, "bvb_int_quant_to_int: range [%g, %g] does not fit the output dtype", qmin, qmax);
    const int scale_f32 = (scale_dtype == BVB_F32 && dtype != BVB_F32) ? 1 : 0;
    if (scale_dtype != dtype && !(scale_f32 && scale_count == 1))
        return fail(BVB_EUNSUPPORTED, "bvb_int_quant_to_int: scale dtype must equal the tensor dtype, or be fp32 with one element");
    if (n == 0) return BVB_OK;
    if (!x || !scale || !out) return fail(BVB_EINVAL, "bvb_int_quant_to_int: null pointer");
    if (scale_inner < 1 || scale_count < 1) return fail(BVB_EINVAL, "bvb_int_quant_to_int: bad scale broadcast pattern");
    QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    cudaStream_t st = (cudaStream_t)stream;
    BVB_DISPATCH_DTYPE(dtype, BVB_DISPATCH_ROUND(round_mode, {
        constexpr int V = DT<T>::VEC;
        const int out_bytes = out_kind == BVB_OUT_I32 ? 4 : 1;
        const bool vec_ok = aligned16(x) && ((reinterpret_cast<uintptr_t>(out) & (uintptr_t)(V * out_bytes > 16 ? 15 : V * out_bytes - 1)) == 0) &&
                            (scale_count == 1 || (scale_inner % V) == 0);
        const int64_t nvec = vec_ok ? n / V : 0;
        const unsigned grid = stream_grid(vec_ok ? nvec + 1 : n, ST_THREADS * ST_UNROLL);
        if (out_kind == BVB_OUT_I8)
            to_int_kernel<T, RM, int8_t><<<grid, ST_THREADS, 0, st>>>((const T*)x, (const T*)scale, (int8_t*)out, n, nvec,
                                                                      scale_inner, scale_count, scale_f32, p);
        else if (out_kind == BVB_OUT_U8)
            to_int_kernel<T, RM, uint8_t><<<grid, ST_THREADS, 0, st>>>((const T*)x, (const T*)scale, (uint8_t*)out, n, nvec,
                                                                       scale_inner, scale_count, scale_f32, p);
        else
            to_int_kernel<T, RM, int32_t><<<grid, ST_THREADS, 0, st>>>((const T*)x, (const T*)scale, (int32_t*)out, n, nvec,
                                                                       scale_inner, scale_count, scale_f32, p);
    }));
    return check_launch("bvb_int_quant_to_int");
}

extern "C" int bvb_rows_absmax_int_quant_fwd(const void* x, void* y, void* scale_out, void* absmax_out,
                                             int64_t rows, int64_t cols, float scaling_min_val, float int_threshold,
                                             float zero_point, float qmin, float qmax, int round_mode, int dtype,
                                             void* stream) {
    BVB_CHECK_COMMON("bvb_rows_absmax_int_quant_fwd", rows)
    if (cols < 0) return fail(BVB_EINVAL, "bvb_rows_absmax_int_quant_fwd: negative cols");
    if (rows == 0) return BVB_OK;
    if (cols == 0) return fail(BVB_EINVAL, "bvb_rows_absmax_int_quant_fwd: abs-max over an empty row is undefined");
    if (!x || !y || !scale_out) return fail(BVB_EINVAL, "bvb_rows_absmax_int_quant_fwd: null pointer");
    QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    const int has_min = scaling_min_val > 0.f;
    const float mv = round_to_dtype(scaling_min_val, dtype);
    const float thr = int_threshold;      // 0-dim divisor: ATen keeps its fp32 value in opmath (not rounded to T)
    BVB_DISPATCH_DTYPE(dtype, BVB_DISPATCH_MODE_FWD(round_mode, zero_point == 0.f, return (launch_rows_fwd<T, RM>(
        x, y, scale_out, absmax_out, rows, cols, mv, has_min, thr, p, (cudaStream_t)stream))));
    return BVB_OK;
}

extern "C" int bvb_rows_absmax_int_quant_bwd(const void* gy, const void* x, const void* scale, const void* gscale, void* gx,
                                             int64_t rows, int64_t cols, float int_threshold, float zero_point,
                                             float qmin, float qmax, int round_mode, int clamp_mode, int dtype,
                                             void* stream) {
    BVB_CHECK_COMMON("bvb_rows_absmax_int_quant_bwd", rows)
    if (cols < 0) return fail(BVB_EINVAL, "bvb_rows_absmax_int_quant_bwd: negative cols");
    if (rows == 0 || cols == 0) return BVB_OK;
    if (!gy || !x || !scale || !gx) return fail(BVB_EINVAL, "bvb_rows_absmax_int_quant_bwd: null pointer");
    QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    const int masked = clamp_mode == BVB_CLAMP_MASKED;
    const float thr = int_threshold;
    BVB_DISPATCH_DTYPE(dtype, BVB_DISPATCH_MODE_BWD(round_mode, zero_point == 0.f, masked, return (launch_rows_bwd<T, RM>(
        gy, x, scale, gscale, gx, rows, cols, thr, p, masked, (cudaStream_t)stream))));
    return BVB_OK;
}

extern "C" int bvb_tensor_absmax_int_quant_fwd(const void* x, void* y, void* scale_out, void* absmax_out, int64_t n,
                                               int scale_dtype, float scaling_min_val, float int_threshold,
                                               float zero_point, float qmin, float qmax, int round_mode, int dtype,
                                               void* workspace, void* stream) {
    BVB_CHECK_COMMON("bvb_tensor_absmax_int_quant_fwd", n)
    const int64_t scale_count = 1;
    BVB_CHECK_SCALE_DTYPE("bvb_tensor_absmax_int_quant_fwd")
    if (n == 0) return fail(BVB_EINVAL, "bvb_tensor_absmax_int_quant_fwd: abs-max over an empty tensor is undefined");
    if (!x || !y || !scale_out || !workspace) return fail(BVB_EINVAL, "bvb_tensor_absmax_int_quant_fwd: null pointer");
    QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    const int has_min = scaling_min_val > 0.f;
    const float mv = round_to_dtype(scaling_min_val, dtype);
    const float thr = int_threshold;
    cudaStream_t st = (cudaStream_t)stream;
    BVB_DISPATCH_DTYPE(dtype, {
        int rc = launch_absmax_tensor<T>(x, n, workspace, scale_out, absmax_out, mv, has_min, thr, scale_f32, st);
        if (rc != BVB_OK) return rc;
    });
    // second phase runs back to front: the tail of the tensor is what phase 1 left in L2
    BVB_DISPATCH_DTYPE(dtype, BVB_DISPATCH_MODE_FWD(round_mode, zero_point == 0.f, return (launch_int_quant_fwd<T, RM>(
        x, scale_out, y, nullptr, n, 1, 1, scale_f32, p, 1, st))));
    return BVB_OK;
}

extern "C" int bvb_tensor_absmax_int_quant_bwd(const void* gy, const void* x, const void* scale, const void* absmax,
                                               const void* gscale, void* gx, int64_t n, int scale_dtype,
                                               float int_threshold, float zero_point, float qmin, float qmax,
                                               int round_mode, int clamp_mode, int dtype, void* workspace, void* stream) {
    BVB_CHECK_COMMON("bvb_tensor_absmax_int_quant_bwd", n)
    const int64_t scale_count = 1;
    BVB_CHECK_SCALE_DTYPE("bvb_tensor_absmax_int_quant_bwd")
    if (n == 0) return BVB_OK;
    if (!gy || !x || !scale || !absmax || !gx || !workspace)
        return fail(BVB_EINVAL, "bvb_tensor_absmax_int_quant_bwd: null pointer");
    QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    const int masked = clamp_mode == BVB_CLAMP_MASKED;
    const float thr = int_threshold;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, 16, st);
    if (e != cudaSuccess) return fail(BVB_ECUDA, "tensor_bwd: memset: %s", cudaGetErrorString(e));
    BVB_DISPATCH_DTYPE(dtype, BVB_DISPATCH_MODE_BWD(round_mode, zero_point == 0.f, masked, {
        constexpr int V = DT<T>::VEC;
        int vec_ok = (aligned16(gy) && aligned16(x) && aligned16(gx)) ? 1 : 0;
        int64_t work = vec_ok ? (n / V + 1) : n;
        unsigned grid = stream_grid(work, ST_THREADS * ST_UNROLL);
        tensor_bwd_kernel<T, RM><<<grid, ST_THREADS, 0, st>>>((const T*)gy, (const T*)x, (const T*)scale, (const T*)absmax,
                                                              (T*)gx, n, vec_ok, (uint32_t*)workspace, scale_f32, masked, p);
        tensor_bwd_fixup_kernel<T><<<(unsigned)sm_count(), 256, 0, st>>>((const T*)x, (const T*)absmax, (const T*)gscale,
                                                                          (T*)gx, n, (const uint32_t*)workspace, thr, scale_f32);
    }));
    return check_launch("bvb_tensor_absmax_int_quant_bwd");
}

extern "C" int bvb_absmax_rows(const void* x, void* out, int64_t rows, int64_t cols, int dtype, void* stream) {
    if (rows < 0 || cols < 0) return fail(BVB_EINVAL, "bvb_absmax_rows: negative size");
    if (rows == 0) return BVB_OK;
    if (cols == 0) return fail(BVB_EINVAL, "bvb_absmax_rows: abs-max over an empty row is undefined");
    if (!x || !out) return fail(BVB_EINVAL, "bvb_absmax_rows: null pointer");
    QParams p = make_qparams(0.f, 0.f, 0.f, dtype);
    BVB_DISPATCH_DTYPE(dtype, {
        constexpr int V = DT<T>::VEC;
        if (aligned16(x) && (cols % V) == 0) {
            const int64_t row_vecs = cols / V;
            const int group = cols * (int64_t)sizeof(T) <= 4096 ? 32 : AMR_THREADS;
            const int64_t grid = stat_grid((rows + (AMR_THREADS / group) - 1) / (AMR_THREADS / group), 6);
            absmax_rows_vec_kernel<T><<<(unsigned)grid, AMR_THREADS, 0, (cudaStream_t)stream>>>(
                (const T*)x, (T*)out, rows, row_vecs, group);
            return check_launch("bvb_absmax_rows");
        }
        int threads = cols >= 4096 ? 512 : (cols >= 512 ? 256 : 64);
        int64_t grid = rows;
        int64_t cap = (int64_t)sm_count() * (2048 / threads);
        if (grid > cap) grid = cap;
        rows_fwd_generic_kernel<T, 0><<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(
            (const T*)x, nullptr, nullptr, (T*)out, rows, cols, 0.f, 0, 1.f, 0, p);
    });
    return check_launch("bvb_absmax_rows");
}

extern "C" int bvb_absmax_tensor(const void* x, void* out, int64_t n, int dtype, void* workspace, void* stream) {
    if (n < 0) return fail(BVB_EINVAL, "bvb_absmax_tensor: negative size");
    if (n == 0) return fail(BVB_EINVAL, "bvb_absmax_tensor: abs-max over an empty tensor is undefined");
    if (!x || !out || !workspace) return fail(BVB_EINVAL, "bvb_absmax_tensor: null pointer");
    BVB_DISPATCH_DTYPE(dtype, return launch_absmax_tensor<T>(x, n, workspace, nullptr, out, 0.f, 0, 1.f, 0, (cudaStream_t)stream));
    return BVB_OK;
}
