// brevitas_b200 :: batch-norm + ReLU + activation quantizer of a conv-net block in fused passes (SURVEY.md §8f rank 4,
// VERDICT r1 item 8).
//
// The reference's block `QuantReLU(BatchNorm2d(conv(x)))` (brevitas_examples/imagenet_classification/models/
// mobilenetv1.py:111-115; FusedActivationQuantProxy, src/brevitas/proxy/runtime_quant.py:73-84) runs, per step and per
// activation element, batch-norm forward (statistics read, normalise read + write), the ReLU + quantizer forward
// (read + write), the quantizer + ReLU backward (2 reads + write) and batch-norm backward (2 reads for the
// reductions, 2 reads + write for dx): 13 passes over the activation.  Here:
//
//   forward   bn_stats (1 read)  ->  finalize [C]  ->  bn_act_quant_apply (1 read + 1 write: normalise, ReLU, quantize)
//   backward  bn_act_quant_bwd_reduce (2 reads: recompute y, quantizer + ReLU backward, per-channel sums, d(scale))
//             ->  finalize [C]  ->  bn_act_quant_bwd_dx (2 reads + 1 write)
//
// = 8 passes; the normalised tensor and the ReLU output are never materialised.  Layout: channels-last (NHWC) seen as
// [rows = N*H*W, C]; a thread keeps the same V channels on every row it visits (C / V divides the block), so the
// per-channel constants and divisor set-ups are hoisted and every access is a coalesced 16-byte vector.
//
// Arithmetic: the quantizer part is the literal reference chain of common.cuh (to_int_chain / bwd_elem), bit-identical to
// bvb_relu_int_quant_fwd/bwd applied to the same normalised value.  The normalisation is
// y = ((x - mean) * invstd) * gamma + beta with batch statistics accumulated in fp32 per thread and combined in fp64 in
// a FIXED order (deterministic); it is NOT bit-identical to cuDNN's batch-norm (different summation order: ~1e-7
// relative), which is outside the reference's fake-quant arithmetic.  Tests state the resulting tolerance.
#include "common.cuh"
#include "host.cuh"

namespace bvb {

constexpr int BN_THREADS = 256;
constexpr int BN_UNROLL = 4;
constexpr int BN_CTAS_PER_SM = 3;        // 78 registers x 256 threads: three resident CTAs = one wave
constexpr int BN_SLOTS = 4;            // per-channel partial sums per CTA: fwd {sum, sumsq}; bwd {dbeta, dgamma, dscale}

struct BnGeom { int cv, rl; unsigned grid; bool ok; };

template <typename T>
static BnGeom bn_geometry(int64_t rows, int64_t channels, int ctas_per_sm = BN_CTAS_PER_SM) {
    constexpr int V = DT<T>::VEC;
    BnGeom g = {0, 0, 0, false};
    if (channels < V || channels % V != 0) return g;
    const int64_t cv = channels / V;
    if (cv > BN_THREADS || BN_THREADS % cv != 0) return g;
    g.cv = (int)cv;
    g.rl = BN_THREADS / (int)cv;
    int64_t want = (rows + g.rl - 1) / g.rl;
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    g.grid = (unsigned)want;
    g.ok = true;
    return g;
}

// reduce the V per-thread accumulators of every slot over the row-lanes of a CTA and store the CTA's partial sums
template <int V, int NS>
__device__ __forceinline__ void cta_store_partials(const float (&acc)[NS][V], float* smem, float* partial, int C, int cv_n,
                                                   int rl_n) {
    // smem: [NS][BN_THREADS][V]
    const int tid = threadIdx.x;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < V; ++i) smem[(s * BN_THREADS + tid) * V + i] = acc[s][i];
    __syncthreads();
    if (tid < cv_n) {          // row-lane 0 of every channel vector sums its column, in a fixed order
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            float t[V];
#pragma unroll
            for (int i = 0; i < V; ++i) t[i] = 0.f;
            for (int r = 0; r < rl_n; ++r)
#pragma unroll
                for (int i = 0; i < V; ++i) t[i] += smem[(s * BN_THREADS + r * cv_n + tid) * V + i];
#pragma unroll
            for (int i = 0; i < V; ++i) partial[((size_t)blockIdx.x * BN_SLOTS + s) * C + tid * V + i] = t[i];
        }
    }
}

// ---- forward 1: per-channel sum and sum of squares ------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BN_THREADS, BN_CTAS_PER_SM) bn_stats_kernel(const T* __restrict__ x, float* __restrict__ partial,
                                                             int64_t rows, int C, int cv_n, int rl_n) {
    constexpr int V = DT<T>::VEC;
    __shared__ float smem[2 * BN_THREADS * V];
    const int cv = threadIdx.x % cv_n, rlane = threadIdx.x / cv_n;
    float acc[2][V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[0][i] = acc[1][i] = 0.f;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const int64_t rstride = (int64_t)gridDim.x * rl_n;
    for (int64_t r0 = (int64_t)blockIdx.x * rl_n + rlane; r0 < rows; r0 += rstride * BN_UNROLL) {
        uint4 q[BN_UNROLL];
        bool ok[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const int64_t r = r0 + (int64_t)u * rstride;
            ok[u] = r < rows;
            if (ok[u]) q[u] = ldg_stream(xv + r * cv_n + cv);
        }
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            if (!ok[u]) continue;
            float e[V];
            DT<T>::unpack(q[u], e);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                acc[0][i] += e[i];
                acc[1][i] = fmaf(e[i], e[i], acc[1][i]);
            }
        }
    }
    cta_store_partials<V, 2>(acc, smem, partial, C, cv_n, rl_n);
}

// ---- forward 2: combine the partials (fp64, fixed order: lane-strided, then a shuffle tree), batch statistics,
// running statistics.  One warp per channel.
constexpr int FIN_THREADS = 256;

__device__ __forceinline__ double warp_strided_sum(const float* partial, int nparts, int slot, int C, int c) {
    const int lane = threadIdx.x & 31;
    double t = 0.0;
    for (int p = lane; p < nparts; p += 32) t += (double)partial[((size_t)p * BN_SLOTS + slot) * C + c];
    return warp_sum_d(t);
}

__global__ void __launch_bounds__(FIN_THREADS) bn_fwd_finalize_kernel(
        const float* __restrict__ partial, int nparts, int C, double inv_count, double unbias, float eps, float momentum,
        float* __restrict__ save_mean, float* __restrict__ save_invstd, float* running_mean, float* running_var) {
    const int c = blockIdx.x * (FIN_THREADS / 32) + (threadIdx.x >> 5);
    if (c >= C) return;
    const double s = warp_strided_sum(partial, nparts, 0, C, c);
    const double ss = warp_strided_sum(partial, nparts, 1, C, c);
    if ((threadIdx.x & 31) != 0) return;
    const double mean = s * inv_count;
    double var = ss * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
}

struct BnCh { float mean, invstd, gamma, beta; };

__device__ __forceinline__ float bn_apply(float x, const BnCh& c) {
    return fadd(fmul(fmul(fsub(x, c.mean), c.invstd), c.gamma), c.beta);
}

template <typename T>
__device__ __forceinline__ float load_scale(const void* scale, int scale_f32, int64_t idx) {
    return scale_f32 ? reinterpret_cast<const float*>(scale)[idx] : DT<T>::to_f(reinterpret_cast<const T*>(scale)[idx]);
}

// per-thread channel constants: V channels of the batch-norm and 1 or V divisor set-ups of the quantizer
template <typename T>
struct BnThread {
    static constexpr int V = DT<T>::VEC;
    BnCh ch[V];
    DivBy dv[V];
    float inv_s[V];
    // backward, fp32, zero zero-point: the clamp mask of the reference chain as two comparisons on the normalised value.
    // t1 = y / s (correctly rounded, monotone in y for s > 0) and round() are monotone, so
    //   round(y / s) > qmax  <=>  y >= y_hi,      round(y / s) < qmin  <=>  y <= y_lo
    // for thresholds found once per scale with the exact division (a few probes around (qmax + 0.5) * s); exact, not an
    // approximation (NaN compares false on both sides = gradient kept, like the reference's where-based clamp).
    float y_hi[V], y_lo[V];
    bool fast;              // thresholds valid for every scale of this thread
    __device__ __forceinline__ BnThread(int c0, const float* mean, const float* invstd, const float* gamma,
                                        const float* beta, const void* scale, int scale_count, int scale_f32,
                                        const QParams* p = nullptr) {
        fast = p != nullptr && !DT<T>::LOWP && !p->zp_nonzero;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            ch[i] = {mean[c0 + i], invstd[c0 + i], gamma ? gamma[c0 + i] : 1.f, beta ? beta[c0 + i] : 0.f};
            const float s = load_scale<T>(scale, scale_f32, scale_count == 1 ? 0 : c0 + i);
            dv[i] = DivBy(s, DT<T>::MUL_DIV_EXACT && !scale_f32);
            inv_s[i] = dv[i].approx_recip();
            y_hi[i] = y_lo[i] = 0.f;
            if (fast && (scale_count != 1 || i == 0)) {
                if (!(s > 0.f) || !(s < 3.0e38f) || !dv[i].fast) { fast = false; continue; }
                // smallest t with rintf(t) > qmax / largest t with rintf(t) < qmin (round-half-even)
                float t_hi = p->qmax + 0.5f, t_lo = p->qmin - 0.5f;
                if (!(rintf(t_hi) > p->qmax)) t_hi = nextafterf(t_hi, 3.0e38f);
                if (!(rintf(t_lo) < p->qmin)) t_lo = nextafterf(t_lo, -3.0e38f);
                float yh = fmul(t_hi, s), yl = fmul(t_lo, s);
                int guard = 0;
                while (__fdiv_rn(yh, s) >= t_hi && ++guard < 64) yh = nextafterf(yh, -3.0e38f);     // below the threshold
                while (!(__fdiv_rn(yh, s) >= t_hi) && ++guard < 64) yh = nextafterf(yh, 3.0e38f);   // first value at / above
                while (__fdiv_rn(yl, s) <= t_lo && ++guard < 64) yl = nextafterf(yl, 3.0e38f);
                while (!(__fdiv_rn(yl, s) <= t_lo) && ++guard < 64) yl = nextafterf(yl, -3.0e38f);
                if (guard >= 64) { fast = false; continue; }
                y_hi[i] = yh;
                y_lo[i] = yl;
            } else if (fast) {
                y_hi[i] = y_hi[0];
                y_lo[i] = y_lo[0];
            }
        }
    }
};

// ---- forward 3: normalise, ReLU, quantize ---------------------------------------------------------------------------
// RES: a residual input is added before the ReLU (compile-time: its loads and registers exist only in that variant, which
// runs two resident CTAs per SM instead of three)
constexpr int BN_CTAS_RES = 2;

template <typename T, int RM, bool RES>
__global__ void __launch_bounds__(BN_THREADS, RES ? BN_CTAS_RES : BN_CTAS_PER_SM) bn_act_quant_apply_kernel(
        const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y, const float* __restrict__ mean,
        const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
        const void* __restrict__ scale, int scale_count, int scale_f32, int64_t rows, int cv_n, int rl_n, int relu,
        QParams p) {
    constexpr int V = DT<T>::VEC;
    const int cv = threadIdx.x % cv_n, rlane = threadIdx.x / cv_n;
    const BnThread<T> th(cv * V, mean, invstd, gamma, beta, scale, scale_count, scale_f32);
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const uint4* rv = reinterpret_cast<const uint4*>(res);       // RES: y = quant(relu(bn(x) + res))
    uint4* yv = reinterpret_cast<uint4*>(y);
    const int64_t rstride = (int64_t)gridDim.x * rl_n;
    for (int64_t r0 = (int64_t)blockIdx.x * rl_n + rlane; r0 < rows; r0 += rstride * BN_UNROLL) {
        uint4 q[BN_UNROLL], qr[RES ? BN_UNROLL : 1];
        bool ok[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            const int64_t r = r0 + (int64_t)u * rstride;
            ok[u] = r < rows;
            if (ok[u]) {
                q[u] = ldg_stream(xv + r * cv_n + cv);
                if constexpr (RES) qr[u] = ldg_stream(rv + r * cv_n + cv);
            }
        }
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            if (!ok[u]) continue;
            float e[V], er[RES ? V : 1];
            DT<T>::unpack(q[u], e);
            if constexpr (RES) DT<T>::unpack(qr[u], er);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float v = DT<T>::rnd(bn_apply(e[i], th.ch[i]));
                if constexpr (RES) v = DT<T>::rnd(fadd(v, er[i]));
                if (relu) v = relu_f(v);
                // the reference chain on v (quant_dequant in common.cuh), with the division that does not single zeros
                // out: after a ReLU half of the values are +0
                const float t1 = DT<T>::rnd(th.dv[i].div_zero_unsigned(v));
                float t2 = fadd(t1, p.zp);
                if (DT<T>::LOWP && p.zp_nonzero) t2 = DT<T>::rnd(t2);
                const float t5 = where_clamp(float_to_int<T, RM>(t2), p.qmin, p.qmax);
                float t6 = fsub(t5, p.zp);
                if (DT<T>::LOWP && p.zp_nonzero) t6 = DT<T>::rnd(t6);
                e[i] = fmul(t6, th.dv[i].b);
            }
            stg_stream(yv + (r0 + (int64_t)u * rstride) * cv_n + cv, DT<T>::pack(e));
        }
    }
}

// one element of the backward up to the batch-norm output: returns d(loss)/d(bn output); accumulates d(scale)
template <typename T, int RM>
__device__ __forceinline__ float bn_bwd_elem(float g, float x, const BnThread<T>& th, int i, const QParams& p, int masked,
                                             int relu, bool want_gs, float& gs_acc, float& xhat, bool has_res = false,
                                             float res = 0.f) {
    xhat = fmul(fsub(x, th.ch[i].mean), th.ch[i].invstd);
    float yb = DT<T>::rnd(fadd(fmul(xhat, th.ch[i].gamma), th.ch[i].beta));
    if (has_res) yb = DT<T>::rnd(fadd(yb, res));
    const float v = relu ? relu_f(yb) : yb;
    if (th.fast) {
        // The clamp mask is exact (thresholds above).  The kept gradient is g itself: the reference's ((g * s) / s) equals
        // g to within one rounding, far below the batch-norm reductions this value feeds (dx of the fused path carries a
        // stated tolerance; the unfused path stays bit-exact).
        const bool keep = !(masked && (v >= th.y_hi[i] || v <= th.y_lo[i]));
        const float r0 = keep ? g : 0.f;
        if (want_gs) {
            // d(scale) = g * t5 - r * (y / s): an order-dependent sum; reciprocal-multiply and magic-number rounding of the
            // clamped quotient are within its tolerance (as in bwd_elem)
            const float t1 = v * th.inv_s[i];
            const float c = fminf(fmaxf(t1, p.qmin), p.qmax);
            const float t5 = (c + 12582912.f) - 12582912.f;
            // one term per element: g * (t5 - t1) where the gradient is kept (the rounding residual, |.| <= 0.5: the two
            // large sums sum g*t5 and sum g*t1 never meet in an fp32 accumulator), g * t5 where it is clamped
            gs_acc = fmaf(g, keep ? (t5 - t1) : t5, gs_acc);
        }
        return (relu && yb <= 0.f) ? 0.f : r0;
    }
    float r = bwd_elem<T, RM>(g, v, th.dv[i], th.inv_s[i], p, masked, want_gs, gs_acc);
    r = DT<T>::rnd(r);
    if (relu && yb <= 0.f) r = 0.f;          // ATen's threshold_backward: keeps the gradient unless y <= 0
    return r;
}

// ---- backward 1: per-channel sum(gy), sum(gy * xhat), d(scale) -------------------------------------------------------
template <typename T, int RM, bool RES>
__global__ void __launch_bounds__(BN_THREADS, RES ? BN_CTAS_RES : BN_CTAS_PER_SM) bn_act_quant_bwd_reduce_kernel(
        const T* __restrict__ g, const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ mean,
        const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
        const void* __restrict__ scale, int scale_count, int scale_f32, float* __restrict__ partial, int64_t rows, int C,
        int cv_n, int rl_n, int relu, int masked, int want_gs, QParams p) {
    constexpr int V = DT<T>::VEC;
    __shared__ float smem[3 * BN_THREADS * V];
    const int cv = threadIdx.x % cv_n, rlane = threadIdx.x / cv_n;
    const BnThread<T> th(cv * V, mean, invstd, gamma, beta, scale, scale_count, scale_f32, &p);
    float acc[3][V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[0][i] = acc[1][i] = acc[2][i] = 0.f;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const uint4* gv = reinterpret_cast<const uint4*>(g);
    const uint4* rv = reinterpret_cast<const uint4*>(res);
    const int64_t rstride = (int64_t)gridDim.x * rl_n;
    for (int64_t r0 = (int64_t)blockIdx.x * rl_n + rlane; r0 < rows; r0 += rstride * (BN_UNROLL / 2)) {
        uint4 qx[BN_UNROLL / 2], qg[BN_UNROLL / 2], qr[RES ? BN_UNROLL / 2 : 1];
        bool ok[BN_UNROLL / 2];
#pragma unroll
        for (int u = 0; u < BN_UNROLL / 2; ++u) {
            const int64_t r = r0 + (int64_t)u * rstride;
            ok[u] = r < rows;
            if (ok[u]) {
                qx[u] = ldg_stream(xv + r * cv_n + cv);
                qg[u] = ldg_stream(gv + r * cv_n + cv);
                if constexpr (RES) qr[u] = ldg_stream(rv + r * cv_n + cv);
            }
        }
#pragma unroll
        for (int u = 0; u < BN_UNROLL / 2; ++u) {
            if (!ok[u]) continue;
            float ex[V], eg[V], er[RES ? V : 1];
            DT<T>::unpack(qx[u], ex);
            DT<T>::unpack(qg[u], eg);
            if constexpr (RES) DT<T>::unpack(qr[u], er);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float xhat;
                const float r = bn_bwd_elem<T, RM>(eg[i], ex[i], th, i, p, masked, relu, want_gs != 0, acc[2][i], xhat,
                                                   RES, RES ? er[i] : 0.f);
                acc[0][i] += r;
                acc[1][i] = fmaf(r, xhat, acc[1][i]);
            }
        }
    }
    cta_store_partials<V, 3>(acc, smem, partial, C, cv_n, rl_n);
}

// ---- backward 2: combine -> d(beta), d(gamma), d(scale) per channel (one warp per channel); a scalar scale takes a
// second, single-CTA pass over the C per-channel values (fixed order)
__global__ void __launch_bounds__(FIN_THREADS) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nparts, int C,
                                                                      float* __restrict__ gbeta, float* __restrict__ ggamma,
                                                                      float* __restrict__ gs_channel) {
    const int c = blockIdx.x * (FIN_THREADS / 32) + (threadIdx.x >> 5);
    if (c >= C) return;
    const double b = warp_strided_sum(partial, nparts, 0, C, c);
    const double g = warp_strided_sum(partial, nparts, 1, C, c);
    const double sg = gs_channel ? warp_strided_sum(partial, nparts, 2, C, c) : 0.0;
    if ((threadIdx.x & 31) != 0) return;
    gbeta[c] = (float)b;
    ggamma[c] = (float)g;
    if (gs_channel) gs_channel[c] = (float)sg;
}

__global__ void __launch_bounds__(FIN_THREADS) bn_scalar_gscale_kernel(const float* __restrict__ gs_channel, int C,
                                                                       float* __restrict__ gscale) {
    __shared__ double red[FIN_THREADS / 32];
    double t = 0.0;
    for (int c = threadIdx.x; c < C; c += FIN_THREADS) t += (double)gs_channel[c];
    t = warp_sum_d(t);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x < 32) {
        double u = threadIdx.x < FIN_THREADS / 32 ? red[threadIdx.x] : 0.0;
        u = warp_sum_d(u);
        if (threadIdx.x == 0) gscale[0] = (float)u;
    }
}

// ---- backward 3: dx = gamma * invstd * (gy - mean(gy) - xhat * mean(gy * xhat)) --------------------------------------
template <typename T, int RM, bool RES>
__global__ void __launch_bounds__(BN_THREADS, RES ? BN_CTAS_RES : BN_CTAS_PER_SM) bn_act_quant_bwd_dx_kernel(
        const T* __restrict__ g, const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ gx,
        T* __restrict__ gres, const float* __restrict__ mean,
        const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
        const void* __restrict__ scale, int scale_count, int scale_f32, const float* __restrict__ gbeta,
        const float* __restrict__ ggamma, float inv_count, int64_t rows, int cv_n, int rl_n, int relu, int masked,
        QParams p) {
    constexpr int V = DT<T>::VEC;
    const int cv = threadIdx.x % cv_n, rlane = threadIdx.x / cv_n;
    const BnThread<T> th(cv * V, mean, invstd, gamma, beta, scale, scale_count, scale_f32, &p);
    float a[V], mb[V], mg[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        a[i] = fmul(th.ch[i].gamma, th.ch[i].invstd);
        mb[i] = fmul(gbeta[cv * V + i], inv_count);
        mg[i] = fmul(ggamma[cv * V + i], inv_count);
    }
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const uint4* gv = reinterpret_cast<const uint4*>(g);
    const uint4* rv = reinterpret_cast<const uint4*>(res);
    uint4* ov = reinterpret_cast<uint4*>(gx);
    uint4* orv = reinterpret_cast<uint4*>(gres);                 // nullable: gradient of the residual input
    const int64_t rstride = (int64_t)gridDim.x * rl_n;
    for (int64_t r0 = (int64_t)blockIdx.x * rl_n + rlane; r0 < rows; r0 += rstride * (BN_UNROLL / 2)) {
        uint4 qx[BN_UNROLL / 2], qg[BN_UNROLL / 2], qr[RES ? BN_UNROLL / 2 : 1];
        bool ok[BN_UNROLL / 2];
#pragma unroll
        for (int u = 0; u < BN_UNROLL / 2; ++u) {
            const int64_t r = r0 + (int64_t)u * rstride;
            ok[u] = r < rows;
            if (ok[u]) {
                qx[u] = ldg_stream(xv + r * cv_n + cv);
                qg[u] = ldg_stream(gv + r * cv_n + cv);
                if constexpr (RES) qr[u] = ldg_stream(rv + r * cv_n + cv);
            }
        }
#pragma unroll
        for (int u = 0; u < BN_UNROLL / 2; ++u) {
            if (!ok[u]) continue;
            float ex[V], eg[V], er[RES ? V : 1];
            DT<T>::unpack(qx[u], ex);
            DT<T>::unpack(qg[u], eg);
            if constexpr (RES) DT<T>::unpack(qr[u], er);
            const int64_t o = (r0 + (int64_t)u * rstride) * cv_n + cv;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float xhat, unused = 0.f;
                const float r = bn_bwd_elem<T, RM>(eg[i], ex[i], th, i, p, masked, relu, false, unused, xhat,
                                                   RES, RES ? er[i] : 0.f);
                if constexpr (RES) er[i] = r;                   // d(loss) / d(residual) = d(loss) / d(bn output)
                eg[i] = fmul(a[i], fsub(fsub(r, mb[i]), fmul(xhat, mg[i])));
            }
            stg_stream(ov + o, DT<T>::pack(eg));
            if constexpr (RES) {
                if (orv) stg_stream(orv + o, DT<T>::pack(er));
            }
        }
    }
}

template <typename T>
static int launch_bn_fwd(const void* x, const void* res, const float* gamma, const float* beta, float* running_mean, float* running_var,
                         float momentum, float eps, int use_running, const void* scale, int64_t scale_count,
                         int scale_f32, void* y, float* save_mean, float* save_invstd, int64_t rows, int64_t channels,
                         const QParams& p, int relu, float* workspace, cudaStream_t st) {
    const BnGeom g = bn_geometry<T>(rows, channels, res ? BN_CTAS_RES : BN_CTAS_PER_SM);
    if (!g.ok || !aligned16(x) || !aligned16(y))
        return fail(BVB_EUNSUPPORTED, "bvb_bn_act_quant_fwd: channels = %lld must divide or be divided by %d vectors of "
                    "16 bytes, tensors 16-byte aligned", (long long)channels, BN_THREADS);
    const int C = (int)channels;
    if (!use_running) {
        bn_stats_kernel<T><<<g.grid, BN_THREADS, 0, st>>>((const T*)x, workspace, rows, C, g.cv, g.rl);
        const double inv_count = 1.0 / (double)rows;
        const double unbias = rows > 1 ? (double)rows / (double)(rows - 1) : 1.0;
        bn_fwd_finalize_kernel<<<(C + FIN_THREADS / 32 - 1) / (FIN_THREADS / 32), FIN_THREADS, 0, st>>>(
            workspace, (int)g.grid, C, inv_count, unbias, eps, momentum, save_mean, save_invstd, running_mean, running_var);
    }
    if (res && !aligned16(res)) return fail(BVB_EUNSUPPORTED, "bvb_bn_act_quant_fwd: residual not 16-byte aligned");
    if (res)
        bn_act_quant_apply_kernel<T, RM_ROUND, true><<<g.grid, BN_THREADS, 0, st>>>(
            (const T*)x, (const T*)res, (T*)y, save_mean, save_invstd, gamma, beta, scale, (int)scale_count, scale_f32, rows,
            g.cv, g.rl, relu, p);
    else
        bn_act_quant_apply_kernel<T, RM_ROUND, false><<<g.grid, BN_THREADS, 0, st>>>(
            (const T*)x, nullptr, (T*)y, save_mean, save_invstd, gamma, beta, scale, (int)scale_count, scale_f32, rows,
            g.cv, g.rl, relu, p);
    return check_launch("bvb_bn_act_quant_fwd");
}

template <typename T>
static int launch_bn_bwd(const void* gy, const void* x, const void* res, void* gres, const float* gamma, const float* beta, const float* save_mean,
                         const float* save_invstd, const void* scale, int64_t scale_count, int scale_f32, void* gx,
                         float* ggamma, float* gbeta, float* gscale, int64_t rows, int64_t channels, const QParams& p,
                         int relu, int masked, float* workspace, cudaStream_t st) {
    const BnGeom g = bn_geometry<T>(rows, channels, res ? BN_CTAS_RES : BN_CTAS_PER_SM);
    if (!g.ok || !aligned16(x) || !aligned16(gy) || !aligned16(gx))
        return fail(BVB_EUNSUPPORTED, "bvb_bn_act_quant_bwd: unsupported channel count %lld or alignment", (long long)channels);
    const int C = (int)channels;
    if ((res && !aligned16(res)) || (gres && !aligned16(gres)))
        return fail(BVB_EUNSUPPORTED, "bvb_bn_act_quant_bwd: residual tensors not 16-byte aligned");
    if (res)
        bn_act_quant_bwd_reduce_kernel<T, RM_ROUND, true><<<g.grid, BN_THREADS, 0, st>>>(
            (const T*)gy, (const T*)x, (const T*)res, save_mean, save_invstd, gamma, beta, scale, (int)scale_count, scale_f32,
            workspace, rows, C, g.cv, g.rl, relu, masked, gscale != nullptr, p);
    else
        bn_act_quant_bwd_reduce_kernel<T, RM_ROUND, false><<<g.grid, BN_THREADS, 0, st>>>(
            (const T*)gy, (const T*)x, nullptr, save_mean, save_invstd, gamma, beta, scale, (int)scale_count, scale_f32,
            workspace, rows, C, g.cv, g.rl, relu, masked, gscale != nullptr, p);
    // d(scale): per channel straight into gscale, or (one scale) into scratch behind the partials and then summed
    float* gs_channel = !gscale ? nullptr : (scale_count > 1 ? gscale : workspace + (size_t)g.grid * BN_SLOTS * C);
    bn_bwd_finalize_kernel<<<(C + FIN_THREADS / 32 - 1) / (FIN_THREADS / 32), FIN_THREADS, 0, st>>>(
        workspace, (int)g.grid, C, gbeta, ggamma, gs_channel);
    if (gscale && scale_count == 1) bn_scalar_gscale_kernel<<<1, FIN_THREADS, 0, st>>>(gs_channel, C, gscale);
    const float inv_count = (float)(1.0 / (double)rows);
    if (res)
        bn_act_quant_bwd_dx_kernel<T, RM_ROUND, true><<<g.grid, BN_THREADS, 0, st>>>(
            (const T*)gy, (const T*)x, (const T*)res, (T*)gx, (T*)gres, save_mean, save_invstd, gamma, beta, scale,
            (int)scale_count, scale_f32, gbeta, ggamma, inv_count, rows, g.cv, g.rl, relu, masked, p);
    else
        bn_act_quant_bwd_dx_kernel<T, RM_ROUND, false><<<g.grid, BN_THREADS, 0, st>>>(
            (const T*)gy, (const T*)x, nullptr, (T*)gx, nullptr, save_mean, save_invstd, gamma, beta, scale,
            (int)scale_count, scale_f32, gbeta, ggamma, inv_count, rows, g.cv, g.rl, relu, masked, p);
    return check_launch("bvb_bn_act_quant_bwd");
}

}  // namespace bvb

using namespace bvb;

extern "C" int64_t bvb_bn_act_quant_workspace_bytes(int64_t channels) {
    if (channels < 1) channels = 1;
    return (int64_t)sizeof(float) * channels * (BN_SLOTS * (int64_t)sm_count() * BN_CTAS_PER_SM + 1);
}

extern "C" int bvb_bn_act_quant_fwd(const void* x, const void* residual, const float* gamma, const float* beta, float* running_mean,
                                    float* running_var, float momentum, float eps, int use_running_stats,
                                    const void* scale, int64_t scale_count, int scale_dtype, void* y, float* save_mean,
                                    float* save_invstd, int64_t rows, int64_t channels, float zero_point, float qmin,
                                    float qmax, int round_mode, int relu, int dtype, void* workspace, void* stream) {
    if (rows < 0 || channels < 0) return fail(BVB_EINVAL, "bvb_bn_act_quant_fwd: negative size");
    if (rows == 0 || channels == 0) return BVB_OK;
    if (!x || !y || !scale || !save_mean || !save_invstd || !workspace)
        return fail(BVB_EINVAL, "bvb_bn_act_quant_fwd: null pointer");
    if (round_mode != BVB_ROUND) return fail(BVB_EUNSUPPORTED, "bvb_bn_act_quant_fwd: round-half-even only");
    if (scale_count != 1 && scale_count != channels)
        return fail(BVB_EINVAL, "bvb_bn_act_quant_fwd: one scale or one per channel");
    if (scale_dtype != dtype && !(scale_dtype == BVB_F32 && scale_count == 1))
        return fail(BVB_EINVAL, "bvb_bn_act_quant_fwd: scale dtype must match (a one-element scale may be fp32)");
    const QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    const int scale_f32 = (scale_dtype == BVB_F32 && dtype != BVB_F32) ? 1 : 0;
    BVB_DISPATCH_DTYPE(dtype, return launch_bn_fwd<T>(x, residual, gamma, beta, running_mean, running_var, momentum, eps,
                                                      use_running_stats, scale, scale_count, scale_f32, y, save_mean,
                                                      save_invstd, rows, channels, p, relu, (float*)workspace,
                                                      (cudaStream_t)stream));
    return BVB_OK;
}

extern "C" int bvb_bn_act_quant_bwd(const void* gy, const void* x, const void* residual, void* gresidual,
                                    const float* gamma, const float* beta,
                                    const float* save_mean, const float* save_invstd, const void* scale,
                                    int64_t scale_count, int scale_dtype, void* gx, float* ggamma, float* gbeta,
                                    float* gscale, int64_t rows, int64_t channels, float zero_point, float qmin, float qmax,
                                    int round_mode, int clamp_mode, int relu, int dtype, void* workspace, void* stream) {
    if (rows < 0 || channels < 0) return fail(BVB_EINVAL, "bvb_bn_act_quant_bwd: negative size");
    if (rows == 0 || channels == 0) return BVB_OK;
    if (!gy || !x || !gx || !scale || !save_mean || !save_invstd || !ggamma || !gbeta || !workspace)
        return fail(BVB_EINVAL, "bvb_bn_act_quant_bwd: null pointer");
    if (round_mode != BVB_ROUND) return fail(BVB_EUNSUPPORTED, "bvb_bn_act_quant_bwd: round-half-even only");
    if (scale_count != 1 && scale_count != channels)
        return fail(BVB_EINVAL, "bvb_bn_act_quant_bwd: one scale or one per channel");
    const QParams p = make_qparams(zero_point, qmin, qmax, dtype);
    const int scale_f32 = (scale_dtype == BVB_F32 && dtype != BVB_F32) ? 1 : 0;
    BVB_DISPATCH_DTYPE(dtype, return launch_bn_bwd<T>(gy, x, residual, gresidual, gamma, beta, save_mean, save_invstd, scale, scale_count,
                                                      scale_f32, gx, ggamma, gbeta, gscale, rows, channels, p, relu,
                                                      clamp_mode == BVB_CLAMP_MASKED, (float*)workspace,
                                                      (cudaStream_t)stream));
    return BVB_OK;
}
