// brevitas_b200 :: scale statistics that are not fused into a quantizer kernel.
//
//   bvb_abs_kth_value_rows    AbsPercentile.forward   src/brevitas/core/stats/stats_op.py:41-66
//                             = x.abs().kthvalue(k) flat or per row of a 2-D view.  Exact MSB-first radix
//                             select on the bit pattern of |x| (IEEE ordering of non-negative floats equals
//                             the unsigned ordering of their bits; NaN patterns sort above +inf, which is
//                             where torch.kthvalue puts NaN).  One HBM read of x per 8-bit digit
//                             (fp32: 4 passes, bf16/fp16: 2 passes) instead of a sort.
//   bvb_running_stats_update  _RuntimeStats.forward   src/brevitas/core/stats/stats_wrapper.py:56-65
//
// Workspace layout for the select: uint32 hist[4][rows][256], then (8-byte aligned) int64 first_index[rows][256]: the
// smallest element index seen per bin of the LAST digit, recorded during the last pass so that locating the k-th
// value's position (needed by the backward) costs no extra scan of the tensor.  For rows <= KTH_COMPACT_ROWS a
// candidate buffer follows: uint32 count[rows] (256 bytes), uint32 keys[rows][KTH_COMPACT_CAP], int64
// index[rows][KTH_COMPACT_CAP].  As soon as a histogram shows that at most min(KTH_COMPACT_CAP, cols / 8) elements still
// share the resolved prefix, the next pass copies those candidates (key, element index) aside while it scans, and every later
// pass reads the copy instead of the tensor: a 99.999th percentile of an fp32 tensor costs two reads of the tensor
// instead of four.  The decision is a pure function of the (exact) histograms, so every CTA of every pass takes the
// same one without any flag.
#include <type_traits>

#include "common.cuh"
#include "host.cuh"

namespace bvb {

constexpr int KTH_THREADS = 256;
constexpr int KTH_UNROLL = 4;
constexpr int KTH_BINS = 256;
constexpr int KTH_COMPACT_ROWS = 2;
constexpr uint32_t KTH_COMPACT_CAP = 1u << 22;
constexpr int KTH_NO_COMPACT = 99;

// KeyTraits<T, SIGNED>: SIGNED = false orders |x| (AbsPercentile); SIGNED = true orders x itself (the signed k-th value of
// NegativePercentileOrZero / PercentileInterval, stats_op.py:69-126): the usual order-preserving map of IEEE bits to
// unsigned integers (flip all bits of negatives, the sign bit of non-negatives), every NaN mapped above +inf, which is
// where torch.kthvalue puts NaN.
template <typename T, bool SIGNED = false> struct KeyTraits;
// Digit layout, most significant first.  The FIRST digit is the whole 8-bit exponent, so that from the second pass on
// only the elements in the answer's binade are candidates (with byte-aligned digits the first one holds just the top
// 7 exponent bits -- two binades -- and the second pass still histograms a large part of the tensor).
template <> struct KeyTraits<float, false> {
    static constexpr int PASSES = 4;
    __device__ __forceinline__ static int shift(int pass) { return pass == 0 ? 23 : (pass == 1 ? 15 : (pass == 2 ? 7 : 0)); }
    __device__ __forceinline__ static int width(int pass) { return pass == 3 ? 7 : 8; }
    __device__ __forceinline__ static uint32_t key(float v) { return __float_as_uint(v) & 0x7fffffffu; }
    __device__ __forceinline__ static float value(uint32_t k) { return __uint_as_float(k); }
};
template <> struct KeyTraits<__nv_bfloat16, false> {
    static constexpr int PASSES = 2;
    __device__ __forceinline__ static int shift(int pass) { return pass == 0 ? 7 : 0; }      // exponent | 7-bit mantissa
    __device__ __forceinline__ static int width(int pass) { return pass == 0 ? 8 : 7; }
    __device__ __forceinline__ static uint32_t key(float v) { return (__float_as_uint(v) >> 16) & 0x7fffu; }
    __device__ __forceinline__ static float value(uint32_t k) { return __uint_as_float(k << 16); }
};
template <> struct KeyTraits<__half, false> {
    static constexpr int PASSES = 2;
    __device__ __forceinline__ static int shift(int pass) { return pass == 0 ? 8 : 0; }      // 15-bit key: 7 | 8 bits
    __device__ __forceinline__ static int width(int pass) { return pass == 0 ? 7 : 8; }
    __device__ __forceinline__ static uint32_t key(float v) { return (uint32_t)(__half_as_ushort(__float2half_rn(v)) & 0x7fffu); }
    __device__ __forceinline__ static float value(uint32_t k) {
        __half_raw r; r.x = (unsigned short)k; return __half2float(__half(r));
    }
};
// signed keys: 32 (fp32) / 16 (bf16, fp16) bits in byte-aligned digits
__device__ __forceinline__ uint32_t ordered32(uint32_t b) { return b ^ ((b & 0x80000000u) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ uint32_t unordered32(uint32_t k) { return k ^ ((k & 0x80000000u) ? 0x80000000u : 0xffffffffu); }
template <> struct KeyTraits<float, true> {
    static constexpr int PASSES = 4;
    __device__ __forceinline__ static int shift(int pass) { return 24 - 8 * pass; }
    __device__ __forceinline__ static int width(int) { return 8; }
    __device__ __forceinline__ static uint32_t key(float v) { return v != v ? 0xffffffffu : ordered32(__float_as_uint(v)); }
    __device__ __forceinline__ static float value(uint32_t k) { return __uint_as_float(k == 0xffffffffu ? 0x7fc00000u : unordered32(k)); }
};
template <> struct KeyTraits<__nv_bfloat16, true> {
    static constexpr int PASSES = 2;
    __device__ __forceinline__ static int shift(int pass) { return pass == 0 ? 8 : 0; }
    __device__ __forceinline__ static int width(int) { return 8; }
    __device__ __forceinline__ static uint32_t key(float v) { return v != v ? 0xffffu : (ordered32(__float_as_uint(v)) >> 16); }
    __device__ __forceinline__ static float value(uint32_t k) {
        return __uint_as_float(k == 0xffffu ? 0x7fc00000u : unordered32(k << 16) & 0xffff0000u);
    }
};
template <> struct KeyTraits<__half, true> {
    static constexpr int PASSES = 2;
    __device__ __forceinline__ static int shift(int pass) { return pass == 0 ? 8 : 0; }
    __device__ __forceinline__ static int width(int) { return 8; }
    __device__ __forceinline__ static uint32_t key(float v) {
        if (v != v) return 0xffffu;
        const uint32_t h = (uint32_t)__half_as_ushort(__float2half_rn(v));
        return (h ^ ((h & 0x8000u) ? 0xffffu : 0x8000u)) & 0xffffu;
    }
    __device__ __forceinline__ static float value(uint32_t k) {
        if (k == 0xffffu) return __uint_as_float(0x7fc00000u);
        __half_raw r; r.x = (unsigned short)((k ^ ((k & 0x8000u) ? 0x8000u : 0xffffu)) & 0xffffu); return __half2float(__half(r));
    }
};

// Resolve the digits fixed by passes [0, upto) for one row.  Executed by warp 0 of a block; returns
// (prefix, remaining k) to every lane.  hist counts are exact, so the walk is deterministic.
// compact_at (optional): 1 + the first pass whose selected bin holds <= cap elements = the pass that copies the
// candidates aside (KTH_NO_COMPACT if there is none among the resolved passes).
template <typename T, bool SIGNED = false>
__device__ __forceinline__ void kth_resolve(const uint32_t* hist, int64_t rows, int64_t row, int upto, int64_t k,
                                            uint32_t& prefix, int64_t& krem, uint32_t cap = 0, int* compact_at = nullptr) {
    const int lane = threadIdx.x & 31;
    prefix = 0;
    krem = k;
    int cat = KTH_NO_COMPACT;
    for (int q = 0; q < upto; ++q) {
        const uint32_t* h = hist + ((int64_t)q * rows + row) * KTH_BINS;
        // each lane owns 8 consecutive bins
        uint32_t c[8];
        int64_t mine = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = h[lane * 8 + j]; mine += c[j]; }
        // exclusive prefix over lanes
        int64_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int64_t excl = incl - mine;
        // the lane whose range contains the krem-th element (1-indexed)
        bool has = (krem > excl) && (krem <= incl);
        uint32_t who = __ballot_sync(0xffffffffu, has);
        int src = who ? (__ffs(who) - 1) : 31;
        int bin = 0;
        int64_t before = excl;
        uint32_t pop = 0;
        if (lane == src) {
            int64_t run = excl;
            bin = 7;
            before = excl + mine - c[7];
            pop = c[7];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (krem <= run + c[j]) { bin = j; before = run; pop = c[j]; break; }
                run += c[j];
            }
            bin += lane * 8;
        }
        bin = __shfl_sync(0xffffffffu, bin, src);
        before = __shfl_sync(0xffffffffu, before, src);
        pop = __shfl_sync(0xffffffffu, pop, src);
        if (cat == KTH_NO_COMPACT && pop <= cap) cat = q + 1;
        prefix = (prefix << KeyTraits<T, SIGNED>::width(q)) | (uint32_t)bin;
        krem -= before;
    }
    if (compact_at) *compact_at = cat;
}

// one radix pass: histogram digit `pass` of the keys whose higher digits equal the resolved prefix
template <typename T, bool SIGNED, bool RELU>
__global__ void __launch_bounds__(KTH_THREADS, 5) kth_hist_kernel(const T* __restrict__ x, int64_t rows, int64_t cols,
                                                                int vec_ok, int pass, int64_t k, uint32_t* hist,
                                                                unsigned long long* first_index, uint32_t cap,
                                                                uint32_t* ccount, uint32_t* ckeys,
                                                                unsigned long long* cidx) {
    constexpr int V = DT<T>::VEC;
    using KT = KeyTraits<T, SIGNED>;
    constexpr int P = KT::PASSES;
    constexpr bool CAN_COMPACT = P >= 3;       // 2-digit keys: the copying pass would be the last one
    __shared__ uint32_t sh[KTH_THREADS / 32][KTH_BINS];
    __shared__ uint32_t s_prefix;
    __shared__ int s_mode;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int shift = KT::shift(pass);
    const uint32_t digit_mask = (1u << KT::width(pass)) - 1u;
    const int prefix_shift = shift + KT::width(pass);      // the bits above this digit are the resolved prefix
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        for (int i = threadIdx.x; i < (KTH_THREADS / 32) * KTH_BINS; i += KTH_THREADS) (&sh[0][0])[i] = 0;
        if (warp == 0) {
            uint32_t prefix; int64_t krem; int cat;
            kth_resolve<T, SIGNED>(hist, rows, row, pass, k, prefix, krem, cap, &cat);
            if (lane == 0) {
                s_prefix = prefix;
                // 0: scan the tensor   1: scan it and copy the candidates aside   2: scan the copy made by pass `cat`
                s_mode = cat < pass ? 2 : ((cat == pass && pass < P - 1) ? 1 : 0);
                if (pass == 0 && blockIdx.x == 0 && ccount) ccount[row] = 0;
            }
        }
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const int mode = CAN_COMPACT ? s_mode : 0;
        const T* xr = x + row * cols;
        uint32_t* myh = sh[warp];
        uint32_t* ck = ckeys + (size_t)row * KTH_COMPACT_CAP;
        unsigned long long* ci = cidx + (size_t)row * KTH_COMPACT_CAP;
        // Pass 0 counts every element: plain shared atomics on the warp's private histogram (the hardware serialises
        // same-bin lanes; measured faster than match_any aggregation, whose cost is paid per element).  Later passes
        // count only the keys below the resolved prefix -- a small minority for the high percentiles this is used
        // for -- so a warp first votes and skips the histogram update when no lane has a candidate.
        unsigned long long* fi = first_index ? first_index + row * KTH_BINS : nullptr;   // non-null in the last pass only
        // NV elements of one thread (a 16-byte vector, or one element of a ragged tail) at element index j0.  ONE warp
        // vote per vector decides whether any lane holds a candidate at all; everything behind it is rare.
        auto visit = [&](const float* v, auto nv_tag, bool valid, int64_t j0) {
            constexpr int NV = decltype(nv_tag)::value;
            uint32_t key[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) key[i] = KT::key(RELU ? relu_f(v[i]) : v[i]);   // RELU: statistic of relu(x)
            if (pass == 0) {
                if (valid) {
#pragma unroll
                    for (int i = 0; i < NV; ++i) atomicAdd(&myh[key[i] >> shift], 1u);
                }
                return;
            }
            bool take[NV];
            bool mine = false;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                take[i] = valid && ((key[i] >> prefix_shift) == prefix);
                mine = mine || take[i];
            }
            if (__any_sync(0xffffffffu, mine)) {
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    if (take[i]) {
                        const uint32_t bin = (key[i] >> shift) & digit_mask;
                        atomicAdd(&myh[bin], 1u);
                        if (fi) atomicMin(fi + bin, (unsigned long long)(j0 + i));
                    }
                    if (CAN_COMPACT && mode == 1) {   // warp-aggregated append; the total is known to fit (the bin's count)
                        const uint32_t m = __ballot_sync(0xffffffffu, take[i]);
                        if (m) {
                            const int leader = __ffs(m) - 1;
                            uint32_t base = 0;
                            if (lane == leader) base = atomicAdd(ccount + row, (uint32_t)__popc(m));
                            base = __shfl_sync(0xffffffffu, base, leader);
                            if (take[i]) {
                                const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
                                ck[slot] = key[i];
                                ci[slot] = (unsigned long long)(j0 + i);
                            }
                        }
                    }
                }
            }
        };
        if (CAN_COMPACT && mode == 2) {
            const uint32_t cnt = ccount[row];
            for (uint32_t b0 = blockIdx.x * KTH_THREADS; b0 < cnt; b0 += gridDim.x * KTH_THREADS) {
                const uint32_t i = b0 + threadIdx.x;
                if (i < cnt) {
                    const uint32_t key = ck[i];
                    if ((key >> prefix_shift) == prefix) {
                        const uint32_t bin = (key >> shift) & digit_mask;
                        atomicAdd(&myh[bin], 1u);
                        if (fi) atomicMin(fi + bin, ci[i]);
                    }
                }
            }
        }
        const int64_t nvec = (mode == 2) ? 0 : (vec_ok ? cols / V : 0);
        const int64_t scan_end = (mode == 2) ? 0 : cols;
        const uint4* xv = reinterpret_cast<const uint4*>(xr);
        const int64_t gstride = (int64_t)gridDim.x * KTH_THREADS;
        // KTH_UNROLL independent 16-byte loads in flight per thread; every lane of a warp runs the same trip count
        // (bounds are rounded up to the block), so the warp votes above always see a full, converged warp
        for (int64_t b0 = (int64_t)blockIdx.x * KTH_THREADS; b0 < nvec; b0 += gstride * KTH_UNROLL) {
            uint4 q[KTH_UNROLL];
            if (b0 + (int64_t)(KTH_UNROLL - 1) * gstride + KTH_THREADS <= nvec) {     // CTA-uniform: no predication
#pragma unroll
                for (int u = 0; u < KTH_UNROLL; ++u) q[u] = ldg_stream(xv + b0 + (int64_t)u * gstride + threadIdx.x);
#pragma unroll
                for (int u = 0; u < KTH_UNROLL; ++u) {
                    float e[V];
                    DT<T>::unpack(q[u], e);
                    visit(e, std::integral_constant<int, V>{}, true, (b0 + (int64_t)u * gstride + threadIdx.x) * V);
                }
                continue;
            }
            bool ok[KTH_UNROLL];
#pragma unroll
            for (int u = 0; u < KTH_UNROLL; ++u) {
                const int64_t v0 = b0 + (int64_t)u * gstride + threadIdx.x;
                ok[u] = v0 < nvec;
                q[u] = ok[u] ? ldg_stream(xv + v0) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < KTH_UNROLL; ++u) {
                float e[V];
                DT<T>::unpack(q[u], e);
                visit(e, std::integral_constant<int, V>{}, ok[u], (b0 + (int64_t)u * gstride + threadIdx.x) * V);
            }
        }
        for (int64_t b0 = nvec * V + (int64_t)blockIdx.x * KTH_THREADS; b0 < scan_end; b0 += gstride) {
            const int64_t j = b0 + threadIdx.x;
            const bool valid = j < cols;
            const float e1 = valid ? DT<T>::to_f(xr[j]) : 0.f;
            visit(&e1, std::integral_constant<int, 1>{}, valid, j);
        }
        __syncthreads();
        uint32_t* gh = hist + ((int64_t)pass * rows + row) * KTH_BINS;
        for (int b = threadIdx.x; b < KTH_BINS; b += KTH_THREADS) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < KTH_THREADS / 32; ++w) t += sh[w][b];
            if (t) atomicAdd(gh + b, t);
        }
        __syncthreads();
    }
}

// all digits resolved: write the value; optionally locate the smallest index attaining it
template <typename T, bool SIGNED>
__global__ void __launch_bounds__(KTH_THREADS) kth_final_kernel(const T* __restrict__ x, int64_t rows, int64_t cols,
                                                                 int64_t k, const uint32_t* hist, T* out,
                                                                 long long* index_out,
                                                                 const unsigned long long* first_index) {
    using KT = KeyTraits<T, SIGNED>;
    constexpr int P = KT::PASSES;
    __shared__ uint32_t s_key;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        if (warp == 0) {
            uint32_t prefix; int64_t krem;
            kth_resolve<T, SIGNED>(hist, rows, row, P, k, prefix, krem);
            if (lane == 0) s_key = prefix;
        }
        __syncthreads();
        const uint32_t key = s_key;
        if (blockIdx.x == 0 && threadIdx.x == 0) out[row] = DT<T>::from_f(KT::value(key));
        if (index_out && blockIdx.x == 0 && threadIdx.x == 0) {
            const uint32_t last_bin = key & ((1u << KT::width(P - 1)) - 1u);
            index_out[row] = (long long)first_index[row * KTH_BINS + last_bin];
        }
        __syncthreads();
    }
}

static inline int64_t kth_base_bytes(int64_t rows) {
    return (int64_t)sizeof(uint32_t) * 4 * rows * KTH_BINS + (int64_t)sizeof(unsigned long long) * rows * KTH_BINS;
}

template <typename T, bool SIGNED, bool RELU = false>
static int launch_kth(const void* x, void* out, int64_t* index_out, int64_t rows, int64_t cols, int64_t k,
                      void* workspace, cudaStream_t st) {
    constexpr int P = KeyTraits<T, SIGNED>::PASSES;
    constexpr int V = DT<T>::VEC;
    uint32_t* hist = (uint32_t*)workspace;
    unsigned long long* first_index = (unsigned long long*)(hist + (size_t)4 * (size_t)rows * KTH_BINS);
    cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)P * (size_t)rows * KTH_BINS, st);
    if (e != cudaSuccess) return fail(BVB_ECUDA, "bvb_abs_kth_value_rows: memset: %s", cudaGetErrorString(e));
    if (index_out) {
        e = cudaMemsetAsync(first_index, 0xff, sizeof(unsigned long long) * (size_t)rows * KTH_BINS, st);
        if (e != cudaSuccess) return fail(BVB_ECUDA, "bvb_abs_kth_value_rows: memset: %s", cudaGetErrorString(e));
    }
    const int vec_ok = (aligned16(x) && (cols % V) == 0) ? 1 : 0;
    const int64_t work = vec_ok ? cols / V : cols;
    int64_t gx = (work + (int64_t)KTH_THREADS * KTH_UNROLL - 1) / ((int64_t)KTH_THREADS * KTH_UNROLL);
    int64_t gy = rows < 65535 ? rows : 65535;
    const int64_t wave = (int64_t)stat_grid((int64_t)1 << 40, 5);       // CTAs of one wave (tools/statbench.py)
    if (gx * gy > wave) gx = (wave + gy - 1) / gy;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)gy);
    // candidate buffer (see the layout at the top); only worth it when there are passes left after the copying one
    // ... and when the candidates are a small part of the row (copying costs 12 bytes per candidate)
    uint32_t cap = (P >= 3 && rows <= KTH_COMPACT_ROWS) ? KTH_COMPACT_CAP : 0u;
    if ((int64_t)cap > cols / 8) cap = (uint32_t)(cols / 8);
    unsigned char* cbase = (unsigned char*)workspace + kth_base_bytes(rows);
    uint32_t* ccount = cap ? (uint32_t*)cbase : nullptr;
    uint32_t* ckeys = (uint32_t*)(cbase + 256);
    unsigned long long* cidx = (unsigned long long*)(cbase + 256 + sizeof(uint32_t) * (size_t)rows * KTH_COMPACT_CAP);
    for (int pass = 0; pass < P; ++pass)
        kth_hist_kernel<T, SIGNED, RELU><<<grid, KTH_THREADS, 0, st>>>((const T*)x, rows, cols, vec_ok, pass, k, hist,
                                                         (index_out && pass == P - 1) ? first_index : nullptr, cap,
                                                         ccount, ckeys, cidx);
    const dim3 fgrid(1u, (unsigned)gy);                                  // one CTA per row resolves the last digit
    kth_final_kernel<T, SIGNED><<<fgrid, KTH_THREADS, 0, st>>>((const T*)x, rows, cols, k, hist, (T*)out, (long long*)index_out,
                                                       first_index);
    return check_launch("bvb_abs_kth_value_rows");
}

template <typename T>
__global__ void running_stats_kernel(float* running, const T* stat, int64_t count, float momentum,
                                     float one_minus_momentum, int first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float s = DT<T>::to_f(stat[i]);
    float r = running[i];
    if (first) {
        r = fmul(r, s);                                     // running_stats *= out.detach()
    } else {
        r = fmul(r, one_minus_momentum);                    // running_stats *= (1 - momentum)  (python double -> fp32)
        r = fadd(r, DT<T>::rnd(fmul(s, momentum)));         // running_stats += momentum * out.detach()  (T product)
    }
    running[i] = r;
}

}  // namespace bvb

using namespace bvb;

extern "C" int64_t bvb_kth_workspace_bytes(int64_t rows) {
    if (rows < 1) rows = 1;
    int64_t bytes = kth_base_bytes(rows);
    if (rows <= KTH_COMPACT_ROWS) bytes += 256 + (int64_t)rows * KTH_COMPACT_CAP * (sizeof(uint32_t) + sizeof(unsigned long long));
    return bytes;
}

extern "C" int bvb_abs_kth_value_rows(const void* x, void* out, int64_t* index_out, int64_t rows, int64_t cols,
                                      int64_t k, int dtype, void* workspace, void* stream) {
    if (rows < 0 || cols < 0) return fail(BVB_EINVAL, "bvb_abs_kth_value_rows: negative size");
    if (rows == 0) return BVB_OK;
    if (k < 1 || k > cols)
        return fail(BVB_EINVAL, "bvb_abs_kth_value_rows: k = %lld out of range [1, %lld]", (long long)k, (long long)cols);
    if (!x || !out || !workspace) return fail(BVB_EINVAL, "bvb_abs_kth_value_rows: null pointer");
    BVB_DISPATCH_DTYPE(dtype, return (launch_kth<T, false>(x, out, index_out, rows, cols, k, workspace, (cudaStream_t)stream)));
    return BVB_OK;
}

// AbsPercentile of relu(x) without materialising relu(x): the statistic of a QuantReLU's quantizer while it still
// collects (FusedActivationQuantProxy: activation_impl then tensor_quant, proxy/runtime_quant.py:81-84; scaling from
// ParameterFromRuntimeStatsScaling, core/scaling/standalone.py:230-244).  |relu(x)| = relu(x): keys of max(x, +0).
extern "C" int bvb_relu_abs_kth_value_rows(const void* x, void* out, int64_t* index_out, int64_t rows, int64_t cols,
                                           int64_t k, int dtype, void* workspace, void* stream) {
    if (rows < 0 || cols < 0) return fail(BVB_EINVAL, "bvb_relu_abs_kth_value_rows: negative size");
    if (rows == 0) return BVB_OK;
    if (k < 1 || k > cols)
        return fail(BVB_EINVAL, "bvb_relu_abs_kth_value_rows: k = %lld out of range [1, %lld]", (long long)k, (long long)cols);
    if (!x || !out || !workspace) return fail(BVB_EINVAL, "bvb_relu_abs_kth_value_rows: null pointer");
    BVB_DISPATCH_DTYPE(dtype, return (launch_kth<T, false, true>(x, out, index_out, rows, cols, k, workspace, (cudaStream_t)stream)));
    return BVB_OK;
}

// signed k-th smallest value: x.kthvalue(k) of NegativePercentileOrZero / PercentileInterval (stats_op.py:69-126)
extern "C" int bvb_kth_value_rows(const void* x, void* out, int64_t* index_out, int64_t rows, int64_t cols,
                                  int64_t k, int dtype, void* workspace, void* stream) {
    if (rows < 0 || cols < 0) return fail(BVB_EINVAL, "bvb_kth_value_rows: negative size");
    if (rows == 0) return BVB_OK;
    if (k < 1 || k > cols)
        return fail(BVB_EINVAL, "bvb_kth_value_rows: k = %lld out of range [1, %lld]", (long long)k, (long long)cols);
    if (!x || !out || !workspace) return fail(BVB_EINVAL, "bvb_kth_value_rows: null pointer");
    BVB_DISPATCH_DTYPE(dtype, return (launch_kth<T, true>(x, out, index_out, rows, cols, k, workspace, (cudaStream_t)stream)));
    return BVB_OK;
}

// ---- minimum AND maximum of every row with their positions, ONE read ------------------------------------------------
// The asymmetric weight quantizers (ShiftedUint8Weight*: scale from AbsMinMax, stats_op.py:144-158; zero-point from
// NegativeMinOrZero, stats_op.py:22-39) take torch.max once and torch.min twice over the same tensor; this is the three
// reductions in one pass.  Selection semantics of ATen's CUDA reductions: NaN wins, then the value, then the LOWEST index
// (-0.0 and +0.0 compare equal); the outputs are the selected ELEMENTS (read back by position), so their bits are exact.
namespace bvb {
constexpr int MM_THREADS = 256;

// per-thread running extrema: a thread meets its elements in increasing index order, so plain float compares keep the
// FIRST position of a tie (-0.0 == +0.0 included) and NaN is tracked on the side; folded into the packed form once
struct MinMaxThread {
    float lo_v, hi_v;
    uint32_t lo_i, hi_i, nan_i, first_i;
    __device__ __forceinline__ MinMaxThread()
        : lo_v(INFINITY), hi_v(-INFINITY), lo_i(0xffffffffu), hi_i(0xffffffffu),
          nan_i(0xffffffffu), first_i(0xffffffffu) {}
    __device__ __forceinline__ void start(uint32_t idx) { first_i = idx; }
    __device__ __forceinline__ void add(float v, uint32_t idx) {
        nan_i = min(nan_i, (v != v) ? idx : 0xffffffffu);
        if (v < lo_v) { lo_v = v; lo_i = idx; }
        if (v > hi_v) { hi_v = v; hi_i = idx; }
    }
};

struct MinMaxAcc {
    unsigned long long lo = ~0ull, hi = 0ull;
    __device__ __forceinline__ void add(float v, uint32_t idx) {
        uint32_t b = __float_as_uint(v);
        if (b == 0x80000000u) b = 0u;
        const bool nan = v != v;
        const uint32_t k = ordered32(b);
        const unsigned long long a = ((unsigned long long)(nan ? 0u : k) << 32) | idx;
        const unsigned long long c = ((unsigned long long)(nan ? 0xffffffffu : k) << 32) | (uint32_t)~idx;
        lo = a < lo ? a : lo;
        hi = c > hi ? c : hi;
    }
    __device__ __forceinline__ void fold(const MinMaxThread& t) {
        if (t.first_i == 0xffffffffu) return;                        // the thread met no element
        if (t.nan_i != 0xffffffffu) { add(__int_as_float(0x7fc00000), t.nan_i); return; }
        // nothing below +inf (above -inf) was met: every element was +inf (-inf), the first one is selected
        add(t.lo_v, t.lo_i != 0xffffffffu ? t.lo_i : t.first_i);
        add(t.hi_v, t.hi_i != 0xffffffffu ? t.hi_i : t.first_i);
    }
    __device__ __forceinline__ void merge(unsigned long long l, unsigned long long h) {
        lo = l < lo ? l : lo;
        hi = h > hi ? h : hi;
    }
    __device__ __forceinline__ void warp_reduce() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) merge(__shfl_xor_sync(0xffffffffu, lo, o), __shfl_xor_sync(0xffffffffu, hi, o));
    }
};

template <typename T>
__global__ void __launch_bounds__(MM_THREADS) minmax_rows_kernel(const T* __restrict__ x, unsigned long long* partial,
                                                                 int64_t cols, int splits, int64_t per_split, int vec_ok) {
    constexpr int V = DT<T>::VEC;
    __shared__ unsigned long long red[2][MM_THREADS / 32];
    const int64_t row = blockIdx.x / splits;
    const int split = blockIdx.x % splits;
    const int64_t begin = (int64_t)split * per_split;
    int64_t end = begin + per_split;
    if (end > cols) end = cols;
    const T* xr = x + row * cols;
    MinMaxThread acc;
    if (vec_ok) {                                       // per_split and cols are multiples of V, rows 16-byte aligned
        const uint4* xv = reinterpret_cast<const uint4*>(xr);
        const int64_t v1 = end / V;
        int64_t v = begin / V + threadIdx.x;
        if (v < v1) acc.start((uint32_t)(v * V));
        for (; v + 3 * MM_THREADS < v1; v += 4 * MM_THREADS) {          // four independent 16-byte loads in flight
            uint4 q4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q4[u] = ldg_stream(xv + v + u * MM_THREADS);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float e[V];
                DT<T>::unpack(q4[u], e);
#pragma unroll
                for (int i = 0; i < V; ++i) acc.add(e[i], (uint32_t)((v + u * MM_THREADS) * V + i));
            }
        }
        for (; v < v1; v += MM_THREADS) {
            float e[V];
            DT<T>::unpack(ldg_stream(xv + v), e);
#pragma unroll
            for (int i = 0; i < V; ++i) acc.add(e[i], (uint32_t)(v * V + i));
        }
    } else {
        if (begin + threadIdx.x < end) acc.start((uint32_t)(begin + threadIdx.x));
        for (int64_t i = begin + threadIdx.x; i < end; i += MM_THREADS) acc.add(DT<T>::to_f(xr[i]), (uint32_t)i);
    }
    MinMaxAcc packed;
    packed.fold(acc);
    packed.warp_reduce();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = packed.lo; red[1][warp] = packed.hi; }
    __syncthreads();
    if (warp == 0) {
        MinMaxAcc t;
        if (lane < MM_THREADS / 32) t.merge(red[0][lane], red[1][lane]);
        t.warp_reduce();
        if (lane == 0) {
            partial[2 * (size_t)blockIdx.x] = t.lo;
            partial[2 * (size_t)blockIdx.x + 1] = t.hi;
        }
    }
}

template <typename T>
__global__ void minmax_finalize_kernel(const T* __restrict__ x, const unsigned long long* partial, T* mn, T* mx,
                                       int64_t* imn, int64_t* imx, int64_t rows, int64_t cols, int splits) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    MinMaxAcc t;
    for (int sidx = lane; sidx < splits; sidx += 32)
        t.merge(partial[2 * ((size_t)row * splits + sidx)], partial[2 * ((size_t)row * splits + sidx) + 1]);
    t.warp_reduce();
    if (lane == 0) {
        const uint32_t il = (uint32_t)t.lo, ih = ~(uint32_t)t.hi;
        mn[row] = x[row * cols + il];
        mx[row] = x[row * cols + ih];
        if (imn) imn[row] = (int64_t)il;
        if (imx) imx[row] = (int64_t)ih;
    }
}

constexpr int64_t MM_MIN_CTAS = 512;            // few rows: split them until about this many CTAs read (3-4 per SM)
static int minmax_splits(int64_t rows, int64_t cols, int vec) {
    int64_t want = (MM_MIN_CTAS + rows - 1) / rows;
    const int64_t most = (cols + (int64_t)MM_THREADS * vec * 4 - 1) / ((int64_t)MM_THREADS * vec * 4);
    if (want > most) want = most;
    if (want < 1) want = 1;
    return (int)want;
}
}  // namespace bvb

extern "C" int64_t bvb_minmax_workspace_bytes(int64_t rows) {           // one (min, max) pair of 64-bit words per CTA
    if (rows < 1) rows = 1;
    return 16 * rows * ((MM_MIN_CTAS + rows - 1) / rows);
}

extern "C" int bvb_minmax_rows(const void* x, void* min_out, void* max_out, int64_t* argmin_out, int64_t* argmax_out,
                               int64_t rows, int64_t cols, int dtype, void* workspace, void* stream) {
    if (rows < 0 || cols < 0) return fail(BVB_EINVAL, "bvb_minmax_rows: negative size");
    if (rows == 0) return BVB_OK;
    if (cols < 1 || cols >= ((int64_t)1 << 32)) return fail(BVB_EINVAL, "bvb_minmax_rows: 1 <= cols < 2^32 required");
    if (!x || !min_out || !max_out || !workspace) return fail(BVB_EINVAL, "bvb_minmax_rows: null pointer");
    const int vec = 16 / dtype_size(dtype);
    const int splits = minmax_splits(rows, cols, vec);
    if ((int64_t)rows * splits >= ((int64_t)1 << 31)) return fail(BVB_EUNSUPPORTED, "bvb_minmax_rows: too many rows");
    int64_t per = (cols + splits - 1) / splits;
    per = (per + vec - 1) / vec * vec;
    const int vec_ok = aligned16(x) && (cols % vec) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    BVB_DISPATCH_DTYPE(dtype, {
        minmax_rows_kernel<T><<<(unsigned)(rows * splits), MM_THREADS, 0, st>>>(
            (const T*)x, (unsigned long long*)workspace, cols, splits, per, vec_ok);
        minmax_finalize_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(
            (const T*)x, (const unsigned long long*)workspace, (T*)min_out, (T*)max_out, argmin_out, argmax_out, rows, cols,
            splits);
    });
    return check_launch("bvb_minmax_rows");
}

extern "C" int bvb_running_stats_update(float* running, const void* stat, int64_t count, float momentum,
                                        float one_minus_momentum, int first, int dtype, void* stream) {
    if (count < 0) return fail(BVB_EINVAL, "bvb_running_stats_update: negative size");
    if (count == 0) return BVB_OK;
    if (!running || !stat) return fail(BVB_EINVAL, "bvb_running_stats_update: null pointer");
    const unsigned blocks = (unsigned)((count + 255) / 256);
    BVB_DISPATCH_DTYPE(dtype, running_stats_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(
                                  running, (const T*)stat, count, momentum, one_minus_momentum, first));
    return check_launch("bvb_running_stats_update");
}
