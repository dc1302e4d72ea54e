// brevitas_b200 :: library-level C-ABI (version, errors, device info, tuning knobs, numerics self-test)
#include "common.cuh"
#include "host.cuh"

namespace bvb {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(BVB_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    return BVB_OK;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached_dev = dev;
        cached_sms = n;
    }
    return cached_sms;
}

#ifdef BVB_TUNING_BUILD
// sweep build only (make TUNING=1 -> libbrevitas_b200_tuning.so): process-wide launch-geometry overrides
static Tuning g_tuning = {0, 0, 0, 0, 0};
const Tuning& tuning() { return g_tuning; }
#else
// product build: the heuristics are the only geometry source; nothing mutable lives in the library
const Tuning& tuning() {
    static const Tuning none = {0, 0, 0, 0, 0};
    return none;
}
#endif

// DivBy (common.cuh) against the compiler's IEEE division, over `count` consecutive numerator bit patterns
__global__ void selftest_div_kernel(float divisor, uint32_t first_bits, unsigned long long count,
                                    unsigned long long* mismatches) {
    const DivBy dv(divisor);
    unsigned long long bad = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float a = __uint_as_float(first_bits + (uint32_t)i);
        const uint32_t got = __float_as_uint(dv(a));
        const uint32_t ref = __float_as_uint(__fdiv_rn(a, divisor));
        bad += (got != ref) ? 1ull : 0ull;
    }
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}

// the bf16 (fp16) "product with the correctly rounded reciprocal" shortcut against IEEE division rounded to T,
// for ALL pairs (numerator, divisor) of T values: blockIdx.y enumerates divisors, threads enumerate numerators
template <typename T>
__global__ void selftest_lowp_div_kernel(unsigned long long* mismatches, unsigned long long* eligible) {
    unsigned long long bad = 0, elig = 0;
    for (uint32_t bb = blockIdx.x; bb < 65536u; bb += gridDim.x) {
        const unsigned short bs = (unsigned short)bb;
        const float b = DT<T>::to_f(*reinterpret_cast<const T*>(&bs));
        const DivBy dv(b, true);
        if (!dv.mul_only) continue;              // divisor outside the window: the kernels use the full sequence
        for (uint32_t aa = threadIdx.x; aa < 65536u; aa += blockDim.x) {
            const unsigned short as = (unsigned short)aa;
            const float a = DT<T>::to_f(*reinterpret_cast<const T*>(&as));
            const float got = DT<T>::rnd(dv(a));
            const float ref = DT<T>::rnd(__fdiv_rn(a, b));
            const bool same = (__float_as_uint(got) == __float_as_uint(ref)) || (got != got && ref != ref);
            bad += same ? 0ull : 1ull;
            elig += 1ull;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        elig += __shfl_xor_sync(0xffffffffu, elig, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(mismatches, bad);
        atomicAdd(eligible, elig);
    }
}

}  // namespace bvb

extern "C" {

int bvb_selftest_lowp_div(int dtype, uint64_t* out2, void* stream) {
    if (!out2) return bvb::fail(BVB_EINVAL, "bvb_selftest_lowp_div: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(uint64_t), st);
    if (e != cudaSuccess) return bvb::fail(BVB_ECUDA, "bvb_selftest_lowp_div: memset: %s", cudaGetErrorString(e));
    unsigned long long* o = (unsigned long long*)out2;
    if (dtype == BVB_BF16) bvb::selftest_lowp_div_kernel<__nv_bfloat16><<<bvb::sm_count() * 8, 256, 0, st>>>(o, o + 1);
    else if (dtype == BVB_F16) bvb::selftest_lowp_div_kernel<__half><<<bvb::sm_count() * 8, 256, 0, st>>>(o, o + 1);
    else return bvb::fail(BVB_EINVAL, "bvb_selftest_lowp_div: dtype must be BVB_BF16 or BVB_F16");
    return bvb::check_launch("bvb_selftest_lowp_div");
}

int bvb_selftest_div(float divisor, uint32_t first_bits, uint64_t count, uint64_t* mismatches, void* stream) {
    if (!mismatches) return bvb::fail(BVB_EINVAL, "bvb_selftest_div: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(mismatches, 0, sizeof(uint64_t), st);
    if (e != cudaSuccess) return bvb::fail(BVB_ECUDA, "bvb_selftest_div: memset: %s", cudaGetErrorString(e));
    bvb::selftest_div_kernel<<<bvb::sm_count() * 8, 256, 0, st>>>(divisor, first_bits, (unsigned long long)count,
                                                                  (unsigned long long*)mismatches);
    return bvb::check_launch("bvb_selftest_div");
}

int bvb_version(void) { return 100; }

const char* bvb_last_error(void) { return bvb::err_buf(); }

int bvb_sm_count(void) { return bvb::sm_count(); }

#ifdef BVB_TUNING_BUILD
void bvb_set_tuning(int rows_threads, int rows_stages, int rows_ctas_per_sm, int stream_threads, int stream_ctas_per_sm) {
    bvb::g_tuning = {rows_threads, rows_stages, rows_ctas_per_sm, stream_threads, stream_ctas_per_sm};
}
#endif

int64_t bvb_workspace_bytes(void) { return 1 << 20; }

}  // extern "C"
