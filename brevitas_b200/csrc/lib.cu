// brevitas_b200 :: library-level C-ABI (version, errors, device info, tuning knobs)
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

Tuning& tuning() {
    static Tuning t = {0, 0, 0, 0, 0};
    return t;
}

}  // namespace bvb

extern "C" {

int bvb_version(void) { return 100; }

const char* bvb_last_error(void) { return bvb::err_buf(); }

int bvb_sm_count(void) { return bvb::sm_count(); }

void bvb_set_tuning(int rows_threads, int rows_stages, int rows_ctas_per_sm, int stream_threads, int stream_ctas_per_sm) {
    bvb::Tuning& t = bvb::tuning();
    t.rows_threads = rows_threads;
    t.rows_stages = rows_stages;
    t.rows_ctas_per_sm = rows_ctas_per_sm;
    t.stream_threads = stream_threads;
    t.stream_ctas_per_sm = stream_ctas_per_sm;
}

int64_t bvb_workspace_bytes(void) { return 1 << 20; }

}  // extern "C"
