// brevitas_b200 :: host-buffer entry point of the per-row weight fake-quant step.
//
// The end-to-end cost of the path when tensors live in HOST memory (checkpoint shards being quantized, an optimizer
// that keeps fp32 master weights on the CPU, the bench's e2e leg) is PCIe, not HBM: W and G go in, dW (and
// optionally Wq) and the scales come out.  Rows are independent, so the call cuts the weight into row chunks and
// runs a three-stream pipeline -- H2D of chunk c+1, fwd+bwd kernels of chunk c, D2H of chunk c-1 -- over a small
// ring of device staging slots supplied by the caller.  Both PCIe directions stay busy at the same time (the two
// copies use different copy engines), which the naive "copy everything, compute, copy back" sequence cannot do.
//
// Replaces, for host-resident tensors, the same reference chain as bvb_rows_absmax_int_quant_{fwd,bwd}:
// RescalingIntQuant.forward (src/brevitas/core/quant/int.py:156-163) + autograd through it.
#include <new>

#include "common.cuh"
#include "host.cuh"

using namespace bvb;

namespace {

struct Pipe {
    cudaStream_t in = nullptr, comp = nullptr, out = nullptr;
    cudaEvent_t start = nullptr, done = nullptr;
    cudaEvent_t ev_in[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_comp[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_out[4] = {nullptr, nullptr, nullptr, nullptr};
    int nslots = 0;
    cudaError_t create(int slots) {
        nslots = slots;
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&in, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&comp, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&out, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&start, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming)) != cudaSuccess) return e;
        for (int i = 0; i < slots; ++i) {
            if ((e = cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&ev_comp[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    ~Pipe() {      // destruction is deferred by the runtime until the enqueued work has drained
        for (int i = 0; i < 4; ++i) {
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_comp[i]) cudaEventDestroy(ev_comp[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
        }
        if (start) cudaEventDestroy(start);
        if (done) cudaEventDestroy(done);
        if (in) cudaStreamDestroy(in);
        if (comp) cudaStreamDestroy(comp);
        if (out) cudaStreamDestroy(out);
    }
};

constexpr int HOST_SLOTS = 3;

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace

// bytes of device staging the pipeline needs for a given geometry (chunk_rows rows per chunk)
extern "C" int64_t bvb_host_pipeline_workspace_bytes(int64_t rows, int64_t cols, int64_t chunk_rows, int want_y, int dtype) {
    if (rows < 0 || cols < 0 || chunk_rows <= 0) return -1;
    const int64_t chunk = align_up(chunk_rows * cols * dtype_size(dtype), 256);
    const int64_t per_slot = chunk * (want_y ? 4 : 3);               // W, G, dW (+ Wq)
    return per_slot * HOST_SLOTS + align_up(rows * dtype_size(dtype), 256);
}

#define BVB_CUDA_OK(expr)                                                                          \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(BVB_ECUDA, "bvb_host_rows_fakequant_fwd_bwd: %s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// A caller-owned set of the pipeline's three streams and its events, so that repeated calls do not create and destroy
// 3 streams + 11 events each time.  The library itself still holds no state: the handle lives with the caller.
extern "C" int bvb_host_pipeline_create(void** handle) {
    if (!handle) return fail(BVB_EINVAL, "bvb_host_pipeline_create: null pointer");
    Pipe* p = new (std::nothrow) Pipe();
    if (!p) return fail(BVB_ECUDA, "bvb_host_pipeline_create: out of host memory");
    cudaError_t e = p->create(HOST_SLOTS);
    if (e != cudaSuccess) {
        delete p;
        return fail(BVB_ECUDA, "bvb_host_pipeline_create: %s", cudaGetErrorString(e));
    }
    *handle = p;
    return BVB_OK;
}

extern "C" int bvb_host_pipeline_destroy(void* handle) {
    delete static_cast<Pipe*>(handle);      // the runtime defers stream / event destruction until their work drained
    return BVB_OK;
}

static int host_rows_impl(Pipe& pp, const void* h_x, const void* h_gy, void* h_y, void* h_gx, void* h_scale,
                          int64_t rows, int64_t cols, int64_t chunk_rows, float scaling_min_val, float int_threshold,
                          float zero_point, float qmin, float qmax, int round_mode, int clamp_mode, int dtype,
                          void* workspace, int64_t workspace_bytes, void* stream);

extern "C" int bvb_host_rows_fakequant_fwd_bwd_on(void* pipeline, const void* h_x, const void* h_gy, void* h_y, void* h_gx,
                                                  void* h_scale, int64_t rows, int64_t cols, int64_t chunk_rows,
                                                  float scaling_min_val, float int_threshold, float zero_point, float qmin,
                                                  float qmax, int round_mode, int clamp_mode, int dtype, void* workspace,
                                                  int64_t workspace_bytes, void* stream) {
    if (!pipeline) return fail(BVB_EINVAL, "bvb_host_rows_fakequant_fwd_bwd_on: null pipeline handle");
    return host_rows_impl(*static_cast<Pipe*>(pipeline), h_x, h_gy, h_y, h_gx, h_scale, rows, cols, chunk_rows,
                          scaling_min_val, int_threshold, zero_point, qmin, qmax, round_mode, clamp_mode, dtype, workspace,
                          workspace_bytes, stream);
}

extern "C" int bvb_host_rows_fakequant_fwd_bwd(const void* h_x, const void* h_gy, void* h_y, void* h_gx, void* h_scale,
                                               int64_t rows, int64_t cols, int64_t chunk_rows, float scaling_min_val,
                                               float int_threshold, float zero_point, float qmin, float qmax,
                                               int round_mode, int clamp_mode, int dtype, void* workspace,
                                               int64_t workspace_bytes, void* stream) {
    Pipe pp;                                 // one-shot form: transient streams and events
    cudaError_t e = pp.create(HOST_SLOTS);
    if (e != cudaSuccess) return fail(BVB_ECUDA, "bvb_host_rows_fakequant_fwd_bwd: %s", cudaGetErrorString(e));
    return host_rows_impl(pp, h_x, h_gy, h_y, h_gx, h_scale, rows, cols, chunk_rows, scaling_min_val, int_threshold,
                          zero_point, qmin, qmax, round_mode, clamp_mode, dtype, workspace, workspace_bytes, stream);
}

static int host_rows_impl(Pipe& pp, const void* h_x, const void* h_gy, void* h_y, void* h_gx, void* h_scale,
                          int64_t rows, int64_t cols, int64_t chunk_rows, float scaling_min_val, float int_threshold,
                          float zero_point, float qmin, float qmax, int round_mode, int clamp_mode, int dtype,
                          void* workspace, int64_t workspace_bytes, void* stream) {
    if (rows < 0 || cols < 0) return fail(BVB_EINVAL, "bvb_host_rows_fakequant_fwd_bwd: negative size");
    if (rows == 0) return BVB_OK;
    if (cols == 0) return fail(BVB_EINVAL, "bvb_host_rows_fakequant_fwd_bwd: abs-max over an empty row is undefined");
    if (!h_x || !h_gy || !h_gx || !h_scale || !workspace)
        return fail(BVB_EINVAL, "bvb_host_rows_fakequant_fwd_bwd: null pointer");
    if (dtype != BVB_F32 && dtype != BVB_BF16 && dtype != BVB_F16) return fail(BVB_EINVAL, "unknown dtype tag %d", dtype);
    if (chunk_rows <= 0) chunk_rows = rows;
    if (chunk_rows > rows) chunk_rows = rows;
    const int want_y = h_y != nullptr;
    const int64_t need = bvb_host_pipeline_workspace_bytes(rows, cols, chunk_rows, want_y, dtype);
    if (workspace_bytes < need)
        return fail(BVB_EINVAL, "bvb_host_rows_fakequant_fwd_bwd: workspace of %lld bytes, need %lld",
                    (long long)workspace_bytes, (long long)need);
    const int64_t esz = dtype_size(dtype);
    const int64_t chunk = align_up(chunk_rows * cols * esz, 256);
    const int64_t per_slot = chunk * (want_y ? 4 : 3);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    unsigned char* d_scale = ws + per_slot * HOST_SLOTS;
    cudaStream_t user = (cudaStream_t)stream;

    // order the pipeline after whatever the caller enqueued before (the staging buffers may still be in use)
    BVB_CUDA_OK(cudaEventRecord(pp.start, user));
    BVB_CUDA_OK(cudaStreamWaitEvent(pp.in, pp.start, 0));
    BVB_CUDA_OK(cudaStreamWaitEvent(pp.comp, pp.start, 0));
    BVB_CUDA_OK(cudaStreamWaitEvent(pp.out, pp.start, 0));

    const int64_t nchunks = (rows + chunk_rows - 1) / chunk_rows;
    for (int64_t c = 0; c < nchunks; ++c) {
        const int slot = (int)(c % HOST_SLOTS);
        const int64_t r0 = c * chunk_rows;
        const int64_t nr = (rows - r0 < chunk_rows) ? rows - r0 : chunk_rows;
        const size_t bytes = (size_t)(nr * cols * esz);
        const size_t hoff = (size_t)(r0 * cols * esz);
        unsigned char* d_w = ws + per_slot * slot;
        unsigned char* d_g = d_w + chunk;
        unsigned char* d_gx = d_g + chunk;
        unsigned char* d_y = want_y ? d_gx + chunk : d_gx;     // without h_y the quantized weight is scratch: reuse
        // --- H2D (slot free once the D2H of its previous occupant has finished) ---
        if (c >= HOST_SLOTS) BVB_CUDA_OK(cudaStreamWaitEvent(pp.in, pp.ev_out[slot], 0));
        BVB_CUDA_OK(cudaMemcpyAsync(d_w, static_cast<const unsigned char*>(h_x) + hoff, bytes, cudaMemcpyHostToDevice, pp.in));
        BVB_CUDA_OK(cudaMemcpyAsync(d_g, static_cast<const unsigned char*>(h_gy) + hoff, bytes, cudaMemcpyHostToDevice, pp.in));
        BVB_CUDA_OK(cudaEventRecord(pp.ev_in[slot], pp.in));
        // --- kernels ---
        BVB_CUDA_OK(cudaStreamWaitEvent(pp.comp, pp.ev_in[slot], 0));
        int rc = bvb_rows_absmax_int_quant_fwd(d_w, d_y, d_scale + (size_t)(r0 * esz), nullptr, nr, cols, scaling_min_val,
                                               int_threshold, zero_point, qmin, qmax, round_mode, dtype, pp.comp);
        if (rc != BVB_OK) return rc;
        rc = bvb_rows_absmax_int_quant_bwd(d_g, d_w, d_scale + (size_t)(r0 * esz), nullptr, d_gx, nr, cols, int_threshold,
                                           zero_point, qmin, qmax, round_mode, clamp_mode, dtype, pp.comp);
        if (rc != BVB_OK) return rc;
        BVB_CUDA_OK(cudaEventRecord(pp.ev_comp[slot], pp.comp));
        // --- D2H ---
        BVB_CUDA_OK(cudaStreamWaitEvent(pp.out, pp.ev_comp[slot], 0));
        BVB_CUDA_OK(cudaMemcpyAsync(static_cast<unsigned char*>(h_gx) + hoff, d_gx, bytes, cudaMemcpyDeviceToHost, pp.out));
        if (want_y)
            BVB_CUDA_OK(cudaMemcpyAsync(static_cast<unsigned char*>(h_y) + hoff, d_y, bytes, cudaMemcpyDeviceToHost, pp.out));
        BVB_CUDA_OK(cudaEventRecord(pp.ev_out[slot], pp.out));
    }
    BVB_CUDA_OK(cudaMemcpyAsync(h_scale, d_scale, (size_t)(rows * esz), cudaMemcpyDeviceToHost, pp.out));
    // the caller's stream resumes when the last copy has landed (the H2D and kernel streams finished before it)
    BVB_CUDA_OK(cudaEventRecord(pp.done, pp.out));
    BVB_CUDA_OK(cudaStreamWaitEvent(user, pp.done, 0));
    return BVB_OK;
}
