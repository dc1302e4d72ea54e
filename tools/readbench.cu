// Read-only streaming micro-benchmark (run on the GPU box): how fast can ONE pass over a tensor that only produces a
// scalar (abs-max, histogram) go on a B200?  Decides the geometry of the statistic kernels.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/readbench tools/readbench.cu && tools/_build/readbench
// Variants
//   ldg   : contiguous slice per CTA, U independent 16-byte ld.global.nc.L1::no_allocate per thread
//   tma   : one producer thread per CTA streams TILE-byte bulk copies through an mbarrier ring, consumer warps reduce
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t amax4(uint32_t m, uint4 q) {
    m = max(m, q.x & 0x7fffffffu); m = max(m, q.y & 0x7fffffffu);
    m = max(m, q.z & 0x7fffffffu); m = max(m, q.w & 0x7fffffffu);
    return m;
}

template <int U>
__global__ void ldg_kernel(const uint4* __restrict__ x, int64_t nvec, uint32_t* out) {
    const int64_t per = (nvec + gridDim.x - 1) / gridDim.x;
    const int64_t lo = blockIdx.x * per, hi = min(lo + per, nvec);
    uint32_t m = 0;
    for (int64_t b = lo + threadIdx.x; b < hi; b += (int64_t)blockDim.x * U) {
        uint4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int64_t v = b + (int64_t)u * blockDim.x;
            q[u] = v < hi ? ldg_stream(x + v) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) m = amax4(m, q[u]);
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// full chunks without predication (the compiler then issues all U loads before the first use), ragged tail one by one
template <int U>
__global__ void ldgb_kernel(const uint4* __restrict__ x, int64_t nvec, uint32_t* out) {
    const int64_t per = (nvec + gridDim.x - 1) / gridDim.x;
    const int64_t lo = blockIdx.x * per, hi = min(lo + per, nvec);
    uint32_t m = 0;
    int64_t b = lo + threadIdx.x;
    for (; b + (int64_t)(U - 1) * blockDim.x < hi; b += (int64_t)blockDim.x * U) {
        uint4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = ldg_stream(x + b + (int64_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < U; ++u) m = amax4(m, q[u]);
    }
    for (; b < hi; b += blockDim.x) m = amax4(m, ldg_stream(x + b));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// four loads in ONE asm statement: the scheduler cannot sink a load below the first use
__device__ __forceinline__ void ldg4(const uint4* p0, const uint4* p1, const uint4* p2, const uint4* p3, uint4& a, uint4& b, uint4& c,
                                     uint4& d) {
    asm volatile(
        "ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%16];\n"
        "ld.global.nc.L1::no_allocate.v4.u32 {%4,%5,%6,%7}, [%17];\n"
        "ld.global.nc.L1::no_allocate.v4.u32 {%8,%9,%10,%11}, [%18];\n"
        "ld.global.nc.L1::no_allocate.v4.u32 {%12,%13,%14,%15}, [%19];\n"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w), "=r"(c.x), "=r"(c.y), "=r"(c.z),
          "=r"(c.w), "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w)
        : "l"(p0), "l"(p1), "l"(p2), "l"(p3));
}
template <int U4>      // U4 groups of four loads
__global__ void ldgq_kernel(const uint4* __restrict__ x, int64_t nvec, uint32_t* out) {
    constexpr int U = U4 * 4;
    const int64_t per = (nvec + gridDim.x - 1) / gridDim.x;
    const int64_t lo = blockIdx.x * per, hi = min(lo + per, nvec);
    uint32_t m = 0;
    int64_t b = lo + threadIdx.x;
    const int64_t bd = blockDim.x;
    for (; b + (int64_t)(U - 1) * bd < hi; b += bd * U) {
        uint4 q[U];
#pragma unroll
        for (int g = 0; g < U4; ++g)
            ldg4(x + b + (4 * g) * bd, x + b + (4 * g + 1) * bd, x + b + (4 * g + 2) * bd, x + b + (4 * g + 3) * bd, q[4 * g],
                 q[4 * g + 1], q[4 * g + 2], q[4 * g + 3]);
#pragma unroll
        for (int u = 0; u < U; ++u) m = amax4(m, q[u]);
    }
    for (; b < hi; b += bd) m = amax4(m, ldg_stream(x + b));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// ---- TMA ring -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_load(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(b)) : "memory");
}

__global__ void tma_kernel(const uint4* __restrict__ x, int64_t nvec, uint32_t* out, int tile_vecs, int stages) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 8;
    uint4* buf = reinterpret_cast<uint4*>(smem + 128);
    const int nconsumers = blockDim.x - 32;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, nconsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ntiles = (nvec + tile_vecs - 1) / tile_vecs;
    const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = blockIdx.x * per, t1 = min(t0 + per, ntiles);
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            for (int64_t t = t0; t < t1; ++t) {
                const int s = (int)((t - t0) % stages);
                const uint32_t k = (uint32_t)((t - t0) / stages);
                if (k > 0) mbar_wait(empty + s, (k - 1) & 1);
                const int64_t v0 = t * tile_vecs;
                const uint32_t bytes = (uint32_t)(min((int64_t)tile_vecs, nvec - v0) * 16);
                mbar_expect(full + s, bytes);
                tma_load(buf + (size_t)s * tile_vecs, x + v0, bytes, full + s);
            }
        }
        return;
    }
    const int ct = threadIdx.x - 32;
    uint32_t m = 0;
    for (int64_t t = t0; t < t1; ++t) {
        const int s = (int)((t - t0) % stages);
        const uint32_t k = (uint32_t)((t - t0) / stages);
        mbar_wait(full + s, k & 1);
        const int64_t v0 = t * tile_vecs;
        const int nv = (int)min((int64_t)tile_vecs, nvec - v0);
        const uint4* b = buf + (size_t)s * tile_vecs;
        for (int v = ct; v < nv; v += nconsumers) m = amax4(m, b[v]);
        __syncwarp();
        if ((ct & 31) == 0) mbar_arrive(empty + s);
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((ct & 31) == 0) atomicMax(out, m);
}

int main(int argc, char** argv) {
    const int reps = 30;
    std::vector<int64_t> sizes = {90177536ll, 180355072ll, 721420288ll};
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint32_t* out;
    CK(cudaMalloc(&out, 4));
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    for (int64_t bytes : sizes) {
        const int nbuf = (int)(800000000ll / bytes) + 2;
        std::vector<uint4*> bufs(nbuf);
        for (auto& b : bufs) { CK(cudaMalloc(&b, bytes)); CK(cudaMemset(b, 0x3c, bytes)); }
        const int64_t nvec = bytes / 16;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        auto timeit = [&](const char* name, auto launch) {
            for (int i = 0; i < 3; ++i) launch(bufs[i % nbuf]);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int i = 0; i < reps; ++i) launch(bufs[i % nbuf]);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("%6.0f MB  %-34s %7.1f us  %6.0f GB/s\n", bytes / 1e6, name, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9);
        };
        char name[96];
        for (int threads : {128, 256, 512, 1024}) {
            for (int per_sm : {1, 2, 4, 8, 16}) {
                if (threads * per_sm > 2048) continue;
                const int grid = sms * per_sm;
                snprintf(name, sizeof name, "ldg U=8 thr=%d cta/sm=%d", threads, per_sm);
                timeit(name, [&](uint4* b) { ldg_kernel<8><<<grid, threads>>>(b, nvec, out); });
                snprintf(name, sizeof name, "ldg U=4 thr=%d cta/sm=%d", threads, per_sm);
                timeit(name, [&](uint4* b) { ldg_kernel<4><<<grid, threads>>>(b, nvec, out); });
                snprintf(name, sizeof name, "ldgb U=8 thr=%d cta/sm=%d", threads, per_sm);
                timeit(name, [&](uint4* b) { ldgb_kernel<8><<<grid, threads>>>(b, nvec, out); });
                snprintf(name, sizeof name, "ldgb U=4 thr=%d cta/sm=%d", threads, per_sm);
                timeit(name, [&](uint4* b) { ldgb_kernel<4><<<grid, threads>>>(b, nvec, out); });
                snprintf(name, sizeof name, "ldgb U=16 thr=%d cta/sm=%d", threads, per_sm);
                timeit(name, [&](uint4* b) { ldgb_kernel<16><<<grid, threads>>>(b, nvec, out); });
                snprintf(name, sizeof name, "ldgq U=4 thr=%d cta/sm=%d", threads, per_sm);
                timeit(name, [&](uint4* b) { ldgq_kernel<1><<<grid, threads>>>(b, nvec, out); });
                snprintf(name, sizeof name, "ldgq U=8 thr=%d cta/sm=%d", threads, per_sm);
                timeit(name, [&](uint4* b) { ldgq_kernel<2><<<grid, threads>>>(b, nvec, out); });
                snprintf(name, sizeof name, "ldgq U=16 thr=%d cta/sm=%d", threads, per_sm);
                timeit(name, [&](uint4* b) { ldgq_kernel<4><<<grid, threads>>>(b, nvec, out); });
                if (threads * per_sm <= 1024) {
                    snprintf(name, sizeof name, "ldg U=16 thr=%d cta/sm=%d", threads, per_sm);
                    timeit(name, [&](uint4* b) { ldg_kernel<16><<<grid, threads>>>(b, nvec, out); });
                }
            }
        }
        for (int tile_kb : {8, 16, 32}) {
            for (int stages : {3, 4, 6}) {
                for (int per_sm : {1, 2}) {
                    for (int cw : {4, 8}) {
                        const size_t smem = 128 + (size_t)stages * tile_kb * 1024;
                        if (smem * per_sm > 220 * 1024) continue;
                        snprintf(name, sizeof name, "tma tile=%dK st=%d cta/sm=%d cw=%d", tile_kb, stages, per_sm, cw);
                        timeit(name, [&](uint4* b) {
                            tma_kernel<<<sms * per_sm, 32 + 32 * cw, smem>>>(b, nvec, out, tile_kb * 64, stages);
                        });
                    }
                }
            }
        }
        for (auto& b : bufs) CK(cudaFree(b));
    }
    return 0;
}
