// Pipe micro-benchmarks: the matcher is bound by the integer POPC / LOP3 pipes and the
// RANSAC scoring by the FP64 pipe; MEASURED_PEAKS.json records neither, so the roofline
// denominators for those kernels are measured here, on the box, in the same run.
#include "common.cuh"

namespace {

constexpr int MB_ITERS = 4096;
constexpr int MB_ILP = 8;

template <int WHICH>
__global__ void __launch_bounds__(512) pipe_kernel(uint32_t* out, uint32_t seed) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (WHICH == 0) {                        // POPC.32 (+1 IADD per POPC, on the other pipe)
        uint32_t x[MB_ILP], acc[MB_ILP];
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) { x[k] = tid * 2654435761u + k * seed; acc[k] = 0; }
        for (int i = 0; i < MB_ITERS; ++i) {
#pragma unroll
            for (int k = 0; k < MB_ILP; ++k) { acc[k] += __popc(x[k]); x[k] += acc[k]; }
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) s += acc[k];
        if (s == 0x12345678u) out[tid] = s;
    } else if (WHICH == 1) {                 // LOP3
        uint32_t x[MB_ILP];
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) x[k] = tid * 2654435761u + k * seed;
        uint32_t a = seed | 1u, b = ~seed;
        for (int i = 0; i < MB_ITERS; ++i) {
#pragma unroll
            for (int k = 0; k < MB_ILP; ++k) x[k] = (x[k] ^ a) | (x[(k + 1) % MB_ILP] & b);
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) s ^= x[k];
        if (s == 0x12345678u) out[tid] = s;
    } else if (WHICH == 2) {                 // FP64 FMA
        double x[MB_ILP];
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) x[k] = 1.0 + 1e-9 * (tid + k);
        double a = 1.0 + 1e-12 * seed, b = 1e-13;
        for (int i = 0; i < MB_ITERS; ++i) {
#pragma unroll
            for (int k = 0; k < MB_ILP; ++k) x[k] = __fma_rn(x[k], a, b);
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) s += x[k];
        if (s == 0.123) out[tid] = 1;
    } else if (WHICH == 3) {                 // FP32 FMA
        float x[MB_ILP];
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) x[k] = 1.0f + 1e-6f * (tid + k);
        float a = 1.0f + 1e-7f * seed, b = 1e-8f;
        for (int i = 0; i < MB_ITERS; ++i) {
#pragma unroll
            for (int k = 0; k < MB_ILP; ++k) x[k] = __fmaf_rn(x[k], a, b);
        }
        float s = 0;
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) s += x[k];
        if (s == 0.123f) out[tid] = 1;
    } else if (WHICH == 5) {                 // FP64 tensor-core MMA m8n8k4 (DMMA): 256 FMA per warp instruction
        double c[MB_ILP][2];
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) { c[k][0] = 1e-9 * (tid + k); c[k][1] = 1e-9 * k; }
        const double a = 1.0 + 1e-12 * seed, b = 1e-3 + 1e-13 * tid;
        for (int i = 0; i < MB_ITERS; ++i) {
#pragma unroll
            for (int k = 0; k < MB_ILP; ++k)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) s += c[k][0] + c[k][1];
        if (s == 0.123) out[tid] = 1;
    } else {                                 // IADD3
        uint32_t x[MB_ILP];
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) x[k] = tid + k * seed;
        for (int i = 0; i < MB_ITERS; ++i) {
#pragma unroll
            for (int k = 0; k < MB_ILP; ++k) x[k] = x[k] + x[(k + 1) % MB_ILP] + seed;
        }
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < MB_ILP; ++k) s += x[k];
        if (s == 0x12345678u) out[tid] = s;
    }
}

}  // namespace

extern "C" int epivo_microbench(epivo_ctx* ctx, int which, double* ops_per_sec) {
    if (!ctx || !ops_per_sec) return EPIVO_ERR_INVALID;
    if (which < 0 || which > 5) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "microbench id %d", which);
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const int threads = 512, blocks = ctx->sm_count * 4;
    int rc = epv_ws_reserve(ctx, (size_t)threads * blocks * 4);
    if (rc) return rc;
    uint32_t* out = epv_ws_take<uint32_t>(ctx, (size_t)threads * blocks);
    cudaEvent_t e0, e1;
    EPV_CUDA(ctx, cudaEventCreate(&e0));
    EPV_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        EPV_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        switch (which) {
            case 0: pipe_kernel<0><<<blocks, threads, 0, ctx->stream>>>(out, 7u + rep); break;
            case 1: pipe_kernel<1><<<blocks, threads, 0, ctx->stream>>>(out, 7u + rep); break;
            case 2: pipe_kernel<2><<<blocks, threads, 0, ctx->stream>>>(out, 7u + rep); break;
            case 3: pipe_kernel<3><<<blocks, threads, 0, ctx->stream>>>(out, 7u + rep); break;
            case 5: pipe_kernel<5><<<blocks, threads, 0, ctx->stream>>>(out, 7u + rep); break;
            default: pipe_kernel<4><<<blocks, threads, 0, ctx->stream>>>(out, 7u + rep); break;
        }
        EPV_LAUNCHED(ctx);
        EPV_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        EPV_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        EPV_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // which = 5: one warp instruction = 256 FMA = 8 per thread
    double ops = (double)threads * blocks * MB_ITERS * MB_ILP * (which == 5 ? 8.0 : 1.0);
    *ops_per_sec = ops / (best * 1e-3);
    return EPIVO_OK;
}
