// C ABI: context management and the single-call entry points (host buffers in / out).
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "stages.cuh"

int epv_ws_reserve(epivo_ctx* ctx, size_t bytes) {
    ctx->ws_used = 0;
    if (bytes <= ctx->ws_bytes) return EPIVO_OK;
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->ws) EPV_CUDA(ctx, cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    size_t want = epv_align(bytes + bytes / 4, 1 << 20);
    EPV_CUDA(ctx, cudaMalloc(&ctx->ws, want));
    ctx->ws_bytes = want;
    return EPIVO_OK;
}

int epv_pin_reserve(epivo_ctx* ctx, size_t bytes) {
    ctx->pin_used = 0;
    if (bytes <= ctx->pin_bytes) return EPIVO_OK;
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->pin) EPV_CUDA(ctx, cudaFreeHost(ctx->pin));
    ctx->pin = nullptr;
    ctx->pin_bytes = 0;
    size_t want = epv_align(bytes + bytes / 4, 1 << 16);
    EPV_CUDA(ctx, cudaMallocHost(&ctx->pin, want));
    ctx->pin_bytes = want;
    return EPIVO_OK;
}

extern "C" {

const char* epivo_version(void) { return "epivo_b200 0.1 (sm_100a)"; }

int epivo_create(epivo_ctx** out, int device) {
    if (!out) return EPIVO_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) return EPIVO_ERR_CUDA;   // no CPU fallback
    epivo_ctx* ctx = new epivo_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return EPIVO_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return EPIVO_OK;
}

void epivo_destroy(epivo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* epivo_last_error(const epivo_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* epivo_stream(epivo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int64_t epivo_launch_count(const epivo_ctx* ctx) { return ctx ? ctx->launches : 0; }

int epivo_sync(epivo_ctx* ctx) {
    if (!ctx) return EPIVO_ERR_INVALID;
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPIVO_OK;
}

// ---- M1 -----------------------------------------------------------------------------
static int match_common(epivo_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int desc_bytes,
                        int norm, int mode, float ratio, int32_t* query_idx, int32_t* train_idx, int32_t* dist,
                        int32_t* dist2, int* n_out, int32_t* knn_idx, int32_t* knn_dist) {
    if (!ctx) return EPIVO_ERR_INVALID;
    if (nq < 0 || nt < 0 || (nq > 0 && !q) || (nt > 0 && !t))
        EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null descriptors or negative count");
    if (desc_bytes != 16 && desc_bytes != 32 && desc_bytes != 64)
        EPV_FAIL(ctx, EPIVO_ERR_UNSUPPORTED, "descriptor size %d bytes unsupported (16, 32 or 64)", desc_bytes);
    if (mode < 0 || mode > 2) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "match mode %d", mode);
    if (norm != EPIVO_NORM_HAMMING && norm != EPIVO_NORM_HAMMING2)
        EPV_FAIL(ctx, EPIVO_ERR_INVALID, "norm %d is not NORM_HAMMING(6) / NORM_HAMMING2(7)", norm);
    if (n_out) *n_out = 0;
    if (nq == 0) return EPIVO_OK;
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const int words = desc_bytes / 4;
    const int stride = std::max(nq, nt);
    const bool top2 = (mode == EPIVO_MATCH_RATIO) || knn_idx;
    const size_t rows = (size_t)nq + nt;
    const int tsplits = knn_idx ? 1 : epv_match_splits(ctx, 1, nq, nt);
    size_t need = epv_align(rows * desc_bytes) * 2 + epv_align((size_t)stride * 4) * (6 + 2 * tsplits) + 4096;
    int rc = epv_ws_reserve(ctx, need);
    if (rc) return rc;
    rc = epv_pin_reserve(ctx, (size_t)stride * 4 * 5 + 256);
    if (rc) return rc;
    uint32_t* d_desc = epv_ws_take<uint32_t>(ctx, rows * words);
    uint32_t* d_planes = epv_ws_take<uint32_t>(ctx, rows * words);
    uint32_t* d_row = epv_ws_take<uint32_t>(ctx, (size_t)stride * tsplits);
    uint32_t* d_row2 = epv_ws_take<uint32_t>(ctx, (size_t)stride * tsplits);
    uint32_t* d_col = epv_ws_take<uint32_t>(ctx, stride);
    int32_t* d_mq = epv_ws_take<int32_t>(ctx, stride);
    int32_t* d_mt = epv_ws_take<int32_t>(ctx, stride);
    int32_t* d_md = epv_ws_take<int32_t>(ctx, stride);
    int32_t* d_md2 = epv_ws_take<int32_t>(ctx, stride);
    int32_t* d_n = epv_ws_take<int32_t>(ctx, 1);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_desc, q, (size_t)nq * desc_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (nt > 0)
        EPV_CUDA(ctx, cudaMemcpyAsync((char*)d_desc + (size_t)nq * desc_bytes, t, (size_t)nt * desc_bytes,
                                      cudaMemcpyHostToDevice, ctx->stream));
    MatchPlan mp{};
    mp.desc = d_desc;
    mp.planes = d_planes;
    mp.total_rows = (int64_t)rows;
    mp.words = words;
    mp.norm = norm;
    mp.top2 = top2;
    mp.n_pairs = 1;
    mp.q0 = 0;
    mp.qs = 0;
    mp.t0 = nq;
    mp.ts = 0;
    mp.nq = nq;
    mp.nt = nt;
    mp.tsplits = tsplits;
    mp.rowkey = d_row;
    mp.rowkey2 = d_row2;
    mp.colkey = d_col;
    mp.stride = stride;
    rc = epv_match_launch(ctx, mp, true);
    if (rc) return rc;
    if (knn_idx) {
        std::vector<uint32_t> h1(nq), h2(nq);
        EPV_CUDA(ctx, cudaMemcpyAsync(h1.data(), d_row, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
        EPV_CUDA(ctx, cudaMemcpyAsync(h2.data(), d_row2, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
        EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < nq; ++i) {
            bool v1 = nt >= 1, v2 = nt >= 2;
            knn_idx[2 * i] = v1 ? (int32_t)(h1[i] & EPV_IDX_MASK) : -1;
            knn_dist[2 * i] = v1 ? (int32_t)(h1[i] >> EPV_KEY_SHIFT) : -1;
            knn_idx[2 * i + 1] = v2 ? (int32_t)(h2[i] & EPV_IDX_MASK) : -1;
            knn_dist[2 * i + 1] = v2 ? (int32_t)(h2[i] >> EPV_KEY_SHIFT) : -1;
        }
        return EPIVO_OK;
    }
    FinalizePlan fp{};
    fp.n_pairs = 1;
    fp.nq = nq;
    fp.nt = nt;
    fp.stride = stride;
    fp.mode = mode;
    fp.tsplits = tsplits;
    fp.ratio = ratio;
    fp.rowkey = d_row;
    fp.rowkey2 = d_row2;
    fp.colkey = d_col;
    fp.mq = d_mq;
    fp.mt = d_mt;
    fp.md = d_md;
    fp.md2 = d_md2;
    fp.n_matches = d_n;
    rc = epv_finalize_launch(ctx, fp);
    if (rc) return rc;
    int32_t* h = epv_pin_take<int32_t>(ctx, (size_t)stride * 4 + 1);
    EPV_CUDA(ctx, cudaMemcpyAsync(h, d_mq, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(h + stride, d_mt, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(h + 2 * stride, d_md, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(h + 3 * stride, d_md2, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(h + 4 * stride, d_n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int n = h[4 * stride];
    if (query_idx) memcpy(query_idx, h, (size_t)n * 4);
    if (train_idx) memcpy(train_idx, h + stride, (size_t)n * 4);
    if (dist) memcpy(dist, h + 2 * stride, (size_t)n * 4);
    if (dist2 && mode == EPIVO_MATCH_RATIO) memcpy(dist2, h + 3 * stride, (size_t)n * 4);
    if (n_out) *n_out = n;
    return EPIVO_OK;
}

int epivo_match_hamming(epivo_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int desc_bytes,
                        int norm, int mode, float ratio, int32_t* query_idx, int32_t* train_idx, int32_t* dist,
                        int32_t* dist2, int* n_out) {
    return match_common(ctx, q, nq, t, nt, desc_bytes, norm, mode, ratio, query_idx, train_idx, dist, dist2,
                        n_out, nullptr, nullptr);
}

int epivo_knn2_hamming(epivo_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int desc_bytes,
                       int norm, int32_t* train_idx2, int32_t* dist2) {
    if (ctx && (!train_idx2 || !dist2)) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null output");
    return match_common(ctx, q, nq, t, nt, desc_bytes, norm, EPIVO_MATCH_NN, 0.f, nullptr, nullptr, nullptr,
                        nullptr, nullptr, train_idx2, dist2);
}

}  // extern "C"
