// C ABI: context management and the single-call entry points (host buffers in / out).
#include <stdlib.h>
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
    if (const char* e = getenv("EPIVO_NVTX")) ctx->nvtx = e[0] == '1';
    *out = ctx;
    return EPIVO_OK;
}

void epivo_destroy(epivo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
    if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* epivo_last_error(const epivo_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* epivo_stream(epivo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int64_t epivo_launch_count(const epivo_ctx* ctx) { return ctx ? ctx->launches : 0; }

int epivo_last_kernel_ms(epivo_ctx* ctx, float* ms) {
    if (!ctx || !ms) return EPIVO_ERR_INVALID;
    if (!ctx->ev_k0) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "no timed call yet");
    EPV_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_k0, ctx->ev_k1));
    return EPIVO_OK;
}

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
    EpvRange nvtx_range(ctx, "epivo_match_hamming (BFMatcher::match)");
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

// ---- E1/E2, K2, K3 ---------------------------------------------------------------------
extern "C" {

int epivo_find_essential(epivo_ctx* ctx, const float* p0, const float* p1, int n, const double K[9], int method,
                         double prob, double threshold, int max_iters, const int32_t* samples, int m, double E[9],
                         uint8_t* mask, int* n_inliers, int* iters_run) {
    if (!ctx) return EPIVO_ERR_INVALID;
    EpvRange nvtx_range(ctx, "epivo_find_essential (findEssentialMat)");
    if (n < 0 || (n > 0 && (!p0 || !p1)) || !K || !E) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (method != EPIVO_RANSAC && method != EPIVO_LMEDS)
        EPV_FAIL(ctx, EPIVO_ERR_INVALID, "method %d is not RANSAC(8) / LMEDS(4)", method);
    if (!(prob > 0 && prob < 1)) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "prob must be in (0,1)");   // CV_Assert in ptsetreg.cpp
    if (samples && m <= 0) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "empty sample set");
    if (samples)
        for (int i = 0; i < m * 5; ++i)
            if (samples[i] < 0 || samples[i] >= n) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "sample index out of range");
    if (n_inliers) *n_inliers = 0;
    if (iters_run) *iters_run = 0;
    for (int i = 0; i < 9; ++i) E[i] = 0.0;
    if (n < 5) {
        if (mask) memset(mask, 0, (size_t)std::max(n, 0));
        EPV_FAIL(ctx, EPIVO_ERR_NOMODEL, "fewer than 5 correspondences");
    }
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const int stride = n;
    const size_t nerr = epv_essential_errbuf_floats(1, stride);
    const size_t wbytes = epv_essential_work_bytes(1);
    size_t need = epv_align((size_t)n * 8) * 2 + epv_align((size_t)stride * 32) + epv_align(nerr * 4) +
                  epv_align((size_t)n) + epv_align((size_t)std::max(m, 1) * 20) + 8 * 256 + 4096 +
                  epv_align(wbytes) + 512;
    int rc = epv_ws_reserve(ctx, need);
    if (rc) return rc;
    float* d_p0 = epv_ws_take<float>(ctx, (size_t)n * 2);
    float* d_p1 = epv_ws_take<float>(ctx, (size_t)n * 2);
    double* d_xn = epv_ws_take<double>(ctx, (size_t)stride * 4);
    float* d_err = epv_ws_take<float>(ctx, nerr);
    uint8_t* d_mask = epv_ws_take<uint8_t>(ctx, n);
    int32_t* d_samples = samples ? epv_ws_take<int32_t>(ctx, (size_t)m * 5) : nullptr;
    double* d_E = epv_ws_take<double>(ctx, 9);
    int32_t* d_n = epv_ws_take<int32_t>(ctx, 1);
    int32_t* d_out = epv_ws_take<int32_t>(ctx, 4);
    char* d_work = epv_ws_take<char>(ctx, wbytes);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_p0, p0, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_p1, p1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_n, &n, 4, cudaMemcpyHostToDevice, ctx->stream));
    if (samples)
        EPV_CUDA(ctx, cudaMemcpyAsync(d_samples, samples, (size_t)m * 20, cudaMemcpyHostToDevice, ctx->stream));
    rc = epv_normalize_launch(ctx, d_p0, d_p1, n, stride, K, d_xn);
    if (rc) return rc;
    EssentialPlan ep{};
    ep.n_pairs = 1;
    ep.stride = stride;
    ep.xn = d_xn;
    ep.n = d_n;
    ep.method = method;
    ep.prob = prob;
    ep.thresh = threshold / ((K[0] + K[4]) / 2.0);
    ep.max_iters = max_iters;
    ep.samples = d_samples;
    ep.m = m;
    ep.errbuf = d_err;
    ep.E = d_E;
    ep.mask = d_mask;
    ep.n_inliers = d_out;
    ep.iters = d_out + 1;
    ep.n_models = d_out + 2;
    ep.status = d_out + 3;
    ep.work = d_work;
    ep.work_bytes = wbytes;
    rc = epv_essential_launch(ctx, ep);
    if (rc) return rc;
    int32_t h_out[4];
    EPV_CUDA(ctx, cudaMemcpyAsync(E, d_E, 72, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (mask) EPV_CUDA(ctx, cudaMemcpyAsync(mask, d_mask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_inliers) *n_inliers = h_out[0];
    if (iters_run) *iters_run = h_out[1];
    if (h_out[3] != 0) EPV_FAIL(ctx, EPIVO_ERR_NOMODEL, "no essential matrix found");
    return EPIVO_OK;
}

int epivo_five_point(epivo_ctx* ctx, const double* x1, const double* x2, int m, double* E_out, int32_t* n_models) {
    if (!ctx) return EPIVO_ERR_INVALID;
    if (m < 0 || (m > 0 && (!x1 || !x2 || !E_out || !n_models))) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (m == 0) return EPIVO_OK;
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = epv_ws_reserve(ctx, (size_t)m * (80 * 2 + 720 + 4 + 96 * 8 + 40 + 80) + 16384);
    if (rc) return rc;
    rc = epv_pin_reserve(ctx, (size_t)m * (720 + 4) + 1024);
    if (rc) return rc;
    double* d_x1 = epv_ws_take<double>(ctx, (size_t)m * 10);
    double* d_x2 = epv_ws_take<double>(ctx, (size_t)m * 10);
    double* d_E = epv_ws_take<double>(ctx, (size_t)m * 90);
    uint32_t* d_flags = epv_ws_take<uint32_t>(ctx, m);
    double* d_rec = epv_ws_take<double>(ctx, (size_t)m * 96);
    uint32_t* d_items = epv_ws_take<uint32_t>(ctx, (size_t)m * 10);
    double* d_item_z = epv_ws_take<double>(ctx, (size_t)m * 10);
    int32_t* d_count = epv_ws_take<int32_t>(ctx, 1);
    double* h_E = epv_pin_take<double>(ctx, (size_t)m * 90);
    uint32_t* h_flags = epv_pin_take<uint32_t>(ctx, m);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_x1, x1, (size_t)m * 80, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_x2, x2, (size_t)m * 80, cudaMemcpyHostToDevice, ctx->stream));
    rc = epv_five_point_launch(ctx, d_x1, d_x2, m, d_rec, d_items, d_item_z, d_count, d_E, d_flags);
    if (rc) return rc;
    EPV_CUDA(ctx, cudaMemcpyAsync(h_E, d_E, (size_t)m * 720, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(h_flags, d_flags, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // valid models, in root order, packed to the front of each sample's block of ten
    memset(E_out, 0, (size_t)m * 720);
    for (int i = 0; i < m; ++i) {
        int c = 0;
        for (int j = 0; j < 10; ++j)
            if (h_flags[i] >> j & 1u) memcpy(E_out + ((size_t)i * 10 + c++) * 9, h_E + ((size_t)i * 10 + j) * 9, 72);
        n_models[i] = c;
    }
    return EPIVO_OK;
}

int epivo_eight_point(epivo_ctx* ctx, const double* x1, const double* x2, int m, double* E_out, int32_t* ok) {
    if (!ctx) return EPIVO_ERR_INVALID;
    if (m < 0 || (m > 0 && (!x1 || !x2 || !E_out || !ok))) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (m == 0) return EPIVO_OK;
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = epv_ws_reserve(ctx, (size_t)m * (128 * 2 + 72 + 4) + 16384);
    if (rc) return rc;
    double* d_x1 = epv_ws_take<double>(ctx, (size_t)m * 16);
    double* d_x2 = epv_ws_take<double>(ctx, (size_t)m * 16);
    double* d_E = epv_ws_take<double>(ctx, (size_t)m * 9);
    int32_t* d_ok = epv_ws_take<int32_t>(ctx, m);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_x1, x1, (size_t)m * 128, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_x2, x2, (size_t)m * 128, cudaMemcpyHostToDevice, ctx->stream));
    rc = epv_eight_point_launch(ctx, d_x1, d_x2, m, d_E, d_ok);
    if (rc) return rc;
    EPV_CUDA(ctx, cudaMemcpyAsync(E_out, d_E, (size_t)m * 72, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(ok, d_ok, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPIVO_OK;
}

int epivo_fast_detect(epivo_ctx* ctx, const uint8_t* images, int n_images, int rows, int cols, int threshold,
                      int nonmax, int max_kp, float* kps, float* response, int32_t* counts) {
    if (!ctx) return EPIVO_ERR_INVALID;
    if (n_images < 0 || rows < 0 || cols < 0 || max_kp < 0) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "negative size");
    // outside [0, 255] OpenCV's own code paths disagree with each other (its vector path truncates the threshold to
    // 8 bits, its scalar tail does not), so there is no reference result to reproduce
    if (threshold < 0 || threshold > 255) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "threshold %d outside [0, 255]", threshold);
    if (n_images > 0 && (!images || !counts || (max_kp > 0 && !kps))) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (n_images == 0) return EPIVO_OK;
    if (rows < 7 || cols < 7) {                  // no pixel has a full circle: OpenCV returns no keypoints
        for (int i = 0; i < n_images; ++i) counts[i] = 0;
        return EPIVO_OK;
    }
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t npx = (size_t)n_images * rows * cols;
    const size_t nk = (size_t)n_images * max_kp;
    int rc = epv_ws_reserve(ctx, npx + epv_fast_work_bytes(n_images, rows, cols) + nk * 12 + (size_t)n_images * 4 + 4096);
    if (rc) return rc;
    uint8_t* d_img = epv_ws_take<uint8_t>(ctx, npx);
    uint8_t* d_work = epv_ws_take<uint8_t>(ctx, epv_fast_work_bytes(n_images, rows, cols));
    float* d_kps = epv_ws_take<float>(ctx, nk * 2 + 2);
    float* d_resp = response ? epv_ws_take<float>(ctx, nk + 1) : nullptr;
    int32_t* d_counts = epv_ws_take<int32_t>(ctx, n_images);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_img, images, npx, cudaMemcpyHostToDevice, ctx->stream));
    rc = epv_fast_launch(ctx, d_img, n_images, rows, cols, threshold, nonmax, max_kp, d_kps, d_resp, d_counts, d_work);
    if (rc) return rc;
    EPV_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, (size_t)n_images * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (nk) EPV_CUDA(ctx, cudaMemcpyAsync(kps, d_kps, nk * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nk && response) EPV_CUDA(ctx, cudaMemcpyAsync(response, d_resp, nk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPIVO_OK;
}

int epivo_lk_track(epivo_ctx* ctx, const uint8_t* images, int n_frames, int rows, int cols, const float* pts,
                   const int32_t* counts, int max_pts, int max_level, int max_count, double epsilon,
                   double min_eig_threshold, float* next_pts, uint8_t* status, float* err) {
    if (!ctx) return EPIVO_ERR_INVALID;
    if (n_frames < 0 || rows < 0 || cols < 0 || max_pts < 0) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "negative size");
    if (n_frames < 2 || max_pts == 0) return EPIVO_OK;
    if (!images || !pts || !counts || !next_pts || !status) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (max_level < 0 || max_level > 7) EPV_FAIL(ctx, EPIVO_ERR_UNSUPPORTED, "max_level %d outside [0, 7]", max_level);
    // lkpyramid.cpp clamps the criteria the same way
    max_count = std::min(std::max(max_count, 0), 100);
    epsilon = std::min(std::max(epsilon, 0.0), 10.0);
    const int n_pairs = n_frames - 1;
    for (int i = 0; i < n_pairs; ++i)
        if (counts[i] < 0 || counts[i] > max_pts) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "counts[%d] = %d outside [0, %d]", i, counts[i], max_pts);
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t npx = (size_t)n_frames * rows * cols, np = (size_t)n_pairs * max_pts;
    const size_t wb = epv_lk_work_bytes(n_frames, rows, cols, max_level);
    int rc = epv_ws_reserve(ctx, npx + wb + np * 21 + (size_t)n_pairs * 4 + 8192);
    if (rc) return rc;
    uint8_t* d_img = epv_ws_take<uint8_t>(ctx, npx);
    uint8_t* d_work = epv_ws_take<uint8_t>(ctx, wb);
    float* d_pts = epv_ws_take<float>(ctx, np * 2);
    float* d_next = epv_ws_take<float>(ctx, np * 2);
    int32_t* d_counts = epv_ws_take<int32_t>(ctx, n_pairs);
    uint8_t* d_status = epv_ws_take<uint8_t>(ctx, np);
    float* d_err = err ? epv_ws_take<float>(ctx, np) : nullptr;
    if (d_err) EPV_CUDA(ctx, cudaMemsetAsync(d_err, 0, np * 4, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_img, images, npx, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_pts, pts, np * 8, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_counts, counts, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemsetAsync(d_next, 0, np * 8, ctx->stream));
    EPV_CUDA(ctx, cudaMemsetAsync(d_status, 0, np, ctx->stream));
    rc = epv_lk_launch(ctx, d_img, n_frames, rows, cols, d_pts, d_counts, max_pts, max_level, max_count, epsilon,
                       min_eig_threshold, d_next, d_status, d_err, d_work);
    if (rc) return rc;
    EPV_CUDA(ctx, cudaMemcpyAsync(next_pts, d_next, np * 8, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(status, d_status, np, cudaMemcpyDeviceToHost, ctx->stream));
    if (err) EPV_CUDA(ctx, cudaMemcpyAsync(err, d_err, np * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPIVO_OK;
}

int epivo_remap(epivo_ctx* ctx, const uint8_t* images, int n_images, int rows, int cols, const int16_t* map_xy,
                const uint16_t* map_frac, int drows, int dcols, int border_value, uint8_t* out) {
    if (!ctx) return EPIVO_ERR_INVALID;
    if (n_images < 0 || rows < 0 || cols < 0 || drows < 0 || dcols < 0) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "negative size");
    if (n_images == 0 || drows == 0 || dcols == 0) return EPIVO_OK;
    if (!images || !map_xy || !map_frac || !out) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (rows == 0 || cols == 0) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "empty source image");
    if (border_value < 0 || border_value > 255) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "border value outside [0, 255]");
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nsrc = (size_t)n_images * rows * cols, nmap = (size_t)drows * dcols, ndst = (size_t)n_images * nmap;
    int rc = epv_ws_reserve(ctx, nsrc + nmap * 6 + ndst + 4096);
    if (rc) return rc;
    uint8_t* d_img = epv_ws_take<uint8_t>(ctx, nsrc);
    int16_t* d_xy = epv_ws_take<int16_t>(ctx, nmap * 2);
    uint16_t* d_fr = epv_ws_take<uint16_t>(ctx, nmap);
    uint8_t* d_out = epv_ws_take<uint8_t>(ctx, ndst);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_img, images, nsrc, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_xy, map_xy, nmap * 4, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_fr, map_frac, nmap * 2, cudaMemcpyHostToDevice, ctx->stream));
    rc = epv_remap_launch(ctx, d_img, n_images, rows, cols, d_xy, d_fr, drows, dcols, border_value, d_out);
    if (rc) return rc;
    EPV_CUDA(ctx, cudaMemcpyAsync(out, d_out, ndst, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPIVO_OK;
}

int epivo_orb_detect_and_compute(epivo_ctx* ctx, const uint8_t* images, int n_images, int rows, int cols, int nfeatures,
                                 float scale_factor, int nlevels, int edge_threshold, int fast_threshold, int max_kp,
                                 epivo_keypoint* kps, uint8_t* desc, int32_t* counts) {
    if (!ctx) return EPIVO_ERR_INVALID;
    if (n_images < 0 || rows < 0 || cols < 0 || max_kp < 0) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "negative size");
    if (fast_threshold < 0 || fast_threshold > 255) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "fast_threshold %d outside [0, 255]", fast_threshold);
    if (n_images > 0 && (!images || !counts || (max_kp > 0 && (!kps || !desc)))) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (n_images == 0) return EPIVO_OK;
    if (rows == 0 || cols == 0) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "empty image");
    EpvOrbPlan plan;
    int rc = epv_orb_plan(ctx, n_images, rows, cols, nfeatures, scale_factor, nlevels, edge_threshold, max_kp, &plan);
    if (rc) return rc;
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nk = (size_t)n_images * max_kp;
    rc = epv_ws_reserve(ctx, epv_orb_work_bytes(plan) + nk * (sizeof(epivo_keypoint) + 32) + (size_t)n_images * 4 + 4096);
    if (rc) return rc;
    rc = epv_pin_reserve(ctx, (plan.tab_entries + 1) * 4 + 512);
    if (rc) return rc;
    uint8_t* d_pyr = epv_ws_take<uint8_t>(ctx, plan.pyr_bytes);
    epivo_keypoint* d_kps = epv_ws_take<epivo_keypoint>(ctx, nk + 1);
    uint8_t* d_desc = epv_ws_take<uint8_t>(ctx, nk * 32 + 32);
    int32_t* d_counts = epv_ws_take<int32_t>(ctx, n_images);
    uint32_t* h_tab = epv_pin_take<uint32_t>(ctx, plan.tab_entries + 1);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_pyr, images, (size_t)n_images * rows * cols, cudaMemcpyHostToDevice, ctx->stream));
    rc = epv_orb_launch(ctx, plan, fast_threshold, d_pyr, d_kps, d_desc, d_counts, h_tab);
    if (rc) return rc;
    EPV_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, (size_t)n_images * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // only what was found travels back: image i's first min(counts[i], max_kp) keypoints and descriptors
    for (int i = 0; i < n_images; ++i) {
        const size_t k = (size_t)std::min(std::max(counts[i], 0), max_kp);
        if (!k) continue;
        EPV_CUDA(ctx, cudaMemcpyAsync(kps + (size_t)i * max_kp, d_kps + (size_t)i * max_kp, k * sizeof(epivo_keypoint),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        EPV_CUDA(ctx, cudaMemcpyAsync(desc + (size_t)i * max_kp * 32, d_desc + (size_t)i * max_kp * 32, k * 32,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    }
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPIVO_OK;
}

int epivo_score_sampson(epivo_ctx* ctx, const double* E, int m, const float* p0, const float* p1, int n,
                        const double K[9], double threshold, int32_t* counts, float* medians, int* best,
                        uint8_t* best_mask) {
    if (!ctx) return EPIVO_ERR_INVALID;
    if (m < 0 || n < 0 || (m > 0 && !E) || (n > 0 && (!p0 || !p1)) || !K || !counts)
        EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (best) *best = -1;
    if (m == 0) return EPIVO_OK;
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const int stride = std::max(n, 1);
    size_t need = epv_align((size_t)m * 72) + epv_align((size_t)stride * 8) * 2 + epv_align((size_t)stride * 32) +
                  epv_align((size_t)m * 4) * 2 + epv_align(stride) + 4096 +
                  (medians ? epv_align((size_t)m * stride * 4) : 0);
    int rc = epv_ws_reserve(ctx, need);
    if (rc) return rc;
    double* d_E = epv_ws_take<double>(ctx, (size_t)m * 9);
    float* d_p0 = epv_ws_take<float>(ctx, (size_t)stride * 2);
    float* d_p1 = epv_ws_take<float>(ctx, (size_t)stride * 2);
    double* d_xn = epv_ws_take<double>(ctx, (size_t)stride * 4);
    int32_t* d_counts = epv_ws_take<int32_t>(ctx, m);
    float* d_med = epv_ws_take<float>(ctx, m);
    uint8_t* d_mask = epv_ws_take<uint8_t>(ctx, stride);
    int* d_best = epv_ws_take<int>(ctx, 1);
    float* d_err = medians ? epv_ws_take<float>(ctx, (size_t)m * stride) : nullptr;
    EPV_CUDA(ctx, cudaMemcpyAsync(d_E, E, (size_t)m * 72, cudaMemcpyHostToDevice, ctx->stream));
    if (n > 0) {
        EPV_CUDA(ctx, cudaMemcpyAsync(d_p0, p0, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        EPV_CUDA(ctx, cudaMemcpyAsync(d_p1, p1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = epv_normalize_launch(ctx, d_p0, d_p1, n, stride, K, d_xn);
    if (rc) return rc;
    rc = epv_score_launch(ctx, d_E, m, d_xn, stride, n, threshold / ((K[0] + K[4]) / 2.0), d_counts,
                          medians ? d_med : nullptr, d_err, d_best, d_mask);
    if (rc) return rc;
    int h_best = -1;
    EPV_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (medians) EPV_CUDA(ctx, cudaMemcpyAsync(medians, d_med, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(&h_best, d_best, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (best_mask && n > 0)
        EPV_CUDA(ctx, cudaMemcpyAsync(best_mask, d_mask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (best) *best = h_best;
    return EPIVO_OK;
}

// ---- P1 ------------------------------------------------------------------------------
int epivo_recover_pose(epivo_ctx* ctx, const double E[9], const float* p0, const float* p1, int n, const double K[9],
                       double dist_thresh, const uint8_t* in_mask, double R[9], double t[3], uint8_t* mask,
                       int* n_good) {
    if (!ctx) return EPIVO_ERR_INVALID;
    EpvRange nvtx_range(ctx, "epivo_recover_pose (recoverPose)");
    if (!E || !K || !R || !t || n < 0 || (n > 0 && (!p0 || !p1))) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const int stride = std::max(n, 1);
    int rc = epv_ws_reserve(ctx, (size_t)stride * (16 + 32 + 2) + 16 * 256 + 4096);
    if (rc) return rc;
    float* d_p0 = epv_ws_take<float>(ctx, (size_t)stride * 2);
    float* d_p1 = epv_ws_take<float>(ctx, (size_t)stride * 2);
    double* d_xn = epv_ws_take<double>(ctx, (size_t)stride * 4);
    uint8_t* d_mask = epv_ws_take<uint8_t>(ctx, stride);
    uint8_t* d_in = in_mask ? epv_ws_take<uint8_t>(ctx, stride) : nullptr;
    double* d_E = epv_ws_take<double>(ctx, 9);
    double* d_R = epv_ws_take<double>(ctx, 9);
    double* d_t = epv_ws_take<double>(ctx, 3);
    int32_t* d_n = epv_ws_take<int32_t>(ctx, 2);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_E, E, 72, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_n, &n, 4, cudaMemcpyHostToDevice, ctx->stream));
    if (n > 0) {
        EPV_CUDA(ctx, cudaMemcpyAsync(d_p0, p0, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        EPV_CUDA(ctx, cudaMemcpyAsync(d_p1, p1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (in_mask) EPV_CUDA(ctx, cudaMemcpyAsync(d_in, in_mask, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = epv_normalize_launch(ctx, d_p0, d_p1, n, stride, K, d_xn);
    if (rc) return rc;
    PosePlan pp{};
    pp.n_pairs = 1;
    pp.stride = stride;
    pp.E = d_E;
    pp.xn = d_xn;
    pp.n = d_n;
    pp.in_mask = d_in;
    pp.dist_thresh = dist_thresh;
    pp.R = d_R;
    pp.t = d_t;
    pp.mask = d_mask;
    pp.n_good = d_n + 1;
    rc = epv_pose_launch(ctx, pp);
    if (rc) return rc;
    int ng = 0;
    EPV_CUDA(ctx, cudaMemcpyAsync(R, d_R, 72, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(t, d_t, 24, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(&ng, d_n + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (mask && n > 0) EPV_CUDA(ctx, cudaMemcpyAsync(mask, d_mask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_good) *n_good = ng;
    return EPIVO_OK;
}

// ---- L4 ------------------------------------------------------------------------------
int epivo_lm_rt_batch(epivo_ctx* ctx, int B, int n_zeta, double epsilon, const int32_t* reps, const double* wreps,
                      int n_rep, double lambda0, int max_iters, double huber_delta, double* T0s, const double* pr,
                      const double* p_r, int N, epivo_lm_res* out, int32_t* iters_run) {
    if (!ctx) return EPIVO_ERR_INVALID;
    EpvRange nvtx_range(ctx, "epivo_lm_rt_batch (Levenberg_Marquardt)");
    if (B < 0 || !reps || !wreps || !T0s || !pr || !p_r || !out) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (n_zeta < 1 || n_rep < 1 || N < 1) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "n_zeta, n_rep and N must be >= 1");
    for (int j = 0; j < n_rep; ++j)                                  // sequence.hpp:118-122 asserts
        if (reps[2 * j] < 0 || reps[2 * j] >= n_zeta || reps[2 * j + 1] < 0 || reps[2 * j + 1] >= n_zeta)
            EPV_FAIL(ctx, EPIVO_ERR_INVALID, "rep %d = (%d,%d) outside [0,%d)", j, reps[2 * j], reps[2 * j + 1], n_zeta);
    if (B == 0) return EPIVO_OK;
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nT = (size_t)B * n_zeta * 16, nP = (size_t)B * n_rep * N * 3, nW = (size_t)B * n_rep;
    int rc = epv_ws_reserve(ctx, (nT + 2 * nP + nW) * 8 + (size_t)n_rep * 8 + (size_t)B * (24 + 4) + 8 * 256);
    if (rc) return rc;
    double* d_T = epv_ws_take<double>(ctx, nT);
    double* d_pr = epv_ws_take<double>(ctx, nP);
    double* d_p_r = epv_ws_take<double>(ctx, nP);
    double* d_w = epv_ws_take<double>(ctx, nW);
    int32_t* d_reps = epv_ws_take<int32_t>(ctx, (size_t)n_rep * 2);
    epivo_lm_res* d_out = epv_ws_take<epivo_lm_res>(ctx, B);
    int32_t* d_it = epv_ws_take<int32_t>(ctx, B);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_T, T0s, nT * 8, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_pr, pr, nP * 8, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_p_r, p_r, nP * 8, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_w, wreps, nW * 8, cudaMemcpyHostToDevice, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(d_reps, reps, (size_t)n_rep * 8, cudaMemcpyHostToDevice, ctx->stream));
    LmPlan lp{};
    lp.B = B;
    lp.n_zeta = n_zeta;
    lp.n_rep = n_rep;
    lp.N = N;
    lp.reps = d_reps;
    lp.wreps = d_w;
    lp.epsilon = epsilon;
    lp.lambda0 = lambda0;
    lp.huber_delta = huber_delta;
    lp.max_iters = max_iters;
    lp.T0s = d_T;
    lp.pr = d_pr;
    lp.p_r = d_p_r;
    lp.out = d_out;
    lp.iters = d_it;
    lp.active = nullptr;
    lp.single_pair = (n_zeta == 1 && n_rep == 1 && reps[0] == 0 && reps[1] == 0) ? 1 : 0;
    if (!ctx->ev_k0) {
        EPV_CUDA(ctx, cudaEventCreate(&ctx->ev_k0));
        EPV_CUDA(ctx, cudaEventCreate(&ctx->ev_k1));
    }
    EPV_CUDA(ctx, cudaEventRecord(ctx->ev_k0, ctx->stream));
    rc = epv_lm_launch(ctx, lp);
    if (rc) return rc;
    EPV_CUDA(ctx, cudaEventRecord(ctx->ev_k1, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(T0s, d_T, nT * 8, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)B * sizeof(epivo_lm_res), cudaMemcpyDeviceToHost, ctx->stream));
    if (iters_run) EPV_CUDA(ctx, cudaMemcpyAsync(iters_run, d_it, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPIVO_OK;
}

int epivo_lm_rt(epivo_ctx* ctx, int n_zeta, double epsilon, const int32_t* reps, const double* wreps, int n_rep,
                double lambda0, int max_iters, double huber_delta, double* T0s, const double* pr, const double* p_r,
                int N, epivo_lm_res* out, int* iters_run) {
    int32_t it = 0;
    int rc = epivo_lm_rt_batch(ctx, 1, n_zeta, epsilon, reps, wreps, n_rep, lambda0, max_iters, huber_delta, T0s, pr,
                               p_r, N, out, &it);
    if (iters_run) *iters_run = it;
    return rc;
}

}  // extern "C"
