// N4 front end, detector part: cv::FastFeatureDetector (FAST-9/16, OpenCV features2d/src/fast.cpp FAST_t<16>)
// for a batch of 8-bit images, as the reference calls it:
//   FastFeatureDetector::create(40)->detect(src, kp0, Mat())   kitti_E.cpp:71-74, kitti_ba.cpp:49,62
//   FastFeatureDetector::create()  ->detect(...)  (threshold 10) kitti_ba.cpp:98,117-118
// Integer / byte work bound by HBM (one read of the image, one 2-byte score write and read per pixel): the
// result -- keypoint coordinates, their order and the response -- is bit-exact with OpenCV.
//
// OpenCV's loop, restated:  a pixel (x, y) with 3 <= x < cols-3, 3 <= y < rows-3 is a corner when 9 contiguous
// pixels of its 16-pixel Bresenham circle are all darker than v - t or all brighter than v + t (the pair-wise
// rejection cascade in front of that test never rejects a corner: 9 contiguous of 16 contain one pixel of every
// antipodal pair).  With non-maximum suppression a corner is kept when its score -- cornerScore<16>: the largest
// threshold for which it is still a corner -- is strictly greater than the scores of its 8 neighbours (0 where the
// neighbour is not a corner); the response is the score, or 0 without suppression.  Keypoints come out row by row,
// left to right.
#include "stages.cuh"

namespace {

constexpr int FT_BX = 32, FT_BY = 8;                     // pixels per CTA of the score pass
constexpr int FT_TW = FT_BX + 6, FT_TH = FT_BY + 6;      // tile with the radius-3 halo

// circle offsets (dx, dy) in OpenCV's order (makeOffsets, patternSize 16)
__constant__ int8_t c_circle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                       {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// 9 contiguous set bits in a circular 16-bit mask
__device__ __forceinline__ bool arc9(unsigned m) {
    const unsigned mm = m | (m << 16);
    unsigned r = mm & (mm >> 1);          // runs of 2
    r &= r >> 2;                          // 4
    r &= r >> 4;                          // 8
    r &= mm >> 8;                         // 9
    return (r & 0xFFFFu) != 0;
}

// cornerScore<16> (fast_score.cpp): d[k] = v - p[k];
//   a0 = max(threshold, max over the 16 arcs of 9 of min d);  b0 = min(-a0, min over the arcs of max d);  -b0 - 1
__device__ __forceinline__ int corner_score(const int (&d)[16], int threshold) {
    int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo2[i] = min(d[i], d[(i + 1) & 15]); hi2[i] = max(d[i], d[(i + 1) & 15]); }
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo4[i] = min(lo2[i], lo2[(i + 2) & 15]); hi4[i] = max(hi2[i], hi2[(i + 2) & 15]); }
    int a0 = threshold;
#pragma unroll
    for (int i = 0; i < 16; ++i) a0 = max(a0, min(min(lo4[i], lo4[(i + 4) & 15]), d[(i + 8) & 15]));
    int b0 = -a0;
#pragma unroll
    for (int i = 0; i < 16; ++i) b0 = min(b0, max(max(hi4[i], hi4[(i + 4) & 15]), d[(i + 8) & 15]));
    return -b0 - 1;
}

// score pass: S[img][y][x] = 0x100 | score for corners, 0 elsewhere
__global__ void __launch_bounds__(FT_BX * FT_BY) fast_score_kernel(const uint8_t* __restrict__ img, int rows, int cols,
                                                                   int threshold, int want_score, uint16_t* __restrict__ S) {
    __shared__ uint8_t tile[FT_TH][FT_TW + 2];
    const int x0 = blockIdx.x * FT_BX, y0 = blockIdx.y * FT_BY;
    const uint8_t* im = img + (size_t)blockIdx.z * rows * cols;
    uint16_t* s = S + (size_t)blockIdx.z * rows * cols;
    const int tid = threadIdx.y * FT_BX + threadIdx.x;
    for (int i = tid; i < FT_TH * FT_TW; i += FT_BX * FT_BY) {
        const int ty = i / FT_TW, tx = i % FT_TW;
        const int gx = min(max(x0 + tx - 3, 0), cols - 1), gy = min(max(y0 + ty - 3, 0), rows - 1);
        tile[ty][tx] = im[(size_t)gy * cols + gx];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= cols || y >= rows) return;
    uint16_t out = 0;
    if (x >= 3 && x < cols - 3 && y >= 3 && y < rows - 3) {
        const int cx = threadIdx.x + 3, cy = threadIdx.y + 3;
        const int v = tile[cy][cx];
        int d[16];
        unsigned dark = 0, bright = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int p = tile[cy + c_circle[k][1]][cx + c_circle[k][0]];
            d[k] = v - p;
            dark |= (d[k] > threshold ? 1u : 0u) << k;        // p < v - t
            bright |= (-d[k] > threshold ? 1u : 0u) << k;     // p > v + t
        }
        if (arc9(dark) || arc9(bright)) out = (uint16_t)(0x100 | (want_score ? corner_score(d, threshold) : 0));
    }
    s[(size_t)y * cols + x] = out;
}

__device__ __forceinline__ bool fast_keep(const uint16_t* __restrict__ s, int rows, int cols, int x, int y, int nonmax) {
    const int c = s[(size_t)y * cols + x];
    if (!(c & 0x100)) return false;
    if (!nonmax) return true;
    const int sc = c & 0xFF;
    // corners exist only for 3 <= x < cols-3, 3 <= y < rows-3: the 8 neighbours are inside the image
    const uint16_t* r0 = s + (size_t)(y - 1) * cols + x;
    const uint16_t* r1 = s + (size_t)y * cols + x;
    const uint16_t* r2 = s + (size_t)(y + 1) * cols + x;
    return sc > (r0[-1] & 0xFF) && sc > (r0[0] & 0xFF) && sc > (r0[1] & 0xFF) && sc > (r1[-1] & 0xFF) &&
           sc > (r1[1] & 0xFF) && sc > (r2[-1] & 0xFF) && sc > (r2[0] & 0xFF) && sc > (r2[1] & 0xFF);
}

// one warp per (row, image): number of kept corners in the row
__global__ void __launch_bounds__(128) fast_count_kernel(const uint16_t* __restrict__ S, int rows, int cols, int nonmax,
                                                         int32_t* __restrict__ rowcount) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const uint16_t* s = S + (size_t)blockIdx.y * rows * cols;
    int n = 0;
    if (row >= 3 && row < rows - 3)
        for (int x = lane; x < cols; x += 32) n += fast_keep(s, rows, cols, x, row, nonmax) ? 1 : 0;
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    if (lane == 0) rowcount[(size_t)blockIdx.y * rows + row] = n;
}

// one CTA per image: exclusive scan of the row counts in place, total to counts[img]
__global__ void __launch_bounds__(256) fast_scan_kernel(int32_t* __restrict__ rowcount, int rows, int32_t* __restrict__ counts) {
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    int32_t* rc = rowcount + (size_t)blockIdx.x * rows;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < rows; base += 256) {
        const int i = base + tid;
        const int v = i < rows ? rc[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int off = s_carry;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (i < rows) rc[i] = off + incl - v;
        __syncthreads();
        if (tid == 255) s_carry = off + incl;
        __syncthreads();
    }
    if (tid == 0) counts[blockIdx.x] = s_carry;
}

// one warp per (row, image): write the kept corners of the row, left to right, at the row's offset
__global__ void __launch_bounds__(128) fast_write_kernel(const uint16_t* __restrict__ S, int rows, int cols, int nonmax,
                                                         const int32_t* __restrict__ rowoff, int max_kp,
                                                         float* __restrict__ kps, float* __restrict__ resp) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row < 3 || row >= rows - 3) return;
    const uint16_t* s = S + (size_t)blockIdx.y * rows * cols;
    int pos = rowoff[(size_t)blockIdx.y * rows + row];
    float* kp = kps + (size_t)blockIdx.y * max_kp * 2;
    float* rs = resp ? resp + (size_t)blockIdx.y * max_kp : nullptr;
    for (int x0 = 0; x0 < cols; x0 += 32) {
        const int x = x0 + lane;
        const bool keep = x < cols && fast_keep(s, rows, cols, x, row, nonmax);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) {
            const int k = pos + __popc(bal & ((1u << lane) - 1));
            if (k < max_kp) {
                kp[2 * k] = (float)x;
                kp[2 * k + 1] = (float)row;
                if (rs) rs[k] = nonmax ? (float)(s[(size_t)row * cols + x] & 0xFF) : 0.0f;
            }
        }
        pos += __popc(bal);
    }
}

}  // namespace

size_t epv_fast_work_bytes(int n_images, int rows, int cols) {
    return (size_t)n_images * rows * cols * 2 + (size_t)n_images * rows * 4 + 256;
}

// d_img: [n_images][rows][cols] bytes; d_kps: [n_images][max_kp][2]; d_resp: optional [n_images][max_kp];
// d_counts: [n_images] (the number FOUND, which may exceed max_kp: only the first max_kp are stored)
int epv_fast_launch(epivo_ctx* ctx, const uint8_t* d_img, int n_images, int rows, int cols, int threshold, int nonmax,
                    int max_kp, float* d_kps, float* d_resp, int32_t* d_counts, void* work) {
    if (n_images <= 0) return EPIVO_OK;
    uint16_t* S = (uint16_t*)work;
    int32_t* rowcount = (int32_t*)((uint8_t*)work + (((size_t)n_images * rows * cols * 2 + 127) & ~(size_t)127));
    if (n_images > 65535) EPV_FAIL(ctx, EPIVO_ERR_UNSUPPORTED, "more than 65535 images per call");
    const dim3 g((cols + FT_BX - 1) / FT_BX, (rows + FT_BY - 1) / FT_BY, n_images);
    fast_score_kernel<<<g, dim3(FT_BX, FT_BY), 0, ctx->stream>>>(d_img, rows, cols, threshold, nonmax, S);
    EPV_LAUNCHED(ctx);
    const dim3 gr((rows + 3) / 4, n_images);
    fast_count_kernel<<<gr, 128, 0, ctx->stream>>>(S, rows, cols, nonmax, rowcount);
    EPV_LAUNCHED(ctx);
    fast_scan_kernel<<<n_images, 256, 0, ctx->stream>>>(rowcount, rows, d_counts);
    EPV_LAUNCHED(ctx);
    fast_write_kernel<<<gr, 128, 0, ctx->stream>>>(S, rows, cols, nonmax, rowcount, max_kp, d_kps, d_resp);
    EPV_LAUNCHED(ctx);
    return EPIVO_OK;
}
