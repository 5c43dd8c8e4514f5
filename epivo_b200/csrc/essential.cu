// E1/E2: cv::findEssentialMat(RANSAC | LMEDS) on the GPU, one CTA per frame pair.
//
// Replaces kitti.cpp:98-104, kitti_E.cpp:98-104, euroc_E.cpp:202-208, kitti_ba.cpp:232,308,702.
// OpenCV's estimator (modules/calib3d/src/ptsetreg.cpp) is a sequential loop
//     sample 5 -> solve (<= 10 models) -> score every model -> keep strictly better -> shrink niters
// whose sample stream does not depend on the data (cv::RNG seeded with -1).  The kernel
// keeps those semantics exactly but evaluates CHUNK samples at a time: thread 0 draws the
// next CHUNK samples from the same RNG, CHUNK threads solve them (fivept.cuh), the warps
// score all their models (Sampson error in OpenCV's operation order, float32 compare), and
// thread 0 replays the sequential "strictly better / update niters" bookkeeping over the
// chunk in sample order, stopping where the sequential loop would have stopped.  The result
// is the model the sequential loop would return; at most CHUNK-1 samples are wasted.
#include <float.h>
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "fivept.cuh"
#include "stages.cuh"

namespace {

constexpr int ES_THREADS = 256;
constexpr int ES_WARPS = ES_THREADS / 32;
constexpr int ES_CHUNK = 32;

struct CvRng {   // cv::RNG: multiply-with-carry
    unsigned long long state;
    __device__ unsigned next() {
        state = (unsigned long long)(unsigned)state * 4164903690ULL + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    __device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a)) + a; }
};

// ptsetreg.cpp getSubset: 5 distinct indices, redraw on duplicates
__device__ __forceinline__ void draw_subset(CvRng& rng, int n, int* idx) {
    for (int i = 0; i < 5; ++i) {
        int v;
        bool dup;
        do {
            v = rng.uniform(0, n);
            dup = false;
            for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
        } while (dup);
        idx[i] = v;
    }
}

// ptsetreg.cpp RANSACUpdateNumIters
__device__ int update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmax(p, 0.0);
    p = fmin(p, 1.0);
    ep = fmax(ep, 0.0);
    ep = fmin(ep, 1.0);
    double num = fmax(1.0 - p, DBL_MIN);
    double denom = 1.0 - pow(1.0 - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// EMEstimatorCallback::computeError for one correspondence, OpenCV's operation order
// (Matx products accumulate left to right from 0; OpenCV's calib3d is built without FMA
// contraction, so every operation here is an explicit round-to-nearest intrinsic).
__device__ __forceinline__ float sampson_f32(const double* __restrict__ E, double a1, double b1, double a2,
                                             double b2) {
    // explicit round-to-nearest multiplies and adds: never contracted into FMAs
    const double ex0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], a1), __dmul_rn(E[1], b1)), E[2]);
    const double ex1 = __dadd_rn(__dadd_rn(__dmul_rn(E[3], a1), __dmul_rn(E[4], b1)), E[5]);
    const double ex2 = __dadd_rn(__dadd_rn(__dmul_rn(E[6], a1), __dmul_rn(E[7], b1)), E[8]);
    const double et0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], a2), __dmul_rn(E[3], b2)), E[6]);
    const double et1 = __dadd_rn(__dadd_rn(__dmul_rn(E[1], a2), __dmul_rn(E[4], b2)), E[7]);
    const double x2tEx1 = __dadd_rn(__dadd_rn(__dmul_rn(a2, ex0), __dmul_rn(b2, ex1)), ex2);
    const double den = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(ex0, ex0), __dmul_rn(ex1, ex1)), __dmul_rn(et0, et0)),
                                 __dmul_rn(et1, et1));
    return (float)__ddiv_rn(__dmul_rn(x2tEx1, x2tEx1), den);
}

// Inlier test  (float)(num/den) <= thr32  without the FP64 division.
// Let B be the largest double whose float rounding is <= thr32 (the midpoint between thr32 and the
// next float up, or its predecessor when the tie would round up).  Then
//     num/den <= B                =>  RN64(num/den) <= B        => inlier      (RN is monotone)
//     num/den >  B (1 + 2^-50)    =>  RN64(num/den) >  B        => outlier
// and the sign of B*den - num is exact in one FMA.  Only quotients within 2^-50 of B (or
// den == 0 / NaN) take the division; num and den are the same un-fused values OpenCV computes.
struct SampThr {
    double B, C;     // C = B * 2^-50
    float thr32;
};
__host__ __device__ inline SampThr make_samp_thr(float thr32) {
    SampThr t;
    t.thr32 = thr32;
    unsigned bits;
#ifdef __CUDA_ARCH__
    bits = __float_as_uint(thr32);
    const float up = __uint_as_float(bits + 1u);
#else
    memcpy(&bits, &thr32, 4);
    const unsigned ub = bits + 1u;
    float up;
    memcpy(&up, &ub, 4);
#endif
    double mid = 0.5 * ((double)thr32 + (double)up);          // exact: two adjacent floats
    if ((bits & 1u) && mid > 0.0 && mid < 1e300) {            // odd mantissa: the tie rounds away from thr32
#ifdef __CUDA_ARCH__
        mid = __longlong_as_double(__double_as_longlong(mid) - 1);
#else
        long long mb;
        memcpy(&mb, &mid, 8);
        mb -= 1;
        memcpy(&mid, &mb, 8);
#endif
    }
    if (!(thr32 >= 0.0f) || !(thr32 < 3e38f)) mid = -1.0;     // negative / NaN / huge: always take the exact path
    t.B = mid;
    t.C = mid * 8.881784197001252e-16;                         // 2^-50
    return t;
}

__device__ __forceinline__ bool sampson_inlier(const double* __restrict__ E, double a1, double b1, double a2,
                                               double b2, const SampThr& T) {
    const double ex0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], a1), __dmul_rn(E[1], b1)), E[2]);
    const double ex1 = __dadd_rn(__dadd_rn(__dmul_rn(E[3], a1), __dmul_rn(E[4], b1)), E[5]);
    const double ex2 = __dadd_rn(__dadd_rn(__dmul_rn(E[6], a1), __dmul_rn(E[7], b1)), E[8]);
    const double et0 = __dadd_rn(__dadd_rn(__dmul_rn(E[0], a2), __dmul_rn(E[3], b2)), E[6]);
    const double et1 = __dadd_rn(__dadd_rn(__dmul_rn(E[1], a2), __dmul_rn(E[4], b2)), E[7]);
    const double x2tEx1 = __dadd_rn(__dadd_rn(__dmul_rn(a2, ex0), __dmul_rn(b2, ex1)), ex2);
    const double den = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(ex0, ex0), __dmul_rn(ex1, ex1)), __dmul_rn(et0, et0)),
                                 __dmul_rn(et1, et1));
    const double num = __dmul_rn(x2tEx1, x2tEx1);
    const double r = __fma_rn(T.B, den, -num);
    if (r >= 0.0 && den > 0.0 && T.B > 0.0) return true;
    if (-r > __dmul_rn(T.C, den) && T.B > 0.0) return false;
    return (float)__ddiv_rn(num, den) <= T.thr32;
}

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xFFFFFFFFu, v); }

// k-th smallest (0-based) of n non-negative floats in buf, by bitwise binary search on the
// IEEE bit pattern (OpenCV sorts the float errors as ints: ptsetreg.cpp nth_element on int*).
__device__ float warp_select(const float* buf, int n, int k, int lane) {
    unsigned result = 0;
    for (int bit = 30; bit >= 0; --bit) {
        const unsigned cand = result | (1u << bit);
        int cnt = 0;
        for (int i = lane; i < n; i += 32) cnt += (__float_as_uint(buf[i]) < cand);
        cnt = warp_sum(cnt);
        if (cnt <= k) result = cand;
    }
    return __uint_as_float(result);
}

struct EssArgs {
    int n_pairs;
    int stride;                 // per-pair row stride of xn / masks
    const double* xn;           // [pair][4][stride] K-normalised x1 y1 x2 y2
    const int32_t* n;           // [pair] correspondences
    int method;
    double prob, thresh;        // thresh already divided by the focal length (OpenCV: threshold /= (fx+fy)/2)
    int max_iters;
    const int32_t* samples;     // optional injected samples [m][5] (shared by all pairs) or nullptr
    int m;
    float* errbuf;              // LMedS scratch [pair][ES_WARPS][stride]
    // pre-solved first samples (presolve_kernel): models of samples [0, pre_count) of every pair
    int pre_count;              // multiple of ES_CHUNK (0 = none)
    const double* pre_models;   // [pair][pre_count][10][9]
    const int32_t* pre_nmodels; // [pair][pre_count]
    const unsigned long long* pre_rng;   // [pair] RNG state after pre_count samples
    // outputs
    double* E;                  // [pair][9]
    uint8_t* mask;              // [pair][stride] {0,1}
    int32_t* n_inliers;         // [pair]
    int32_t* iters;             // [pair]
    int32_t* n_models;          // [pair]
    int32_t* status;            // [pair] 0 ok, EPIVO_ERR_NOMODEL
    double* xin;                // optional compacted inliers [pair][4][stride] (E3, kitti_E.cpp:106-112)
};

__global__ void __launch_bounds__(ES_THREADS, 2) essential_kernel(EssArgs a) {
    __shared__ double s_models[ES_CHUNK][10][9];
    __shared__ int s_nmodels[ES_CHUNK];
    __shared__ int s_idx[ES_CHUNK][5];
    __shared__ float s_score[ES_CHUNK][10];      // LMedS: median of every model of the chunk
    __shared__ int s_cnt[ES_WARPS][10];          // RANSAC: inlier counts of the sub-chunk being scored
    __shared__ double s_bestE[9];
    __shared__ int s_niters, s_iter, s_have, s_total_models;
    __shared__ unsigned long long s_rng;
    __shared__ int s_warpcnt[ES_WARPS];
    __shared__ int s_base;

    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n[pair];
    const int64_t so = (int64_t)pair * 4 * a.stride;
    const double* X1 = a.xn + so;
    const double* Y1 = X1 + a.stride;
    const double* X2 = Y1 + a.stride;
    const double* Y2 = X2 + a.stride;
    const bool lmeds = a.method == EPIVO_LMEDS;
    const float thr32 = (float)(a.thresh * a.thresh);

    if (tid == 0) {
        s_rng = a.pre_count > 0 && !a.samples ? a.pre_rng[pair] : 0xFFFFFFFFFFFFFFFFULL;   // RNG rng((uint64)-1)
        s_iter = 0;
        s_have = 0;
        s_total_models = 0;
        int ni = lmeds ? max(update_num_iters(a.prob, 0.45, 5, a.max_iters), 3) : max(a.max_iters, 1);
        if (a.samples) ni = min(ni, a.m);
        if (n < 5) ni = 0;
        if (n == 5) ni = 1;                                  // count == modelPoints: one solve on all points
        s_niters = ni;
    }
    __syncthreads();

    double best_score = lmeds ? DBL_MAX : 0.0;               // thread 0 only
    const SampThr thrR = make_samp_thr(thr32);

    while (true) {
        const int iter0 = s_iter, niters = s_niters;
        if (iter0 >= niters) break;
        // a chunk never straddles the end of the pre-solved samples (the in-kernel RNG state
        // continues from there)
        const bool presolved = iter0 < a.pre_count;
        const int ch = min(min(ES_CHUNK, niters - iter0), presolved ? a.pre_count - iter0 : ES_CHUNK);
        if (presolved) {
            // models of this chunk were solved by presolve_kernel at full occupancy
            const double* gm = a.pre_models + ((int64_t)pair * a.pre_count + iter0) * 90;
            const int32_t* gn = a.pre_nmodels + (int64_t)pair * a.pre_count + iter0;
            for (int i = tid; i < ch * 90; i += ES_THREADS) (&s_models[0][0][0])[i] = gm[i];
            for (int i = tid; i < ch; i += ES_THREADS) s_nmodels[i] = gn[i];
        } else if (tid == 0) {
            if (n == 5) {
                for (int k = 0; k < 5; ++k) s_idx[0][k] = k;
            } else if (a.samples) {
                for (int s = 0; s < ch; ++s)
                    for (int k = 0; k < 5; ++k) s_idx[s][k] = a.samples[(int64_t)(iter0 + s) * 5 + k];
            } else {
                CvRng rng{s_rng};
                for (int s = 0; s < ch; ++s) draw_subset(rng, n, s_idx[s]);
                s_rng = rng.state;
            }
        }
        __syncthreads();
        if (!presolved) {
            if (tid < ch) {
                double x1[5][2], x2[5][2];
                for (int k = 0; k < 5; ++k) {
                    const int i = s_idx[tid][k];
                    x1[k][0] = X1[i]; x1[k][1] = Y1[i];
                    x2[k][0] = X2[i]; x2[k][1] = Y2[i];
                }
                s_nmodels[tid] = fivept::solve(x1, x2, s_models[tid]);
            }
            __syncthreads();
        }
        // score + replay in sub-chunks of at most ES_WARPS samples.  RANSAC usually shrinks niters
        // after the first few samples: samples at or beyond the current niters are never scored,
        // and when fewer than ES_WARPS samples remain their points are split across the idle warps.
        for (int sub = 0; sub * ES_WARPS < ch; ++sub) {
            const int sbase = sub * ES_WARPS;
            const int r = min(ES_WARPS, min(ch, s_niters - iter0) - sbase);     // samples scored now (>= 1)
            if (!lmeds) {
                if (tid < ES_WARPS * 10) (&s_cnt[0][0])[tid] = 0;
                __syncthreads();
                const int slices = (r >= 5) ? 1 : (r >= 3 ? 2 : (r == 2 ? 4 : 8));   // warps per sample
                const int q = warp / slices, slice = warp % slices;
                if (q < r && n != 5) {
                    const int s = sbase + q;
                    const int nm = s_nmodels[s];
                    // point-outer loop: a correspondence is loaded once and scored against every
                    // model of this warp's sample (models broadcast from shared memory)
                    int cnt[10];
#pragma unroll
                    for (int k = 0; k < 10; ++k) cnt[k] = 0;
                    for (int i = lane + 32 * slice; i < n; i += 32 * slices) {
                        const double a1 = X1[i], b1 = Y1[i], a2 = X2[i], b2 = Y2[i];
#pragma unroll
                        for (int k = 0; k < 10; ++k)
                            if (k < nm) cnt[k] += sampson_inlier(s_models[s][k], a1, b1, a2, b2, thrR) ? 1 : 0;
                    }
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
                        if (k < nm) {
                            const int c = warp_sum(cnt[k]);
                            if (lane == 0) atomicAdd(&s_cnt[q][k], c);
                        }
                    }
                }
            } else {
                const int s = sbase + warp;
                if (warp < r && n != 5) {
                    const int nm = s_nmodels[s];
                    for (int k = 0; k < nm; ++k) {
                        const double* E = s_models[s][k];
                        float* buf = a.errbuf + ((int64_t)pair * ES_WARPS + warp) * a.stride;
                        for (int i = lane; i < n; i += 32) buf[i] = sampson_f32(E, X1[i], Y1[i], X2[i], Y2[i]);
                        __syncwarp();
                        const float med = warp_select(buf, n, n / 2, lane);
                        __syncwarp();
                        if (lane == 0) s_score[s][k] = med;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {                                  // sequential bookkeeping of ptsetreg.cpp run()
                int ni = s_niters;
                int q = sbase;
                const int qend = sbase + r;
                for (; q < qend; ++q) {
                    if (iter0 + q >= ni) break;
                    for (int k = 0; k < s_nmodels[q]; ++k) {
                        s_total_models++;
                        if (n == 5) {                        // minimal case: first solution, all inliers
                            if (!s_have) {
                                s_have = 1;
                                for (int c = 0; c < 9; ++c) s_bestE[c] = s_models[q][k][c];
                            }
                            continue;
                        }
                        if (!lmeds) {
                            const int good = s_cnt[q - sbase][k];
                            if (good > max((int)best_score, 4)) {
                                best_score = good;
                                s_have = 1;
                                for (int c = 0; c < 9; ++c) s_bestE[c] = s_models[q][k][c];
                                ni = update_num_iters(a.prob, (double)(n - good) / n, 5, ni);
                            }
                        } else {
                            const double med = (double)s_score[q][k];
                            if (med < best_score) {
                                best_score = med;
                                s_have = 1;
                                for (int c = 0; c < 9; ++c) s_bestE[c] = s_models[q][k][c];
                            }
                        }
                    }
                }
                s_iter = iter0 + q;
                s_niters = ni;
            }
            __syncthreads();
            if (s_iter >= s_niters) break;                   // the sequential loop would have stopped here
        }
        __syncthreads();
    }

    // final mask (ptsetreg.cpp findInliers on the best model), inlier compaction
    __shared__ float s_thr;
    if (tid == 0) {
        float t = thr32;
        if (lmeds && s_have && n > 5) {
            double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 5)) * sqrt(best_score);
            sigma = fmax(sigma, 0.001);
            t = (float)(sigma * sigma);
        }
        s_thr = t;
        s_base = 0;
    }
    __syncthreads();
    const bool have = s_have != 0;
    const SampThr thrF = make_samp_thr(s_thr);
    uint8_t* mask = a.mask + (int64_t)pair * a.stride;
    double* xin = a.xin ? a.xin + so : nullptr;
    for (int start = 0; start < n; start += ES_THREADS) {
        const int i = start + tid;
        bool in = false;
        double p[4] = {0, 0, 0, 0};
        if (i < n && have) {
            p[0] = X1[i]; p[1] = Y1[i]; p[2] = X2[i]; p[3] = Y2[i];
            in = (n == 5) ? true : sampson_inlier(s_bestE, p[0], p[1], p[2], p[3], thrF);
        }
        if (i < n) mask[i] = in ? 1 : 0;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, in);
        if (lane == 0) s_warpcnt[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warpcnt[w];
        if (in && xin) {
            const int k = off + __popc(bal & ((1u << lane) - 1));
            xin[k] = p[0];
            xin[a.stride + k] = p[1];
            xin[2 * a.stride + k] = p[2];
            xin[3 * a.stride + k] = p[3];
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < ES_WARPS; ++w) tot += s_warpcnt[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) {
        a.n_inliers[pair] = s_base;
        a.iters[pair] = s_iter;
        a.n_models[pair] = s_total_models;
        a.status[pair] = have ? 0 : EPIVO_ERR_NOMODEL;
    }
    if (tid < 9) a.E[(int64_t)pair * 9 + tid] = have ? s_bestE[tid] : 0.0;
}

// ---- first samples of every pair, solved at full occupancy --------------------------------
// sample_kernel: one thread per pair draws the first `count` samples of OpenCV's stream;
// presolve_kernel: one thread per (pair, sample) runs the 5-point solver.
__global__ void sample_kernel(int n_pairs, const int32_t* __restrict__ n_arr, int count, int32_t* __restrict__ idx,
                              unsigned long long* __restrict__ rng_out) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    const int n = n_arr[pair];
    CvRng rng{0xFFFFFFFFFFFFFFFFULL};
    int32_t* out = idx + (int64_t)pair * count * 5;
    if (n > 5) {
        for (int s = 0; s < count; ++s) {
            int v[5];
            draw_subset(rng, n, v);
            for (int k = 0; k < 5; ++k) out[s * 5 + k] = v[k];
        }
    } else {
        for (int s = 0; s < count; ++s)
            for (int k = 0; k < 5; ++k) out[s * 5 + k] = (n == 5) ? k : -1;
    }
    rng_out[pair] = rng.state;
}

#ifndef EPV_PRESOLVE_MINBLOCKS
#define EPV_PRESOLVE_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(64, EPV_PRESOLVE_MINBLOCKS)
presolve_kernel(int n_pairs, int stride, const double* __restrict__ xn, const int32_t* __restrict__ n_arr, int count,
                int used, const int32_t* __restrict__ idx, const int32_t* __restrict__ shared_samples,
                double* __restrict__ models, int32_t* __restrict__ nmodels) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (int64_t)n_pairs * count) return;
    const int pair = (int)(g / count), s = (int)(g % count);
    const int n = n_arr[pair];
    int32_t* nm = nmodels + g;
    if (n < 5 || s >= used || (n == 5 && s > 0)) { *nm = 0; return; }
    const int32_t* id = shared_samples ? shared_samples + (int64_t)s * 5 : idx + g * 5;
    const double* X1 = xn + (int64_t)pair * 4 * stride;
    double x1[5][2], x2[5][2];
    for (int k = 0; k < 5; ++k) {
        const int i = (n == 5) ? k : id[k];
        x1[k][0] = X1[i]; x1[k][1] = X1[stride + i];
        x2[k][0] = X1[2 * stride + i]; x2[k][1] = X1[3 * stride + i];
    }
    double E[10][9];
    const int c = fivept::solve(x1, x2, E);
    *nm = c;
    double* out = models + g * 90;
    for (int k = 0; k < c; ++k)
        for (int q = 0; q < 9; ++q) out[k * 9 + q] = E[k][q];
}

// ---- K2 alone: one thread per sample ------------------------------------------------
__global__ void __launch_bounds__(64) five_point_kernel(const double* __restrict__ x1, const double* __restrict__ x2,
                                                        int m, double* __restrict__ Eout, int32_t* __restrict__ nm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double a[5][2], b[5][2];
    for (int k = 0; k < 5; ++k) {
        a[k][0] = x1[(i * 5 + k) * 2];
        a[k][1] = x1[(i * 5 + k) * 2 + 1];
        b[k][0] = x2[(i * 5 + k) * 2];
        b[k][1] = x2[(i * 5 + k) * 2 + 1];
    }
    double E[10][9];
    const int n = fivept::solve(a, b, E);
    nm[i] = n;
    for (int k = 0; k < n; ++k)
        for (int q = 0; q < 9; ++q) Eout[((int64_t)i * 10 + k) * 9 + q] = E[k][q];
}

// ---- K3 alone: fixed hypothesis set, m models x n correspondences ----------------------
// Every thread keeps one correspondence in registers; the CTA walks a block of models staged
// in shared memory (broadcast reads); votes are counted per warp with ballot + popc, per CTA
// in shared memory and per model with one global atomicAdd per (CTA, model).
constexpr int SC_THREADS = 256;
constexpr int SC_MODELS = 128;

__global__ void __launch_bounds__(SC_THREADS)
score_count_kernel(const double* __restrict__ E, int m, const double* __restrict__ xn, int stride, int n,
                   float thr32, int32_t* __restrict__ counts) {
    __shared__ double s_E[SC_MODELS][9];
    __shared__ int s_cnt[SC_MODELS];
    const int tid = threadIdx.x, lane = tid & 31;
    const int m0 = blockIdx.y * SC_MODELS;
    const int mm = min(SC_MODELS, m - m0);
    for (int i = tid; i < mm * 9; i += SC_THREADS) (&s_E[0][0])[i] = E[(int64_t)m0 * 9 + i];
    for (int i = tid; i < mm; i += SC_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const int i = blockIdx.x * SC_THREADS + tid;
    const bool valid = i < n;
    const int ii = valid ? i : 0;
    const double a1 = xn[ii], b1 = xn[stride + ii], a2 = xn[2 * stride + ii], b2 = xn[3 * stride + ii];
    const SampThr T = make_samp_thr(thr32);
    for (int k = 0; k < mm; ++k) {
        const bool in = valid && sampson_inlier(s_E[k], a1, b1, a2, b2, T);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, in);
        if (lane == 0 && bal) atomicAdd(&s_cnt[k], __popc(bal));
    }
    __syncthreads();
    for (int k = tid; k < mm; k += SC_THREADS)
        if (s_cnt[k]) atomicAdd(&counts[m0 + k], s_cnt[k]);
}

// LMedS medians of a fixed hypothesis set: one warp per model
__global__ void __launch_bounds__(256)
score_median_kernel(const double* __restrict__ E, int m, const double* __restrict__ xn, int stride, int n,
                    float* __restrict__ errbuf, float* __restrict__ medians) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= m) return;
    double e[9];
    for (int q = 0; q < 9; ++q) e[q] = E[(int64_t)warp * 9 + q];
    float* buf = errbuf + (int64_t)warp * stride;
    for (int i = lane; i < n; i += 32)
        buf[i] = sampson_f32(e, xn[i], xn[stride + i], xn[2 * stride + i], xn[3 * stride + i]);
    __syncwarp();
    const float med = warp_select(buf, n, n / 2, lane);
    if (lane == 0) medians[warp] = med;
}

// first model with the largest count (block-level argmax, lowest index wins ties)
__global__ void __launch_bounds__(256) argmax_first_kernel(const int32_t* __restrict__ counts, int m, int* best) {
    __shared__ unsigned long long s_key[8];
    unsigned long long key = 0;
    for (int i = threadIdx.x; i < m; i += 256) {
        // larger count wins; among equals the smaller index wins
        unsigned long long k = ((unsigned long long)(unsigned)counts[i] << 32) | (unsigned)(0x7FFFFFFF - i);
        key = k > key ? k : key;
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) key = s_key[w] > key ? s_key[w] : key;
        *best = m > 0 ? 0x7FFFFFFF - (int)(key & 0xFFFFFFFFu) : -1;
    }
}

__global__ void mask_of_model_kernel(const double* __restrict__ E, const int* __restrict__ best,
                                     const double* __restrict__ xn, int stride, int n, float thr32,
                                     uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = *best;
    if (b < 0) { mask[i] = 0; return; }
    double e[9];
    for (int q = 0; q < 9; ++q) e[q] = E[(int64_t)b * 9 + q];
    mask[i] = sampson_inlier(e, xn[i], xn[stride + i], xn[2 * stride + i], xn[3 * stride + i], make_samp_thr(thr32)) ? 1 : 0;
}

// pixel -> K-normalised float64 (five-point.cpp: one scale-and-shift per coordinate)
__global__ void normalize_kernel(const float* __restrict__ p0, const float* __restrict__ p1, int n, int stride,
                                 double ax, double bx, double ay, double by, double* __restrict__ xn) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xn[i] = __fma_rn((double)p0[2 * i], ax, bx);
    xn[stride + i] = __fma_rn((double)p0[2 * i + 1], ay, by);
    xn[2 * stride + i] = __fma_rn((double)p1[2 * i], ax, bx);
    xn[3 * stride + i] = __fma_rn((double)p1[2 * i + 1], ay, by);
}

}  // namespace

int epv_normalize_launch(epivo_ctx* ctx, const float* d_p0, const float* d_p1, int n, int stride, const double K[9],
                         double* d_xn) {
    if (n <= 0) return EPIVO_OK;
    const double ax = 1.0 / K[0], ay = 1.0 / K[4];
    normalize_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_p0, d_p1, n, stride, ax, -K[2] * ax, ay,
                                                               -K[5] * ay, d_xn);
    EPV_LAUNCHED(ctx);
    return EPIVO_OK;
}

int epv_essential_pre_count(int method, double prob, int max_iters, int m_samples) {
    // how many leading samples are solved ahead of the per-pair kernel: two sub-chunks for RANSAC
    // (on KITTI-like data it stops after ~10 samples; a pair that needs more continues with
    // in-kernel solves), every sample for LMedS (its iteration count is fixed)
    int want = 2 * ES_WARPS;
    if (method == EPIVO_LMEDS) {
        double num = log(fmax(1.0 - prob, DBL_MIN)), den = log(1.0 - pow(1.0 - 0.45, 5.0));
        int ni = (int)rint(num / den);
        ni = std::max(std::min(ni, max_iters), 3);
        want = ni;
    }
    if (m_samples > 0) want = std::min(want, m_samples);
    want = std::min(want, std::max(max_iters, 1));
    int c = (want + ES_WARPS - 1) / ES_WARPS * ES_WARPS;
    return std::min(c, 4 * ES_CHUNK);
}

int epv_essential_launch(epivo_ctx* ctx, const EssentialPlan& p) {
    if (p.n_pairs <= 0) return EPIVO_OK;
    EssArgs a{};
    a.n_pairs = p.n_pairs;
    a.stride = p.stride;
    a.xn = p.xn;
    a.n = p.n;
    a.method = p.method;
    a.prob = p.prob;
    a.thresh = p.thresh;
    a.max_iters = p.max_iters;
    a.samples = p.samples;
    a.m = p.m;
    a.errbuf = p.errbuf;
    a.E = p.E;
    a.mask = p.mask;
    a.n_inliers = p.n_inliers;
    a.iters = p.iters;
    a.n_models = p.n_models;
    a.status = p.status;
    a.xin = p.xin;
    a.pre_count = 0;
    if (p.pre_count > 0 && p.pre_models && p.pre_nmodels && p.pre_idx && p.pre_rng) {
        const int count = p.pre_count;
        int used = count;
        if (p.samples) used = std::min(count, p.m);
        if (!p.samples) {
            sample_kernel<<<(p.n_pairs + 127) / 128, 128, 0, ctx->stream>>>(p.n_pairs, p.n, count, p.pre_idx, p.pre_rng);
            EPV_LAUNCHED(ctx);
        }
        const int64_t total = (int64_t)p.n_pairs * count;
        presolve_kernel<<<(unsigned)((total + 63) / 64), 64, 0, ctx->stream>>>(
            p.n_pairs, p.stride, p.xn, p.n, count, used, p.pre_idx, p.samples, p.pre_models, p.pre_nmodels);
        EPV_LAUNCHED(ctx);
        a.pre_count = count;
        a.pre_models = p.pre_models;
        a.pre_nmodels = p.pre_nmodels;
        a.pre_rng = p.pre_rng;
    }
    if (p.ev_presolved) EPV_CUDA(ctx, cudaEventRecord(p.ev_presolved, ctx->stream));
    essential_kernel<<<p.n_pairs, ES_THREADS, 0, ctx->stream>>>(a);
    EPV_LAUNCHED(ctx);
    return EPIVO_OK;
}

size_t epv_essential_errbuf_floats(int n_pairs, int stride) { return (size_t)n_pairs * ES_WARPS * stride; }

int epv_five_point_launch(epivo_ctx* ctx, const double* d_x1, const double* d_x2, int m, double* d_E, int32_t* d_nm) {
    if (m <= 0) return EPIVO_OK;
    five_point_kernel<<<(m + 63) / 64, 64, 0, ctx->stream>>>(d_x1, d_x2, m, d_E, d_nm);
    EPV_LAUNCHED(ctx);
    return EPIVO_OK;
}

int epv_score_launch(epivo_ctx* ctx, const double* d_E, int m, const double* d_xn, int stride, int n, double thresh,
                     int32_t* d_counts, float* d_medians, float* d_errbuf, int* d_best, uint8_t* d_mask) {
    const float thr32 = (float)(thresh * thresh);
    EPV_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (size_t)std::max(m, 1) * 4, ctx->stream));
    if (m > 0 && n > 0) {
        dim3 grid((n + SC_THREADS - 1) / SC_THREADS, (m + SC_MODELS - 1) / SC_MODELS);
        score_count_kernel<<<grid, SC_THREADS, 0, ctx->stream>>>(d_E, m, d_xn, stride, n, thr32, d_counts);
        EPV_LAUNCHED(ctx);
        if (d_medians) {
            score_median_kernel<<<(m * 32 + 255) / 256, 256, 0, ctx->stream>>>(d_E, m, d_xn, stride, n, d_errbuf,
                                                                             d_medians);
            EPV_LAUNCHED(ctx);
        }
    }
    argmax_first_kernel<<<1, 256, 0, ctx->stream>>>(d_counts, m, d_best);
    EPV_LAUNCHED(ctx);
    if (n > 0) {
        mask_of_model_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_E, d_best, d_xn, stride, n, thr32, d_mask);
        EPV_LAUNCHED(ctx);
    }
    return EPIVO_OK;
}

#ifdef EPV_PROFILE_SOLVE
// tuning builds only: read and reset the per-phase clock totals of fivept::solve
extern "C" int epivo_debug_solve_profile(unsigned long long* out8) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(out8, fivept::g_solve_prof, sizeof(z)) != cudaSuccess) return -2;
    if (cudaMemcpyToSymbol(fivept::g_solve_prof, z, sizeof(z)) != cudaSuccess) return -2;
    return 0;
}
#endif
