// N2: the step right after the per-pair core in the reference drivers -- pose chaining and the
// per-inlier depth / point cloud (kitti_E.cpp:203-254, euroc_E.cpp:303-349):
//     dT      = [R | t/|t| * scale]            point transform of pair i, GT-scaled   (:218-223)
//     all_T_i = cT;  cT = cT * dT^-1           camera pose chain                      (:225-228)
//     per E-inlier j:  cp = K^-1 (u0,v0,1), cp' = K^-1 (u1,v1,1), P = [[1,0,-cp'_x],[0,1,-cp'_y]]
//         A = P t, B = P R cp;  if |B| > 1e-2:  d = |A|/|B|,  X = all_T_i[:3,:3] (d cp) + all_T_i[:3,3]   (:239-253)
//     limits_i = number of cloud points before pair i                                 (:238)
// The chain is a prefix product of 4x4 matrices, done here as a block scan (matrix product is
// associative; the result differs from the sequential loop only by rounding).  The cloud is
// written compacted in pair order straight to its final offset: a counting pass, a scan of the
// counts, and a writing pass.
#include <algorithm>

#include "common.cuh"
#include "stages.cuh"

namespace {

struct M34 { double m[12]; };    // rows of [R | t], last row 0 0 0 1 implied

__device__ __forceinline__ M34 m34_identity() {
    M34 o;
#pragma unroll
    for (int i = 0; i < 12; ++i) o.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
    return o;
}
__device__ __forceinline__ M34 m34_mul(const M34& a, const M34& b) {      // a * b
    M34 o;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o.m[i * 4 + j] = a.m[i * 4] * b.m[j] + a.m[i * 4 + 1] * b.m[4 + j] + a.m[i * 4 + 2] * b.m[8 + j] +
                             (j == 3 ? a.m[i * 4 + 3] : 0.0);
    }
    return o;
}
// general inverse of [A t; 0 1] (the reference calls MatrixXd::inverse())
__device__ __forceinline__ M34 m34_inv(const M34& a) {
    const double* m = a.m;
    const double c00 = m[5] * m[10] - m[6] * m[9], c01 = m[6] * m[8] - m[4] * m[10], c02 = m[4] * m[9] - m[5] * m[8];
    const double id = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02);
    M34 o;
    o.m[0] = c00 * id; o.m[1] = (m[2] * m[9] - m[1] * m[10]) * id; o.m[2] = (m[1] * m[6] - m[2] * m[5]) * id;
    o.m[4] = c01 * id; o.m[5] = (m[0] * m[10] - m[2] * m[8]) * id; o.m[6] = (m[2] * m[4] - m[0] * m[6]) * id;
    o.m[8] = c02 * id; o.m[9] = (m[1] * m[8] - m[0] * m[9]) * id;  o.m[10] = (m[0] * m[5] - m[1] * m[4]) * id;
#pragma unroll
    for (int i = 0; i < 3; ++i) o.m[i * 4 + 3] = -(o.m[i * 4] * m[3] + o.m[i * 4 + 1] * m[7] + o.m[i * 4 + 2] * m[11]);
    return o;
}

// dT of pair i from its refined pose and the caller's scale (kitti_E.cpp:220-223)
__device__ __forceinline__ M34 scaled_dT(const epivo_pair_result& r, double scale);
__device__ __forceinline__ M34 scaled_dT(const double* T, double scale) {
    M34 d;
    const double tx = T[3], ty = T[7], tz = T[11];
    const double nrm = sqrt(tx * tx + ty * ty + tz * tz);
    const double k = scale / nrm;                            // reference divides by the norm unguarded
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) d.m[i * 4 + j] = T[i * 4 + j];
    }
    d.m[3] = tx * k; d.m[7] = ty * k; d.m[11] = tz * k;
    return d;
}

__device__ __forceinline__ M34 scaled_dT(const epivo_pair_result& r, double scale) { return scaled_dT(r.T, scale); }

constexpr int CH_THREADS = 256;

// where the refined 4x4 of pair i lives: inside the result records of a run, or in a plain n x 16 array (the poses
// gathered from the ranks of a sharded sequence)
struct PoseSrc {
    const epivo_pair_result* res;
    const double* T;
    __device__ __forceinline__ const double* at(int i) const { return res ? res[i].T : T + (size_t)i * 16; }
};

// poses[i] = inv(dT_0) * ... * inv(dT_{i-1}), i = 0..n (n + 1 poses, poses[0] = I); one CTA
__global__ void __launch_bounds__(CH_THREADS) chain_kernel(PoseSrc src, const double* __restrict__ scales, int n,
                                                           double* __restrict__ poses) {
    __shared__ M34 s_tot[CH_THREADS];
    const int tid = threadIdx.x;
    const int per = (n + CH_THREADS - 1) / CH_THREADS;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    M34 acc = m34_identity();
    for (int i = lo; i < hi; ++i) acc = m34_mul(acc, m34_inv(scaled_dT(src.at(i), scales ? scales[i] : 1.0)));
    s_tot[tid] = acc;
    __syncthreads();
    for (int o = 1; o < CH_THREADS; o <<= 1) {               // inclusive scan of the per-thread products
        M34 v = s_tot[tid];
        if (tid >= o) v = m34_mul(s_tot[tid - o], v);
        __syncthreads();
        s_tot[tid] = v;
        __syncthreads();
    }
    M34 pre = tid > 0 ? s_tot[tid - 1] : m34_identity();     // product of everything before this thread's range
    for (int i = lo; i < hi; ++i) {
        double* out = poses + (size_t)i * 16;
#pragma unroll
        for (int k = 0; k < 12; ++k) out[k] = pre.m[k];
        out[12] = 0.0; out[13] = 0.0; out[14] = 0.0; out[15] = 1.0;
        pre = m34_mul(pre, m34_inv(scaled_dT(src.at(i), scales ? scales[i] : 1.0)));
    }
    if (tid == CH_THREADS - 1) {                             // the pose after the last pair
        double* out = poses + (size_t)n * 16;
#pragma unroll
        for (int k = 0; k < 12; ++k) out[k] = s_tot[tid].m[k];
        out[12] = 0.0; out[13] = 0.0; out[14] = 0.0; out[15] = 1.0;
    }
}

constexpr int CL_THREADS = 256;

// pass 0: counts[pair] = cloud points of the pair; pass 1: write them at offset limits[pair]
__global__ void __launch_bounds__(CL_THREADS) cloud_kernel(int pass, int n_pairs, int stride,
                                                           const epivo_pair_result* __restrict__ res,
                                                           const double* __restrict__ scales,
                                                           const double* __restrict__ poses,
                                                           const double* __restrict__ xin, const int32_t* __restrict__ n_inl,
                                                           int32_t* __restrict__ counts, const int64_t* __restrict__ limits,
                                                           double* __restrict__ points, int64_t cap) {
    __shared__ int s_warp[CL_THREADS / 32];
    __shared__ int s_base;
    const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (pair >= n_pairs) return;
    const M34 dT = scaled_dT(res[pair], scales ? scales[pair] : 1.0);
    const double* pT = poses + (size_t)pair * 16;
    const int n = n_inl[pair];
    const double* x = xin + (size_t)pair * 4 * stride;
    const int64_t base0 = pass ? limits[pair] : 0;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < n; start += CL_THREADS) {
        const int j = start + tid;
        bool keep = false;
        double X0 = 0, X1 = 0, X2 = 0;
        if (j < n) {
            const double c0 = x[j], c1 = x[stride + j], q0 = x[2 * stride + j], q1 = x[3 * stride + j];   // cp, cp'
            const double A0 = dT.m[3] - q0 * dT.m[11], A1 = dT.m[7] - q1 * dT.m[11];
            const double r0 = dT.m[0] * c0 + dT.m[1] * c1 + dT.m[2];
            const double r1 = dT.m[4] * c0 + dT.m[5] * c1 + dT.m[6];
            const double r2 = dT.m[8] * c0 + dT.m[9] * c1 + dT.m[10];
            const double B0 = r0 - q0 * r2, B1 = r1 - q1 * r2;
            const double nb = sqrt(B0 * B0 + B1 * B1);
            keep = nb > 1e-2;                                                                      // kitti_E.cpp:248
            if (keep && pass) {
                const double d = sqrt(A0 * A0 + A1 * A1) / nb;
                const double p0 = d * c0, p1 = d * c1, p2 = d;
                X0 = pT[0] * p0 + pT[1] * p1 + pT[2] * p2 + pT[3];
                X1 = pT[4] * p0 + pT[5] * p1 + pT[6] * p2 + pT[7];
                X2 = pT[8] * p0 + pT[9] * p1 + pT[10] * p2 + pT[11];
            }
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (keep && pass) {
            const int64_t k = base0 + off + __popc(bal & ((1u << lane) - 1));
            if (k < cap) {
                points[k * 3] = X0;
                points[k * 3 + 1] = X1;
                points[k * 3 + 2] = X2;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < CL_THREADS / 32; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0 && !pass) counts[pair] = s_base;
}

// limits[i] = sum of counts[0..i), limits[n] = total; one CTA
__global__ void __launch_bounds__(1024) limits_kernel(const int32_t* __restrict__ counts, int n, int64_t* __restrict__ limits) {
    __shared__ long long s_tot[1024];
    const int tid = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    long long acc = 0;
    for (int i = lo; i < hi; ++i) acc += counts[i];
    s_tot[tid] = acc;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        long long v = s_tot[tid];
        if (tid >= o) v += s_tot[tid - o];
        __syncthreads();
        s_tot[tid] = v;
        __syncthreads();
    }
    long long pre = tid > 0 ? s_tot[tid - 1] : 0;
    for (int i = lo; i < hi; ++i) {
        limits[i] = pre;
        pre += counts[i];
    }
    if (tid == 1023) limits[n] = s_tot[1023];
}

}  // namespace

int epv_chain_launch(epivo_ctx* ctx, const epivo_pair_result* d_res, const double* d_scales, int n, double* d_poses) {
    chain_kernel<<<1, CH_THREADS, 0, ctx->stream>>>(PoseSrc{d_res, nullptr}, d_scales, n, d_poses);
    EPV_LAUNCHED(ctx);
    return EPIVO_OK;
}

// D1 for a sharded sequence: the per-pair poses gathered from all ranks (host, n x 16) -> the chained camera poses
// (host, (n + 1) x 16), kitti_E.cpp:218-228.  Same kernel as epivo_seq_cloud's chain.
extern "C" int epivo_chain_poses(epivo_ctx* ctx, const double* T_pairs, const double* scales, int n, double* poses) {
    if (!ctx) return EPIVO_ERR_INVALID;
    if (n < 0 || !poses || (n > 0 && !T_pairs)) EPV_FAIL(ctx, EPIVO_ERR_INVALID, "null argument");
    if (n == 0) {
        for (int i = 0; i < 16; ++i) poses[i] = (i % 5 == 0) ? 1.0 : 0.0;
        return EPIVO_OK;
    }
    EPV_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t np = (size_t)n;
    int rc = epv_ws_reserve(ctx, epv_align(np * 128) + epv_align(np * 8) + epv_align((np + 1) * 128) + 4096);
    if (rc) return rc;
    double* d_T = epv_ws_take<double>(ctx, np * 16);
    double* d_scales = scales ? epv_ws_take<double>(ctx, np) : nullptr;
    double* d_poses = epv_ws_take<double>(ctx, (np + 1) * 16);
    EPV_CUDA(ctx, cudaMemcpyAsync(d_T, T_pairs, np * 128, cudaMemcpyHostToDevice, ctx->stream));
    if (scales) EPV_CUDA(ctx, cudaMemcpyAsync(d_scales, scales, np * 8, cudaMemcpyHostToDevice, ctx->stream));
    chain_kernel<<<1, CH_THREADS, 0, ctx->stream>>>(PoseSrc{nullptr, d_T}, d_scales, n, d_poses);
    EPV_LAUNCHED(ctx);
    EPV_CUDA(ctx, cudaMemcpyAsync(poses, d_poses, (np + 1) * 128, cudaMemcpyDeviceToHost, ctx->stream));
    EPV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPIVO_OK;
}

int epv_cloud_launch(epivo_ctx* ctx, int n_pairs, int stride, const epivo_pair_result* d_res, const double* d_scales,
                     const double* d_poses, const double* d_xin, const int32_t* d_ninl, int32_t* d_counts,
                     int64_t* d_limits, double* d_points, int64_t cap, int pass) {
    if (n_pairs <= 0) return EPIVO_OK;
    cloud_kernel<<<n_pairs, CL_THREADS, 0, ctx->stream>>>(pass, n_pairs, stride, d_res, d_scales, d_poses, d_xin, d_ninl,
                                                          d_counts, d_limits, d_points, cap);
    EPV_LAUNCHED(ctx);
    if (!pass) {
        limits_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, n_pairs, d_limits);
        EPV_LAUNCHED(ctx);
    }
    return EPIVO_OK;
}
