// Evaluation metrics of the reference run (mcmc_utils.py:28-85 stein_disc, :88-111 max_mean_disc): O(T^2 d) pairwise
// kernels over the flow / resampled samples (exe_flow_matching.py:463-488).
//
// One thread per pair (i, j) in 16 x 16 pair tiles; the coordinates (and scores) of the 16 + 16 rows are staged through
// shared memory 32 dimensions at a time, so every global element is read once per tile and the differences x_i - x_j are
// formed directly (no Gram-matrix expansion: ||x||^2 + ||x'||^2 - 2 x.x' cancels catastrophically for close samples,
// which is exactly where the IMQ kernel has its weight).  Pair values are float32 (the dtype of the samples, x64 off);
// they are summed in float64, per block in a fixed order and then over blocks in a fixed order: deterministic.
// T <= ~13 k (eval_iter * num_chain), d = 2..1600: a few hundred MFLOP at most - latency, not bandwidth.
#include "internal.h"

namespace mfm {
namespace {

constexpr int PT = 16, PK = 32;

// MODE 0: Stein discrepancy integrand with the IMQ kernel (mcmc_utils.py:66-75), b = -beta (> 0):
//   -4 b (b+1) r / (1+r)^(b+2) + 2 b (d + (g - g').(x - x')) / (1+r)^(1+b) + g.g' / (1+r)^b,   r = |x - x'|^2
// MODE 1: Gaussian kernel exp(-r / 2) (max_mean_disc, :98-100)
template <int MODE>
__global__ void __launch_bounds__(PT * PT) pair_kernel(const float* __restrict__ X, const float* __restrict__ GX, int TX,
                                                      const float* __restrict__ Y, const float* __restrict__ GY, int TY, int d, float b,
                                                      double* __restrict__ block_sum, double* __restrict__ block_diag) {
    __shared__ float xi[PT][PK + 1], xj[PT][PK + 1], gi[MODE == 0 ? PT : 1][PK + 1], gj[MODE == 0 ? PT : 1][PK + 1];
    __shared__ double red[PT * PT / 32], redd[PT * PT / 32];
    const int tx = threadIdx.x % PT, ty = threadIdx.x / PT;
    const int i = blockIdx.y * PT + ty, j = blockIdx.x * PT + tx;
    float r = 0.0f, gd = 0.0f, gg = 0.0f;
    for (int k0 = 0; k0 < d; k0 += PK) {
        for (int e = threadIdx.x; e < PT * PK; e += PT * PT) {
            const int row = e / PK, k = e % PK;
            const int ri = blockIdx.y * PT + row, rj = blockIdx.x * PT + row;
            const bool ki = k0 + k < d;
            xi[row][k] = (ri < TX && ki) ? X[(long long)ri * d + k0 + k] : 0.0f;
            xj[row][k] = (rj < TY && ki) ? Y[(long long)rj * d + k0 + k] : 0.0f;
            if (MODE == 0) {
                gi[row][k] = (ri < TX && ki) ? GX[(long long)ri * d + k0 + k] : 0.0f;
                gj[row][k] = (rj < TY && ki) ? GY[(long long)rj * d + k0 + k] : 0.0f;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < PK; ++k) {
            const float diff = xi[ty][k] - xj[tx][k];
            r += diff * diff;
            if (MODE == 0) { gd += (gi[ty][k] - gj[tx][k]) * diff; gg += gi[ty][k] * gj[tx][k]; }
        }
        __syncthreads();
    }
    double v = 0.0;
    if (i < TX && j < TY) {
        if (MODE == 0) {
            const float q = 1.0f + r;
            const float pb = powf(q, b);                         // (1+r)^b
            v = (double)(-4.0f * b * (b + 1.0f) * r / (pb * q * q) + 2.0f * b * ((float)d + gd) / (pb * q) + gg / pb);
        } else {
            v = (double)expf(-0.5f * r);
        }
    }
    double vd = (i == j) ? v : 0.0;
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); vd += __shfl_xor_sync(0xffffffffu, vd, o); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = v; redd[threadIdx.x >> 5] = vd; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0, sd = 0.0;
        for (int w = 0; w < PT * PT / 32; ++w) { s += red[w]; sd += redd[w]; }
        const long long bid = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        block_sum[bid] = s; block_diag[bid] = sd;
    }
}

// fixed-order final sums; MODE 0: out = (U, V); MODE 1: out[slot] = sum (diag unused)
__global__ void __launch_bounds__(1024) pair_final_kernel(const double* __restrict__ block_sum, const double* __restrict__ block_diag,
                                                          long long n_blocks, double* __restrict__ out2) {
    __shared__ double red[32], redd[32];
    double s = 0.0, sd = 0.0;
    const long long per = (n_blocks + blockDim.x - 1) / blockDim.x;
    const long long b0 = threadIdx.x * per, b1 = min(n_blocks, b0 + per);
    for (long long k = b0; k < b1; ++k) { s += block_sum[k]; sd += block_diag[k]; }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); sd += __shfl_xor_sync(0xffffffffu, sd, o); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; redd[threadIdx.x >> 5] = sd; }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = 0.0; sd = 0.0;
        for (int w = 0; w < 32; ++w) { s += red[w]; sd += redd[w]; }
        out2[0] = s; out2[1] = sd;
    }
}

__global__ void stein_finish_kernel(const double* __restrict__ s2, int T, float* __restrict__ out) {
    const double t = (double)T;
    out[0] = (float)((s2[0] - s2[1]) / (t * (t - 1.0)));       // U-statistic (:85)
    out[1] = (float)(s2[0] / (t * t));                         // V-statistic
}

__global__ void mmd_finish_kernel(const double* __restrict__ sxx, const double* __restrict__ syy, const double* __restrict__ sxy, int m,
                                  float* __restrict__ out) {
    const double mm = (double)m, m2 = mm * mm;
    out[0] = (float)((sxx[0] - mm) / (m2 - mm) - 2.0 * sxy[0] / m2 + (syy[0] - mm) / (m2 - mm));   // :104-109
}

inline long long n_pair_blocks(int TX, int TY) { return (long long)ceil_div(TX, PT) * ceil_div(TY, PT); }

}  // namespace
}  // namespace mfm

extern "C" {

size_t mfm_pairwise_workspace_bytes(int t) {
    const long long nb = mfm::n_pair_blocks(t, t);
    return 2 * mfm::ws_slice((size_t)nb, sizeof(double)) + mfm::ws_slice(8, sizeof(double));
}

int mfm_stein_disc(const float* X, const float* grad_logp, int T, int d, float beta, float* out_uv, void* ws, size_t ws_bytes,
                   mfm_stream_t stream) {
    if (!X || !grad_logp || !out_uv || T < 2 || d < 1) { mfm_set_last_error_msg("bad argument (mfm_stein_disc)"); return MFM_ERR_ARG; }
    mfm::Workspace w(ws, ws_bytes);
    const long long nb = mfm::n_pair_blocks(T, T);
    double* bs = w.take<double>((size_t)nb); double* bd = w.take<double>((size_t)nb); double* s2 = w.take<double>(8);
    if (!w.ok) { mfm_set_last_error_msg("workspace too small (mfm_stein_disc)"); return MFM_ERR_WORKSPACE; }
    dim3 grid(ceil_div(T, mfm::PT), ceil_div(T, mfm::PT));
    mfm::pair_kernel<0><<<grid, mfm::PT * mfm::PT, 0, stream>>>(X, grad_logp, T, X, grad_logp, T, d, -beta, bs, bd);
    MFM_LAUNCH_CHECK();
    mfm::pair_final_kernel<<<1, 1024, 0, stream>>>(bs, bd, nb, s2);
    MFM_LAUNCH_CHECK();
    mfm::stein_finish_kernel<<<1, 1, 0, stream>>>(s2, T, out_uv);
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

int mfm_max_mean_disc(const float* X, const float* Y, int m, int d, float* out, void* ws, size_t ws_bytes, mfm_stream_t stream) {
    if (!X || !Y || !out || m < 2 || d < 1) { mfm_set_last_error_msg("bad argument (mfm_max_mean_disc)"); return MFM_ERR_ARG; }
    mfm::Workspace w(ws, ws_bytes);
    const long long nb = mfm::n_pair_blocks(m, m);
    double* bs = w.take<double>((size_t)nb); double* bd = w.take<double>((size_t)nb); double* s2 = w.take<double>(8);
    if (!w.ok) { mfm_set_last_error_msg("workspace too small (mfm_max_mean_disc)"); return MFM_ERR_WORKSPACE; }
    dim3 grid(ceil_div(m, mfm::PT), ceil_div(m, mfm::PT));
    const float* A[3] = {X, Y, X}; const float* B[3] = {X, Y, Y};
    for (int k = 0; k < 3; ++k) {
        mfm::pair_kernel<1><<<grid, mfm::PT * mfm::PT, 0, stream>>>(A[k], nullptr, m, B[k], nullptr, m, d, 0.0f, bs, bd);
        MFM_LAUNCH_CHECK();
        mfm::pair_final_kernel<<<1, 1024, 0, stream>>>(bs, bd, nb, s2 + 2 * k);
        MFM_LAUNCH_CHECK();
    }
    mfm::mmd_finish_kernel<<<1, 1, 0, stream>>>(s2, s2 + 2, s2 + 4, m, out);
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

}  // extern "C"
