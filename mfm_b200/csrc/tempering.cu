// Adaptive tempering: next inverse temperature from an effective-sample-size target.
// Replaces beta_fn (exe_flow_matching.py:391-402): jaxopt.Bisection(optimality_fun=ess_zero,
// lower=prev_beta, upper=1, maxiter=30, tol=1e-5, check_bracket=False).run().params with
// ess_zero(beta) = 1/sum(w^2) - alpha*N, w = softmax(loglik * (beta - prev_beta)).
// jaxopt 0.8.3 Bisection is restated in the kernel (third-party, parity unpinned): the sign of the
// bracket is taken from f(lower), f(upper); every iteration evaluates the midpoint, shrinks the
// bracket, and stops when |f(mid)| <= tol; the last midpoint is returned.
#include "internal.h"

namespace mfm {

__device__ float ess_zero(const float* __restrict__ ll, int n, float dbeta, float alpha_n, float* red) {
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, ll[i] * dbeta);
    // block max
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < (int)((blockDim.x + 31) >> 5); ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const float e = expf(ll[i] * dbeta - mx); s1 += e; s2 += e * e; }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    // weights = e / s1 ; 1 / sum(weights^2) - alpha * N
    return 1.0f / (s2 / (s1 * s1)) - alpha_n;
}

__global__ void __launch_bounds__(1024) tempering_kernel(const float* __restrict__ ll, int n, const float* __restrict__ prev_beta_p,
                                                         float alpha, float* __restrict__ beta_out) {
    __shared__ float red[32];
    const float prev = prev_beta_p[0];
    const float alpha_n = alpha * (float)n;
    float lo = prev, hi = 1.0f;
    const float flo = ess_zero(ll, n, lo - prev, alpha_n, red);
    const float fhi = ess_zero(ll, n, hi - prev, alpha_n, red);
    const float sign = (flo < 0.0f && fhi >= 0.0f) ? 1.0f : ((flo > 0.0f && fhi <= 0.0f) ? -1.0f : 0.0f);
    float mid = lo, err = INFINITY;
    for (int it = 0; it < 30 && err > 1e-5f; ++it) {
        mid = 0.5f * (hi + lo);
        const float v = ess_zero(ll, n, mid - prev, alpha_n, red);
        const bool too_large = sign * v > 0.0f;
        hi = too_large ? mid : hi;
        lo = too_large ? lo : mid;
        err = fabsf(v);
    }
    if (threadIdx.x == 0) beta_out[0] = mid;
}

}  // namespace mfm

extern "C" int mfm_tempering_beta(const float* logliks, int n, const float* prev_beta, float alpha, float* beta_out,
                                  mfm_stream_t stream) {
    if (!logliks || !prev_beta || !beta_out || n <= 0) { mfm_set_last_error_msg("bad argument (mfm_tempering_beta)"); return MFM_ERR_ARG; }
    mfm::tempering_kernel<<<1, 1024, 0, stream>>>(logliks, n, prev_beta, alpha, beta_out);
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}
