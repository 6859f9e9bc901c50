// Final sampling of the reference run (exe_flow_matching.py:453-459): importance weights of the flow samples and
// jax.random.choice(key, flow_samples, (n,), p=weights) - the multinomial resampling that turns them into "exact samples".
//
// jax.random.choice with replace=True and p given (jax 0.4.26 random.py, third-party, restated; parity unpinned):
//     p_cuml = cumsum(p);  r = p_cuml[-1] * (1 - uniform(key, shape, p.dtype));  ind = searchsorted(p_cuml, r)   (side='left')
// The cumulative sum is taken SEQUENTIALLY in float32 (one warp, every lane carrying the same running sum), which is what
// NumPy's cumsum does and therefore what the oracle checks bit for bit; XLA's own summation order is not documented, so
// against JAX an index may differ where r falls within rounding of a boundary.  O(n) on one warp is fine here: this runs
// once per training run on n = eval_iter * num_chain values (12 800 for the 2-d examples, 128..1 024 for phi-four / pines).
#include "internal.h"

namespace mfm {
namespace {

__global__ void __launch_bounds__(1024) importance_weights_kernel(const float* __restrict__ logp, const float* __restrict__ ref_logp,
                                                                  const float* __restrict__ vols, int n, float* __restrict__ log_w,
                                                                  float* __restrict__ w) {
    __shared__ float red[32];
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float lw = logp[i] - ref_logp[i] - vols[i];          // :457
        log_w[i] = lw;
        mx = fmaxf(mx, lw);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmaxf(mx, red[k]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) w[i] = expf(log_w[i] - mx);   // :458 (own writes: no barrier needed)
}

__global__ void __launch_bounds__(32) cumsum_seq_kernel(const float* __restrict__ p, int n, float* __restrict__ cum) {
    const int lane = threadIdx.x;
    float acc = 0.0f;
    for (int base = 0; base < n; base += 32) {
        const float v = base + lane < n ? p[base + lane] : 0.0f;   // one coalesced load per 32 values
        float mine = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            acc += __shfl_sync(0xffffffffu, v, j);                // same order in every lane: ((p0 + p1) + p2) + ...
            if (lane == j) mine = acc;
        }
        if (base + lane < n) cum[base + lane] = mine;
    }
}

__global__ void choice_kernel(const uint32_t* __restrict__ key, const float* __restrict__ cum, int n_pop, int n_draw,
                              int* __restrict__ idx, int x64) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_draw) return;
    const float u = rng_uniform_at(key[0], key[1], (uint32_t)i, (uint32_t)n_draw, x64);
    const float r = cum[n_pop - 1] * (1.0f - u);
    int lo = 0, hi = n_pop;                                        // first index with cum[index] >= r
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] < r) lo = mid + 1; else hi = mid;
    }
    idx[i] = lo;
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, int n_pop, int n_draw, int d,
                                   float* __restrict__ out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)n_draw * d) return;
    const int row = (int)(t / d), col = (int)(t % d);
    const int s = min(max(idx[row], 0), n_pop - 1);               // jnp.take clamps out-of-range indices
    out[t] = src[(long long)s * d + col];
}

}  // namespace
}  // namespace mfm

extern "C" {

int mfm_importance_weights(const float* logdensity, const float* ref_logdensity, const float* vols, int n, float* log_weights,
                           float* weights, mfm_stream_t stream) {
    if (!logdensity || !ref_logdensity || !vols || !log_weights || !weights || n <= 0) { mfm_set_last_error_msg("bad argument (mfm_importance_weights)"); return MFM_ERR_ARG; }
    mfm::importance_weights_kernel<<<1, 1024, 0, stream>>>(logdensity, ref_logdensity, vols, n, log_weights, weights);
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

size_t mfm_random_choice_workspace_bytes(int n_pop) { return mfm::ws_slice((size_t)(n_pop > 0 ? n_pop : 0), sizeof(float)); }

int mfm_random_choice(const uint32_t* key, int n_pop, const float* p, int n_draw, int* idx_out, void* ws, size_t ws_bytes,
                      mfm_stream_t stream) {
    if (!key || !p || !idx_out || n_pop <= 0 || n_draw < 0) { mfm_set_last_error_msg("bad argument (mfm_random_choice)"); return MFM_ERR_ARG; }
    mfm::Workspace w(ws, ws_bytes);
    float* cum = w.take<float>((size_t)n_pop);
    if (!w.ok) { mfm_set_last_error_msg("workspace too small (mfm_random_choice)"); return MFM_ERR_WORKSPACE; }
    if (n_draw == 0) return MFM_OK;
    mfm::cumsum_seq_kernel<<<1, 32, 0, stream>>>(p, n_pop, cum);
    MFM_LAUNCH_CHECK();
    mfm::choice_kernel<<<ceil_div(n_draw, 256), 256, 0, stream>>>(key, cum, n_pop, n_draw, idx_out, mfm::rng_x64());
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

int mfm_gather_rows(const float* src, const int* idx, int n_pop, int n_draw, int d, float* out, mfm_stream_t stream) {
    if (!src || !idx || !out || n_pop <= 0 || n_draw < 0 || d <= 0) { mfm_set_last_error_msg("bad argument (mfm_gather_rows)"); return MFM_ERR_ARG; }
    if (n_draw == 0) return MFM_OK;
    mfm::gather_rows_kernel<<<ceil_div((long long)n_draw * d, 256), 256, 0, stream>>>(src, idx, n_pop, n_draw, d, out);
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

}  // extern "C"
