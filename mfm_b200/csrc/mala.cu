// Fused MALA transition: keys = split(rng_key, N); per chain split -> (key_integrator, key_rmh);
// Euler-Langevin proposal; target value+grad at the proposal; asymmetric transition energies;
// accept/select; info.  One warp per chain, coalesced row access, shuffle reductions.
//
// Replaces (as coded, including the sign convention of proposal_from_energy_diff):
//   bblackjax/mcmc/mala.py:68-79,86-118, diffusions.py:22-33, util.py:57-82,
//   proposal.py:104-112,152-159,178-186, exe_flow_matching.py:303,313.
#include "internal.h"
#include "targets.cuh"

namespace mfm {

constexpr int MALA_WARPS = 4;

struct ChainKeys { u32x2 integrator, rmh; };

// keys[c] = split(rng_key, n_total)[c];  (key_integrator, key_rmh) = split(keys[c])
// n_total == 0: rng_key is the per-chain key array uint32[n,2] (already split).
__device__ __forceinline__ ChainKeys derive_chain_keys(const uint32_t* rng_key, int c, int n_total) {
    u32x2 kc;
    if (n_total > 0) kc = threefry_split_key(rng_key[0], rng_key[1], (uint32_t)c, (uint32_t)n_total);
    else { kc.a = rng_key[2 * c]; kc.b = rng_key[2 * c + 1]; }
    // split(kc, 2): counts [0,1,2,3] -> blocks (0,2),(1,3); keys [[o0(0,2), o0(1,3)], [o1(0,2), o1(1,3)]]
    const u32x2 b0 = threefry2x32(kc.a, kc.b, 0u, 2u);
    const u32x2 b1 = threefry2x32(kc.a, kc.b, 1u, 3u);
    ChainKeys r; r.integrator.a = b0.a; r.integrator.b = b1.a; r.rmh.a = b0.b; r.rmh.b = b1.b;
    return r;
}

// uniform(key, ()) : one-element stream, padded -> block (0,0), word o0
__device__ __forceinline__ float scalar_uniform(u32x2 key, int x64) {
    return rng_uniform_at(key.a, key.b, 0u, 1u, x64);      // float32: block (0,0) word o0; float64 draws: block (0,1), both words
}

// Writes the proposal into xs (smem) and returns sum((x' - x - h g)^2) (warp-reduced).
__device__ __forceinline__ float langevin_propose(u32x2 key, int d, float h, float sq2h, const float* __restrict__ x,
                                                  const float* __restrict__ g, float* xs, int lane, int x64, const double2* ltab) {
    const uint32_t half = ((uint32_t)d + 1u) >> 1;
    float sq = 0.0f;
    for (uint32_t b = lane; b < half; b += 32) {
        const uint32_t hi = b + half;
        const bool has_hi = hi < (uint32_t)d;
        float n_lo, n_hi = 0.0f;
        if (x64) {          // float64 draws: element e is block (e, d + e)
            n_lo = rng_normal_at(key.a, key.b, b, (uint32_t)d, 1);
            if (has_hi) n_hi = rng_normal_at(key.a, key.b, hi, (uint32_t)d, 1);
        } else {
            const u32x2 o = threefry2x32(key.a, key.b, b, has_hi ? hi : 0u);
            n_lo = bits_to_normal_t(o.a, ltab); n_hi = bits_to_normal_t(o.b, ltab);
        }
        {
            const float xv = x[b], gv = g[b];
            // p + step_size*g + sqrt(2 step_size)*n, left to right as diffusions.py:25-30
            const float xn = __fadd_rn(__fadd_rn(xv, __fmul_rn(h, gv)), __fmul_rn(sq2h, n_lo));
            xs[b] = xn;
            const float th = __fadd_rn(__fadd_rn(xn, -xv), -__fmul_rn(h, gv));
            sq += th * th;
        }
        if (has_hi) {
            const float xv = x[hi], gv = g[hi];
            const float xn = __fadd_rn(__fadd_rn(xv, __fmul_rn(h, gv)), __fmul_rn(sq2h, n_hi));
            xs[hi] = xn;
            const float th = __fadd_rn(__fadd_rn(xn, -xv), -__fmul_rn(h, gv));
            sq += th * th;
        }
    }
    return warp_sum(sq);
}

struct MalaIO {
    float* x; float* l; float* g;              // state (in place)
    float* acc_rate; uint8_t* is_acc; float* prop_pos; float* prop_w;   // info
};

// accept/select given proposal quantities; one warp, all lanes hold the scalars.
__device__ __forceinline__ void mala_accept_select(const MalaIO& io, int c, int d, float h, float quarter, float u,
                                                   float l_old, float sq_new, float l_new, const float* xs,
                                                   const float* gs, float beta_scale, int lane) {
    const float* x = io.x + (long long)c * d;
    // theta_prev = x - x' - h g'
    float sq_prev = 0.0f;
    for (int i = lane; i < d; i += 32) {
        const float th = __fadd_rn(__fadd_rn(x[i], -xs[i]), -__fmul_rn(h, beta_scale * gs[i]));
        sq_prev += th * th;
    }
    sq_prev = warp_sum(sq_prev);
    const float e_new = -l_old + quarter * sq_new;     // transition_energy(state, new_state)
    const float e_prev = -l_new + quarter * sq_prev;   // transition_energy(new_state, state)
    float delta = e_prev - e_new;                      // proposal.py:104
    if (isnan(delta)) delta = -INFINITY;
    const float p_accept = fminf(expf(delta), 1.0f);   // proposal.py:178
    const bool acc = u < p_accept;                     // jax.random.bernoulli: strict <
    float* xo = io.x + (long long)c * d;
    float* go = io.g + (long long)c * d;
    float* pp = io.prop_pos ? io.prop_pos + (long long)c * d : nullptr;
    for (int i = lane; i < d; i += 32) {
        const float xn = xs[i];
        if (pp) pp[i] = xn;
        if (acc) { xo[i] = xn; go[i] = beta_scale * gs[i]; }
    }
    if (lane == 0) {
        if (acc) io.l[c] = l_new;
        if (io.acc_rate) io.acc_rate[c] = p_accept;
        if (io.is_acc) io.is_acc[c] = acc ? 1 : 0;
        if (io.prop_w) io.prop_w[c] = expf(l_new + quarter * sq_prev);   // mala.py:113
    }
}

__global__ void __launch_bounds__(MALA_WARPS * 32)
mala_small_kernel(mfm_target_t T, const uint32_t* __restrict__ rng_key, int n, int chain_offset, int n_total, float h,
                  float sq2h, float quarter, MalaIO io, int x64) {
    extern __shared__ float sm[];
    __shared__ double2 ltab[16];
    log_tab_load(ltab);                        // (before any warp leaves: contains a barrier)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * MALA_WARPS + w;
    if (c >= n) return;
    const int d = T.dim;
    float* xs = sm + w * 2 * d;
    float* gs = xs + d;
    const ChainKeys ck = derive_chain_keys(rng_key, chain_offset + c, n_total);
    const float sq_new = langevin_propose(ck.integrator, d, h, sq2h, io.x + (long long)c * d, io.g + (long long)c * d, xs, lane, x64, ltab);
    __syncwarp();
    const float ll = small_target_loglik_grad(T, xs, gs, lane);
    __syncwarp();
    const float l_new = T.beta * ll;
    const float u = scalar_uniform(ck.rmh, x64);
    mala_accept_select(io, c, d, h, quarter, u, io.l[c], sq_new, l_new, xs, gs, T.beta, lane);
}

// ---- pines: propose -> GEMM (K^-1) -> finalize -----------------------------------------------
__global__ void __launch_bounds__(256)
pines_propose_kernel(mfm_target_t T, const uint32_t* __restrict__ rng_key, int n, int chain_offset, int n_total,
                     float h, float sq2h, const float* __restrict__ x, const float* __restrict__ g,
                     float* __restrict__ xprop, float* __restrict__ lik, float* __restrict__ sq_new_out,
                     float* __restrict__ u_out, float* __restrict__ xprop_amax, int x64) {
    __shared__ double2 ltab[16];
    log_tab_load(ltab);                        // (before any warp leaves: contains a barrier)
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    const int d = T.dim;
    const ChainKeys ck = derive_chain_keys(rng_key, chain_offset + c, n_total);
    float* xp = xprop + (long long)c * d;
    const float sq = langevin_propose(ck.integrator, d, h, sq2h, x + (long long)c * d, g + (long long)c * d, xp, lane, x64, ltab);
    __syncwarp();
    float s = 0.0f, vm = 0.0f;
    for (int i = lane; i < d; i += 32) {
        const float xv = xp[i];
        s += xv * T.counts[i] - T.poisson_a * expf(xv);
        vm = fmaxf(vm, fabsf(xv));
    }
    s = warp_sum(s);
    if (lane == 0) { lik[c] = s; sq_new_out[c] = sq; u_out[c] = scalar_uniform(ck.rmh, x64); }
    if (xprop_amax) amax_publish_warp(xprop_amax, vm);      // the proposal is the A operand of the K^-1 GEMM (scaled-fp16 split)
}

// whitened pines: the proposal's log-density and gradient come from target_value_and_grad (two GEMMs against the Cholesky factor)
__global__ void __launch_bounds__(256)
white_mala_finalize_kernel(int n, int d, float h, float quarter, const float* __restrict__ xprop, const float* __restrict__ gnew,
                           const float* __restrict__ lnew, const float* __restrict__ sq_new, const float* __restrict__ u, MalaIO io) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    MalaIO io2 = io; io2.prop_pos = nullptr;    // x' already lives in prop_pos (written by propose)
    mala_accept_select(io2, c, d, h, quarter, u[c], io.l[c], sq_new[c], lnew[c], xprop + (long long)c * d, gnew + (long long)c * d, 1.0f, lane);
}

__global__ void __launch_bounds__(256)
pines_mala_finalize_kernel(mfm_target_t T, int n, int n_tiles, float h, float quarter, const float* __restrict__ xprop,
                           const float* __restrict__ gnew, const float* __restrict__ lik,
                           const float* __restrict__ partial, const float* __restrict__ sq_new,
                           const float* __restrict__ u, MalaIO io) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    float s = 0.0f;
    for (int t = 0; t < n_tiles; ++t) s += partial[(long long)c * n_tiles + t];
    const float l_new = T.beta * lik[c] + (-0.5f * s + T.log_norm);
    MalaIO io2 = io; io2.prop_pos = nullptr;    // x' already lives in prop_pos (written by propose)
    mala_accept_select(io2, c, T.dim, h, quarter, u[c], io.l[c], sq_new[c], l_new, xprop + (long long)c * T.dim,
                       gnew + (long long)c * T.dim, 1.0f, lane);
}

}  // namespace mfm

extern "C" {

size_t mfm_mala_workspace_bytes(const mfm_target_t* t, int n) {
    using namespace mfm;
    if (t->kind == MFM_TARGET_PINES_WHITE)
        return ws_slice((size_t)n * t->dim, 4) * 2 + 4 * ws_slice(n, 4) + ws_slice(64, 4) + target_ws_bytes(*t, n) + 512;
    if (t->kind != MFM_TARGET_PINES) return 256;
    return ws_slice((size_t)n * t->dim, 4) * 2 + ws_slice((size_t)n * pines_n_tiles(t->dim), 4) + 3 * ws_slice(n, 4) + ws_slice(64, 4) + 256;
}

int mfm_mala_step(const mfm_target_t* t, const uint32_t* rng_key, int per_chain_keys, int n, int chain_offset, int n_total, float step_size,
                  float* position, float* logdensity, float* logdensity_grad, float* acceptance_rate,
                  uint8_t* is_accepted, float* proposed_position, float* proposed_weight, void* ws, size_t ws_bytes,
                  mfm_stream_t stream) {
    using namespace mfm;
    if (!t || !rng_key || !position || !logdensity || !logdensity_grad) { mfm_set_last_error_msg("null argument"); return MFM_ERR_ARG; }
    if (n <= 0) return MFM_OK;
    if (per_chain_keys) { chain_offset = 0; n_total = 0; }
    else if (n_total < chain_offset + n || chain_offset < 0) { mfm_set_last_error_msg("bad chain_offset/n_total"); return MFM_ERR_ARG; }
    const mfm_target_t& T = *t;
    const float h = step_size;
    const float sq2h = sqrtf(2.0f * h);                         // jnp.sqrt(2*step_size) in f32
    const float quarter = (float)(0.25 * (1.0 / (double)step_size));   // python-float arithmetic, mala.py:79
    MalaIO io{position, logdensity, logdensity_grad, acceptance_rate, is_accepted, proposed_position, proposed_weight};
    if (T.kind == MFM_TARGET_PINES_WHITE) {
        Workspace w(ws, ws_bytes);
        float* xprop = proposed_position ? proposed_position : w.take<float>((size_t)n * T.dim);
        if (proposed_position) w.take<float>((size_t)n * T.dim);
        float* gnew = w.take<float>((size_t)n * T.dim);
        float* lik = w.take<float>(n); float* sqn = w.take<float>(n); float* u = w.take<float>(n); float* lnew = w.take<float>(n);
        float* xamax = w.take<float>(64);
        Workspace wt((char*)ws + w.off, w.off <= ws_bytes ? ws_bytes - w.off : 0);
        if (!w.ok) { mfm_set_last_error_msg("workspace too small (mfm_mala_step)"); return MFM_ERR_WORKSPACE; }
        // (the propose kernel's Poisson sum over x' is the UNwhitened likelihood: unused here)
        pines_propose_kernel<<<ceil_div(n, 8), 256, 0, stream>>>(T, rng_key, n, chain_offset, n_total, h, sq2h, position,
                                                                  logdensity_grad, xprop, lik, sqn, u, nullptr, rng_x64());
        MFM_LAUNCH_CHECK();
        int rc = target_value_and_grad(T, n, xprop, lnew, gnew, nullptr, wt, stream);
        if (rc) return rc;
        white_mala_finalize_kernel<<<ceil_div(n, 8), 256, 0, stream>>>(n, T.dim, h, quarter, xprop, gnew, lnew, sqn, u, io);
        MFM_LAUNCH_CHECK();
        (void)xamax;
        return MFM_OK;
    }
    if (T.kind == MFM_TARGET_PINES) {
        Workspace w(ws, ws_bytes);
        const int nt = pines_n_tiles(T.dim);
        float* xprop = proposed_position ? proposed_position : w.take<float>((size_t)n * T.dim);
        if (proposed_position) w.take<float>((size_t)n * T.dim);   // keep the layout stable
        float* gnew = w.take<float>((size_t)n * T.dim);
        float* partial = w.take<float>((size_t)n * nt);
        float* lik = w.take<float>(n); float* sqn = w.take<float>(n); float* u = w.take<float>(n);
        float* xamax = w.take<float>(64);
        if (!w.ok) { mfm_set_last_error_msg("workspace too small (mfm_mala_step)"); return MFM_ERR_WORKSPACE; }
        MFM_CUDA_CHECK(cudaMemsetAsync(xamax, 0, sizeof(float), stream));
        pines_propose_kernel<<<ceil_div(n, 8), 256, 0, stream>>>(T, rng_key, n, chain_offset, n_total, h, sq2h, position,
                                                                  logdensity_grad, xprop, lik, sqn, u, xamax, rng_x64());
        MFM_LAUNCH_CHECK();
        int rc = pines_grad_gemm(T, n, xprop, T.dim, T.beta, gnew, T.dim, partial, nullptr, stream, xamax);
        if (rc) return rc;
        pines_mala_finalize_kernel<<<ceil_div(n, 8), 256, 0, stream>>>(T, n, nt, h, quarter, xprop, gnew, lik, partial, sqn, u, io);
        MFM_LAUNCH_CHECK();
        return MFM_OK;
    }
    if (T.kind == MFM_TARGET_GMM && T.dim != 2) { mfm_set_last_error_msg("GMM target requires dim == 2"); return MFM_ERR_ARG; }
    const size_t smem = (size_t)MALA_WARPS * 2 * T.dim * sizeof(float);
    if (smem > 200 * 1024) { mfm_set_last_error_msg("dim too large for warp-per-chain MALA"); return MFM_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) MFM_CUDA_CHECK(cudaFuncSetAttribute(mala_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mala_small_kernel<<<ceil_div(n, MALA_WARPS), MALA_WARPS * 32, smem, stream>>>(T, rng_key, n, chain_offset, n_total, h, sq2h, quarter, io, rng_x64());
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

}  // extern "C"
