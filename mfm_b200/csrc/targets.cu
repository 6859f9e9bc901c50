// Target evaluation kernels: logprob_beta value+grad (mala.init), and the untempered
// grad / Hessian terms the vector field needs.
// Reference: distributions.py:58-67,131-160,299-307; cox_process_utils.py:98-115,142-165;
// bblackjax/mcmc/mala.py:51-54; exe_flow_matching.py:301,316,351.
#include "internal.h"
#include "targets.cuh"
#include "gemm_tf32x3.cuh"

namespace mfm {

// -------------------------------------------------------------------------------------------
// small targets: one warp per chain
// -------------------------------------------------------------------------------------------
constexpr int EVAL_WARPS = 4;

__global__ void small_value_grad_kernel(mfm_target_t T, int n, const float* __restrict__ x, float* __restrict__ logp,
                                        float* __restrict__ grad, float* __restrict__ loglik_out) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * EVAL_WARPS + w;
    if (c >= n) return;
    const int d = T.dim;
    float* xs = sm + w * 2 * d;
    float* gs = xs + d;
    for (int i = lane; i < d; i += 32) xs[i] = x[(long long)c * d + i];
    __syncwarp();
    const float ll = small_target_loglik_grad(T, xs, gs, lane);
    __syncwarp();
    for (int i = lane; i < d; i += 32) grad[(long long)c * d + i] = T.beta * gs[i];
    if (lane == 0) {
        logp[c] = T.beta * ll;     // logprior == 0 for these targets
        if (loglik_out) loglik_out[c] = ll;
    }
}

// field terms for small targets: gc = clip(grad), hvc = inrange * (H z), hdc = inrange * diag(H)
__global__ void small_field_terms_kernel(mfm_target_t T, int n, const int* __restrict__ n_rows_dev, float clip,
                                         const float* __restrict__ x, const float* __restrict__ z,
                                         float* __restrict__ gc, float* __restrict__ hvc, float* __restrict__ hdc) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * EVAL_WARPS + w;
    if (n_rows_dev) n = min(n, *n_rows_dev);
    if (c >= n) return;
    const int d = T.dim;
    float* xs = sm + w * 5 * d;
    float* gs = xs + d; float* zs = gs + d; float* hv = zs + d; float* hd = hv + d;
    for (int i = lane; i < d; i += 32) {
        xs[i] = x[(long long)c * d + i];
        if (z) zs[i] = z[(long long)c * d + i];
    }
    __syncwarp();
    small_target_loglik_grad(T, xs, gs, lane);
    if (hvc || hdc) small_target_hess(T, xs, zs, hvc ? hv : nullptr, hdc ? hd : nullptr, lane);
    __syncwarp();
    for (int i = lane; i < d; i += 32) {
        const float g = gs[i];
        const bool in = !(clip > 0.0f) || (g > -clip && g < clip);
        gc[(long long)c * d + i] = clip > 0.0f ? fminf(fmaxf(g, -clip), clip) : g;
        if (hvc) hvc[(long long)c * d + i] = in ? hv[i] : 0.0f;
        if (hdc) hdc[(long long)c * d + i] = in ? hd[i] : 0.0f;
    }
}

// -------------------------------------------------------------------------------------------
// pines: dense prior as a GEMM against K^-1 with fused epilogues
// -------------------------------------------------------------------------------------------
struct EpiPinesGrad {
    static constexpr bool kRowSum = true;
    const float* X; long long ldx;
    const float* counts; const float* kinv_mu;
    float mu, a, beta;
    float* grad; long long ldg;
    float* partial; int n_tiles;
    struct Aux { float kinv_mu, x, counts; };
    __device__ __forceinline__ Aux load(int row, int col) const {
        Aux u; u.kinv_mu = __ldg(kinv_mu + col); u.x = __ldg(X + (long long)row * ldx + col); u.counts = __ldg(counts + col);
        return u;
    }
    __device__ __forceinline__ float apply(int row, int col, float acc, const Aux& u) const {
        const float q = acc - u.kinv_mu;
        grad[(long long)row * ldg + col] = beta * (u.counts - a * expf(u.x)) - q;
        return (u.x - mu) * q;
    }
    __device__ __forceinline__ float operator()(int row, int col, float acc) const { return apply(row, col, acc, load(row, col)); }
    struct Col4 { float4 kinv_mu, counts; };
    struct Row4 { float4 x; };
    bool vec_ok() const { return aligned16(X) && ldx % 4 == 0 && aligned16(grad) && ldg % 4 == 0 && aligned16(counts) && aligned16(kinv_mu); }
    __device__ __forceinline__ Col4 load_col4(int col) const { Col4 c; c.kinv_mu = ldg4(kinv_mu + col); c.counts = ldg4(counts + col); return c; }
    __device__ __forceinline__ Row4 load_row4(int row, int col) const { Row4 r; r.x = ldg4(X + (long long)row * ldx + col); return r; }
    __device__ __forceinline__ float apply4(int row, int col, const float4& acc, const Col4& c, const Row4& r) const {
        const float4 q = make_float4(acc.x - c.kinv_mu.x, acc.y - c.kinv_mu.y, acc.z - c.kinv_mu.z, acc.w - c.kinv_mu.w);
        float4 g;
        g.x = beta * (c.counts.x - a * expf(r.x.x)) - q.x; g.y = beta * (c.counts.y - a * expf(r.x.y)) - q.y;
        g.z = beta * (c.counts.z - a * expf(r.x.z)) - q.z; g.w = beta * (c.counts.w - a * expf(r.x.w)) - q.w;
        st4(grad + (long long)row * ldg + col, g);
        return ((r.x.x - mu) * q.x + (r.x.y - mu) * q.y) + ((r.x.z - mu) * q.z + (r.x.w - mu) * q.w);
    }
    __device__ __forceinline__ void row_partial(int row, int tile, float s) const {
        if (partial) partial[(long long)row * n_tiles + tile] = s;
    }
    __device__ __forceinline__ void at_z(int) {}
};

struct EpiPinesField {
    static constexpr bool kRowSum = false;
    const float* X; long long ldx;
    const float* counts; const float* kinv_mu; const float* kinv_diag;
    const float* Z; const float* ZK;     // probe and Z K^-1 (hutch), or null
    float a, clip;
    float* gc; float* hvc; float* hdc; long long ld;
    struct Aux { float x, counts, kinv_mu, z, zk, kdiag; };
    __device__ __forceinline__ Aux load(int row, int col) const {
        const long long o = (long long)row * ld + col;
        Aux u; u.x = __ldg(X + (long long)row * ldx + col); u.counts = __ldg(counts + col); u.kinv_mu = __ldg(kinv_mu + col);
        u.z = hvc ? __ldg(Z + o) : 0.0f; u.zk = hvc ? __ldg(ZK + o) : 0.0f; u.kdiag = hdc ? __ldg(kinv_diag + col) : 0.0f;
        return u;
    }
    __device__ __forceinline__ float apply(int row, int col, float acc, const Aux& u) const {
        const long long o = (long long)row * ld + col;
        const float e = a * expf(u.x);
        const float g = u.counts - e - (acc - u.kinv_mu);
        const bool in = !(clip > 0.0f) || (g > -clip && g < clip);
        gc[o] = clip > 0.0f ? fminf(fmaxf(g, -clip), clip) : g;
        if (hvc) hvc[o] = in ? (-e * u.z - u.zk) : 0.0f;
        if (hdc) hdc[o] = in ? (-e - u.kdiag) : 0.0f;
        return 0.0f;
    }
    __device__ __forceinline__ float operator()(int row, int col, float acc) const { return apply(row, col, acc, load(row, col)); }
    struct Col4 { float4 counts, kinv_mu, kdiag; };
    struct Row4 { float4 x, z, zk; };
    bool vec_ok() const {
        return aligned16(X) && ldx % 4 == 0 && ld % 4 == 0 && aligned16(counts) && aligned16(kinv_mu) && aligned16(gc) &&
               (!hvc || (aligned16(hvc) && aligned16(Z) && aligned16(ZK))) && (!hdc || (aligned16(hdc) && aligned16(kinv_diag)));
    }
    __device__ __forceinline__ Col4 load_col4(int col) const {
        Col4 c; c.counts = ldg4(counts + col); c.kinv_mu = ldg4(kinv_mu + col); c.kdiag = hdc ? ldg4(kinv_diag + col) : f4(0.0f);
        return c;
    }
    __device__ __forceinline__ Row4 load_row4(int row, int col) const {
        const long long o = (long long)row * ld + col;
        Row4 r; r.x = ldg4(X + (long long)row * ldx + col);
        r.z = hvc ? ldg4(Z + o) : f4(0.0f); r.zk = hvc ? ldg4(ZK + o) : f4(0.0f);
        return r;
    }
    __device__ __forceinline__ void one(float acc, float x, float cnt, float kmu, float kd, float z, float zk,
                                        float& g_out, float& hv_out, float& hd_out) const {
        const float e = a * expf(x);
        const float g = cnt - e - (acc - kmu);
        const bool in = !(clip > 0.0f) || (g > -clip && g < clip);
        g_out = clip > 0.0f ? fminf(fmaxf(g, -clip), clip) : g;
        hv_out = in ? (-e * z - zk) : 0.0f;
        hd_out = in ? (-e - kd) : 0.0f;
    }
    __device__ __forceinline__ float apply4(int row, int col, const float4& acc, const Col4& c, const Row4& r) const {
        const long long o = (long long)row * ld + col;
        float4 g, hv, hd;
        one(acc.x, r.x.x, c.counts.x, c.kinv_mu.x, c.kdiag.x, r.z.x, r.zk.x, g.x, hv.x, hd.x);
        one(acc.y, r.x.y, c.counts.y, c.kinv_mu.y, c.kdiag.y, r.z.y, r.zk.y, g.y, hv.y, hd.y);
        one(acc.z, r.x.z, c.counts.z, c.kinv_mu.z, c.kdiag.z, r.z.z, r.zk.z, g.z, hv.z, hd.z);
        one(acc.w, r.x.w, c.counts.w, c.kinv_mu.w, c.kdiag.w, r.z.w, r.zk.w, g.w, hv.w, hd.w);
        st4(gc + o, g);
        if (hvc) st4(hvc + o, hv);
        if (hdc) st4(hdc + o, hd);
        return 0.0f;
    }
    __device__ __forceinline__ void row_partial(int, int, float) const {}
    __device__ __forceinline__ void at_z(int) {}
};

int pines_n_tiles(int d) { return gemm_n_tiles(d); }

// the constant K^-1 arrives pre-split (scaled fp16 parts + its maximum behind them) when the host built the mirror once
static inline void kinv_mirror(const mfm_target_t& T, GemmShape& p) {
    if (T.kinv_split && tc2h::gemm_h16()) { p.b_mirror = T.kinv_split; p.b_amax = T.kinv_split + (long long)T.dim * T.dim; }
}

int pines_grad_gemm(const mfm_target_t& T, int n, const float* X, long long ldx, float beta, float* grad_out,
                    long long ldg, float* prior_partial, const int* n_rows_dev, cudaStream_t st, const float* x_amax) {
    GemmShape p{n, T.dim, T.dim, X, ldx, T.kinv, (long long)T.dim, n_rows_dev};
    p.a_amax = x_amax; kinv_mirror(T, p);
    EpiPinesGrad e{X, ldx, T.counts, T.kinv_mu, T.mu, T.poisson_a, beta, grad_out, ldg, prior_partial, gemm_n_tiles(T.dim)};
    MFM_CUDA_CHECK((launch_gemm<true, false>(p, e, st)   /* K^-1 is symmetric: read it as the K-major operand */));
    return MFM_OK;
}

int pines_kinv_gemm(const mfm_target_t& T, int n, const float* Z, long long ldz, float* out, long long ldo,
                    const int* n_rows_dev, cudaStream_t st, float z_bound) {
    GemmShape p{n, T.dim, T.dim, Z, ldz, T.kinv, (long long)T.dim, n_rows_dev};
    p.a_bound = z_bound; kinv_mirror(T, p);
    EpiStd e{out, ldo, nullptr, nullptr, 0, nullptr, 0, 1.0f, 0};
    MFM_CUDA_CHECK((launch_gemm<true, false>(p, e, st)   /* K^-1 is symmetric: read it as the K-major operand */));
    return MFM_OK;
}

// loglik[n] = sum(x c - a e^x)   (cox_process_utils.py:98-115); one warp per chain
__global__ void pines_loglik_kernel(mfm_target_t T, int n, const float* __restrict__ x, float* __restrict__ lik) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    float s = 0.0f;
    for (int i = lane; i < T.dim; i += 32) {
        const float xv = x[(long long)c * T.dim + i];
        s += xv * T.counts[i] - T.poisson_a * expf(xv);
    }
    s = warp_sum(s);
    if (lane == 0) lik[c] = s;
}

__global__ void pines_finish_value_kernel(mfm_target_t T, int n, int n_tiles, const float* __restrict__ lik,
                                          const float* __restrict__ partial, float* __restrict__ logp,
                                          float* __restrict__ loglik_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float s = 0.0f;
    for (int t = 0; t < n_tiles; ++t) s += partial[(long long)c * n_tiles + t];
    logp[c] = T.beta * lik[c] + (-0.5f * s + T.log_norm);
    if (loglik_out) loglik_out[c] = lik[c];
}

// ---- whitened pines (distributions.py:276-297): state e, latents f = L e + mu --------------------------------------------------
// per row: lik = sum(f c - a e^f), r = beta (c - a e^f) (the loglik gradient w.r.t. f), prior = -|e|^2 / 2, neg_e = -e
__global__ void __launch_bounds__(256)
white_lik_kernel(mfm_target_t T, int n, const float* __restrict__ e, const float* __restrict__ f, float* __restrict__ r,
                 float* __restrict__ neg_e, float* __restrict__ lik, float* __restrict__ prior) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    float sl = 0.0f, sp = 0.0f;
    for (int i = lane; i < T.dim; i += 32) {
        const long long o = (long long)c * T.dim + i;
        const float fv = f[o], ev = e[o], ex = T.poisson_a * expf(fv);
        sl += fv * T.counts[i] - ex;
        sp += ev * ev;
        r[o] = T.beta * (T.counts[i] - ex);
        neg_e[o] = -ev;
    }
    sl = warp_sum(sl); sp = warp_sum(sp);
    if (lane == 0) { lik[c] = sl; prior[c] = -0.5f * sp; }
}
__global__ void white_finish_kernel(mfm_target_t T, int n, const float* __restrict__ lik, const float* __restrict__ prior,
                                    float* __restrict__ logp, float* __restrict__ loglik_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    logp[c] = T.beta * lik[c] + (prior[c] + T.log_norm);
    if (loglik_out) loglik_out[c] = lik[c];
}
// field terms: g = (c - a e^f) L - e clipped; hv = -((a e^f o (z L^T)) L) - z; hd_j = -sum_i L_ij^2 a e^f_i - 1, both zeroed where g was clipped
__global__ void white_weights_kernel(mfm_target_t T, long long total, const float* __restrict__ f, const float* __restrict__ q,
                                     float* __restrict__ r, float* __restrict__ s) {
    const long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (o >= total) return;
    const int i = (int)(o % T.dim);
    const float ex = T.poisson_a * expf(f[o]);
    r[o] = T.counts[i] - ex;
    if (s) s[o] = q ? ex * q[o] : ex;
}
__global__ void white_field_finish_kernel(long long total, float clip, const float* __restrict__ g_lin, const float* __restrict__ e,
                                          const float* __restrict__ h_lin, const float* __restrict__ z, float* __restrict__ gc,
                                          float* __restrict__ hvc, float* __restrict__ hdc) {
    const long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (o >= total) return;
    const float g = g_lin[o] - e[o];
    const bool in = !(clip > 0.0f) || (g > -clip && g < clip);
    gc[o] = clip > 0.0f ? fminf(fmaxf(g, -clip), clip) : g;
    if (hvc) hvc[o] = in ? (-h_lin[o] - z[o]) : 0.0f;
    if (hdc) hdc[o] = in ? (-h_lin[o] - 1.0f) : 0.0f;
}
static int white_gemm(int n, int d, const float* A, const float* Bt, const float* bias, const float* add, float* C, cudaStream_t st) {
    GemmShape p{n, d, d, A, (long long)d, Bt, (long long)d, nullptr};
    EpiStd e{C, (long long)d, bias, nullptr, 0, add, (long long)d, 1.0f, 0};
    MFM_CUDA_CHECK((launch_gemm<true, false>(p, e, st)));
    return MFM_OK;
}

size_t target_ws_bytes(const mfm_target_t& T, int n) {
    if (T.kind == MFM_TARGET_PINES_WHITE) return 3 * ws_slice((size_t)n * T.dim, 4) + 2 * ws_slice(n, 4) + 256;
    if (T.kind != MFM_TARGET_PINES) return 256;
    return ws_slice((size_t)n * pines_n_tiles(T.dim), 4) + ws_slice(n, 4) + 256;
}

int target_value_and_grad(const mfm_target_t& T, int n, const float* x, float* logp, float* grad, float* loglik_out,
                          Workspace& ws, cudaStream_t st) {
    if (n <= 0) return MFM_OK;
    if (T.kind == MFM_TARGET_PINES_WHITE) {
        const int d = T.dim;
        float* f = ws.take<float>((size_t)n * d); float* r = ws.take<float>((size_t)n * d); float* neg_e = ws.take<float>((size_t)n * d);
        float* lik = ws.take<float>(n); float* prior = ws.take<float>(n);
        if (!ws.ok) { mfm_set_last_error_msg("workspace too small (target_value_and_grad)"); return MFM_ERR_WORKSPACE; }
        int rc;
        if ((rc = white_gemm(n, d, x, T.chol, T.mu_vec, nullptr, f, st))) return rc;             // f = e L^T + mu   (cox_process_utils.py:137)
        white_lik_kernel<<<ceil_div(n, 8), 256, 0, st>>>(T, n, x, f, r, neg_e, lik, prior);
        MFM_LAUNCH_CHECK();
        if ((rc = white_gemm(n, d, r, T.chol_t, nullptr, neg_e, grad, st))) return rc;           // grad = beta (c - a e^f) L - e
        white_finish_kernel<<<ceil_div(n, 256), 256, 0, st>>>(T, n, lik, prior, logp, loglik_out);
        MFM_LAUNCH_CHECK();
        return MFM_OK;
    }
    if (T.kind == MFM_TARGET_PINES) {
        const int nt = pines_n_tiles(T.dim);
        float* partial = ws.take<float>((size_t)n * nt);
        float* lik = ws.take<float>(n);
        if (!ws.ok) { mfm_set_last_error_msg("workspace too small (target_value_and_grad)"); return MFM_ERR_WORKSPACE; }
        pines_loglik_kernel<<<ceil_div(n, 8), 256, 0, st>>>(T, n, x, lik);
        MFM_LAUNCH_CHECK();
        int rc = pines_grad_gemm(T, n, x, T.dim, T.beta, grad, T.dim, partial, nullptr, st);
        if (rc) return rc;
        pines_finish_value_kernel<<<ceil_div(n, 256), 256, 0, st>>>(T, n, nt, lik, partial, logp, loglik_out);
        MFM_LAUNCH_CHECK();
        return MFM_OK;
    }
    if (T.kind == MFM_TARGET_GMM && T.dim != 2) { mfm_set_last_error_msg("GMM target requires dim == 2"); return MFM_ERR_ARG; }
    const size_t smem = (size_t)EVAL_WARPS * 2 * T.dim * sizeof(float);
    if (smem > 200 * 1024) { mfm_set_last_error_msg("dim too large for warp-per-chain target"); return MFM_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) MFM_CUDA_CHECK(cudaFuncSetAttribute(small_value_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    small_value_grad_kernel<<<ceil_div(n, EVAL_WARPS), EVAL_WARPS * 32, smem, st>>>(T, n, x, logp, grad, loglik_out);
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

int target_field_terms(const mfm_target_t& T, int n, const float* x, const float* z, const float* zkinv, float clip,
                       float* gc, float* hvc, float* hdc, const int* n_rows_dev, cudaStream_t st, const float* x_amax, float* scratch,
                       const float* x_split) {
    if (n <= 0) return MFM_OK;
    if (T.kind == MFM_TARGET_PINES_WHITE) {
        // scratch: 4 x [n, d] (f, r / g_lin, q / s, h_lin).  Rows beyond a device-side active count are computed too (harmless).
        if (!scratch) { mfm_set_last_error_msg("internal: whitened pines field terms need scratch"); return MFM_ERR_ARG; }
        const int d = T.dim; const long long tot = (long long)n * d; const size_t sl = ws_slice((size_t)tot, 4) / 4;
        float* f = scratch; float* r = scratch + sl; float* q = scratch + 2 * sl; float* hl = scratch + 3 * sl;
        int rc;
        if ((rc = white_gemm(n, d, x, T.chol, T.mu_vec, nullptr, f, st))) return rc;
        if (hvc && (rc = white_gemm(n, d, z, T.chol, nullptr, nullptr, q, st))) return rc;       // L z
        white_weights_kernel<<<ceil_div(tot, 256), 256, 0, st>>>(T, tot, f, hvc ? q : nullptr, r, (hvc || hdc) ? q : nullptr);
        MFM_LAUNCH_CHECK();
        if ((hvc || hdc) && (rc = white_gemm(n, d, q, hvc ? T.chol_t : T.chol_sq_t, nullptr, nullptr, hl, st))) return rc;
        if ((rc = white_gemm(n, d, r, T.chol_t, nullptr, nullptr, f, st))) return rc;            // (c - a e^f) L   (f is free now)
        white_field_finish_kernel<<<ceil_div(tot, 256), 256, 0, st>>>(tot, clip, f, x, hl, z, gc, hvc, hdc);
        MFM_LAUNCH_CHECK();
        return MFM_OK;
    }
    if (T.kind == MFM_TARGET_PINES) {
        GemmShape p{n, T.dim, T.dim, x, (long long)T.dim, T.kinv, (long long)T.dim, n_rows_dev};
        p.a_amax = x_amax; kinv_mirror(T, p);
        if (x_split && x_amax) { p.a_split = x_split; p.a_scale_src = x_amax; }      // the caller's pre-split copy of x
        EpiPinesField e{x, (long long)T.dim, T.counts, T.kinv_mu, T.kinv_diag, z, zkinv, T.poisson_a, clip, gc, hvc, hdc, (long long)T.dim};
        MFM_CUDA_CHECK((launch_gemm<true, false>(p, e, st)   /* K^-1 is symmetric: read it as the K-major operand */));
        return MFM_OK;
    }
    const size_t smem = (size_t)EVAL_WARPS * 5 * T.dim * sizeof(float);
    if (smem > 200 * 1024) { mfm_set_last_error_msg("dim too large for warp-per-chain target"); return MFM_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) MFM_CUDA_CHECK(cudaFuncSetAttribute(small_field_terms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    small_field_terms_kernel<<<ceil_div(n, EVAL_WARPS), EVAL_WARPS * 32, smem, st>>>(T, n, n_rows_dev, clip, x, z, gc, hvc, hdc);
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

}  // namespace mfm

extern "C" {

/* Host-side copy of the persistent kernel's work list (gemm_tcgen05_persist.cuh::Sched), for tests without a GPU:
 * rows of 7 ints (pair, tile, kb0, kb1, kind, c_first, c_count); returns the number of rows (<= cap written). */
int mfm_debug_gemm_plan(int M, int N, int K, int n_pairs, int streamk, int* rows, int cap) {
    using namespace mfm::tc2p;
    const Tiles T = make_tiles(M, N, K, 0);
    int n = 0;
    for (int pair = 0; pair < n_pairs; ++pair) {
        const Sched S = make_sched(T.total, n_pairs, pair, K, 0, streamk != 0);
        Item it;
        for (int idx = 0; S.get(idx, it); ++idx, ++n)
            if (n < cap) { int* r = rows + 7 * n; r[0] = pair; r[1] = it.tile; r[2] = it.kb0; r[3] = it.kb1; r[4] = it.kind; r[5] = it.c_first; r[6] = it.c_count; }
    }
    return n;
}

int mfm_gemm_tf32x3(int M, int N, int K, const float* A, long long lda, int a_kmajor, const float* B, long long ldb,
                    int b_nmajor, const float* bias, int relu, float* Cout, long long ldc, mfm_stream_t stream) {
    using namespace mfm;
    GemmShape p{M, N, K, A, lda, B, ldb, nullptr};
    EpiStd e{Cout, ldc, bias, nullptr, 0, nullptr, 0, 1.0f, relu};
    cudaError_t err;
    if (a_kmajor && b_nmajor) err = launch_gemm<true, true>(p, e, stream);
    else if (a_kmajor && !b_nmajor) err = launch_gemm<true, false>(p, e, stream);
    else if (!a_kmajor && b_nmajor) err = launch_gemm<false, true>(p, e, stream);
    else err = launch_gemm<false, false>(p, e, stream);
    MFM_CUDA_CHECK(err);
    return MFM_OK;
}

int mfm_gemm_tf32x3_gated(int M, int N, int K, const float* A, long long lda, const float* Bt, long long ldb, const float* mask,
                          long long ldm, const float* add, long long ldadd, float* Cout, long long ldc, mfm_stream_t stream) {
    using namespace mfm;
    GemmShape p{M, N, K, A, lda, Bt, ldb, nullptr};
    EpiStd e{Cout, ldc, nullptr, mask, ldm, add, ldadd, 1.0f, 0};
    MFM_CUDA_CHECK((launch_gemm<true, false>(p, e, stream)));
    return MFM_OK;
}

int mfm_gemm_presplit(const float* src, float* mirror, long long n_floats, mfm_stream_t stream) {
    const bool h16 = mfm::tc2h::gemm_h16() != 0;
    if (!src || !mirror || n_floats <= 0 || n_floats % (h16 ? 16 : 8) || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(mirror)) & (h16 ? 63 : 31))) {
        mfm_set_last_error_msg("bad argument (mfm_gemm_presplit)"); return MFM_ERR_ARG;
    }
    float* amax = mirror + n_floats;          // the mirror buffer holds n_floats + 16 floats: max |src| lives behind the parts
    if (h16) MFM_CUDA_CHECK(mfm::tc2h::launch_absmax(src, n_floats, 1, (int)n_floats, nullptr, amax, stream));
    return mfm::presplit_weights(src, mirror, n_floats, h16 ? amax : nullptr, stream);
}

int mfm_absmax(const float* x, long long ld, int rows, int cols, float* out, mfm_stream_t stream) {
    if (!x || !out || rows <= 0 || cols <= 0 || cols % 4 || ld % 4 || (reinterpret_cast<uintptr_t>(x) & 15)) { mfm_set_last_error_msg("bad argument (mfm_absmax)"); return MFM_ERR_ARG; }
    MFM_CUDA_CHECK(mfm::tc2h::launch_absmax(x, ld, rows, cols, nullptr, out, stream));
    return MFM_OK;
}

int mfm_gemm_dense(int M, int N, int K, const float* A, long long lda, const float* Bt, long long ldb, const float* bias, int relu,
                   float* Cout, long long ldc, const float* a_amax, float* c_amax, const float* a_split, const float* a_scale_src,
                   mfm_stream_t stream) {
    using namespace mfm;
    GemmShape p{M, N, K, A, lda, Bt, ldb, nullptr};
    p.a_amax = a_amax; p.a_split = a_split; p.a_scale_src = a_scale_src;
    EpiStd e{Cout, ldc, bias, nullptr, 0, nullptr, 0, 1.0f, relu};
    e.amax_out = c_amax;
    MFM_CUDA_CHECK((launch_gemm<true, false>(p, e, stream)));
    return MFM_OK;
}

/* Test hook of the pre-split hand-over between two dense layers (not in the ABI header):
 *   C1 = relu(A B1t^T + b1), written as fp32 AND pre-split (C1s) by the first layer's epilogue with the scale of its output bound;
 *   C2 = C1 B2t^T, whose A operand is loaded pre-split.  slots: device float[4] = {max|A|, max|C1| (exact), bound(C1), -}. */
int mfm_debug_dense_chain(int M, int K, int N1, int N2, const float* A, const float* B1t, const float* b1, const float* B2t,
                          float* C1, float* C1s, float* C2, const float* wnorm1, const float* bias_amax, float* slots, mfm_stream_t stream) {
    using namespace mfm;
    MFM_CUDA_CHECK(tc2h::launch_absmax(A, K, M, K, nullptr, slots, stream));
    MFM_CUDA_CHECK(cudaMemsetAsync(slots + 1, 0, 3 * sizeof(float), stream));
    {
        GemmShape p{M, N1, K, A, (long long)K, B1t, (long long)K, nullptr};
        p.a_amax = slots;
        EpiStdS e{C1, (long long)N1, b1, nullptr, 0, nullptr, 0, 1, 1, C1s, slots, nullptr, 0.0f, wnorm1, bias_amax, nullptr, slots + 2};
        e.amax_out = slots + 1;
        MFM_CUDA_CHECK((launch_gemm<true, false>(p, e, stream)));
    }
    {
        GemmShape p{M, N2, N1, C1, (long long)N1, B2t, (long long)N1, nullptr};
        p.a_amax = slots + 1; p.a_split = C1s; p.a_scale_src = slots + 2;
        EpiStd e{C2, (long long)N2, nullptr, nullptr, 0, nullptr, 0, 1.0f, 0};
        MFM_CUDA_CHECK((launch_gemm<true, false>(p, e, stream)));
    }
    return MFM_OK;
}

int mfm_gemm_tf32x3_rows(int M, int N, int K, const float* A, long long lda, const float* Bt, long long ldb, const float* bias,
                         float* Cout, long long ldc, const int* n_rows_dev, mfm_stream_t stream) {
    using namespace mfm;
    GemmShape p{M, N, K, A, lda, Bt, ldb, n_rows_dev};
    EpiStd e{Cout, ldc, bias, nullptr, 0, nullptr, 0, 1.0f, 0};
    MFM_CUDA_CHECK((launch_gemm<true, false>(p, e, stream)));
    return MFM_OK;
}

size_t mfm_target_workspace_bytes(const mfm_target_t* t, int n) { return mfm::target_ws_bytes(*t, n); }

int mfm_logdensity_and_grad(const mfm_target_t* t, int n, const float* x, float* logp, float* grad, float* loglik_out,
                            void* ws, size_t ws_bytes, mfm_stream_t stream) {
    if (!t || !x || !logp || !grad || n < 0) { mfm_set_last_error_msg("null argument"); return MFM_ERR_ARG; }
    mfm::Workspace w(ws, ws_bytes);
    return mfm::target_value_and_grad(*t, n, x, logp, grad, loglik_out, w, stream);
}

}  // extern "C"
