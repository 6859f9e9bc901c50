// Internal (non-ABI) declarations shared between translation units.
#pragma once
#include "common.cuh"
#include "../../include/mfm_b200.h"

void mfm_set_last_error_msg(const char* msg);

namespace mfm {

// bump allocator over the caller's workspace (256-byte aligned slices)
struct Workspace {
    char* base; size_t size, off; bool ok;
    Workspace(void* p, size_t n) : base((char*)p), size(n), off(0), ok(true) {}
    template <class T> T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        if (base == nullptr || off + bytes > size) { ok = false; off += bytes; return nullptr; }
        T* r = (T*)(base + off); off += bytes; return r;
    }
};
inline size_t ws_slice(size_t count, size_t elt) { return (count * elt + 255) & ~(size_t)255; }

int pines_n_tiles(int d);

// grad_out[n,d] = beta*(c - a e^x) - (x-mu) K^-1 ; prior_partial[n, pines_n_tiles] = per-tile
// sums of (x-mu).q  (logprior = -0.5*sum + log_norm).  n_rows_dev: optional active-row count.
int pines_grad_gemm(const mfm_target_t& T, int n, const float* X, long long ldx, float beta,
                    float* grad_out, long long ldg, float* prior_partial, const int* n_rows_dev,
                    cudaStream_t st);
// out[n,d] = Z K^-1 (Hessian-vector product of the prior is -Z K^-1)
int pines_kinv_gemm(const mfm_target_t& T, int n, const float* Z, long long ldz, float* out, long long ldo,
                    const int* n_rows_dev, cudaStream_t st);

// value+grad of logprob_beta for any target; loglik_out optional; ws from target_ws_bytes
size_t target_ws_bytes(const mfm_target_t& T, int n);
int target_value_and_grad(const mfm_target_t& T, int n, const float* x, float* logp, float* grad,
                          float* loglik_out, Workspace& ws, cudaStream_t st);

// untempered grad logprob (+ optional Hessian-vector product with z / Hessian diagonal) for the
// vector field; for pines hv excludes the constant -zK^-1 part when zkinv is supplied.
// outputs: gc = clip(grad), hvc = 1[|grad|<clip] * (H z), hdc = 1[|grad|<clip] * diag(H)  (hvc/hdc optional)
int target_field_terms(const mfm_target_t& T, int n, const float* x, const float* z, const float* zkinv, float clip,
                       float* gc, float* hvc, float* hdc, const int* n_rows_dev, cudaStream_t st);

// ---- vector field (flow.cu) -------------------------------------------------------------------
struct FieldBufs {
    float *ff, *h0, *cat, *h2, *gt, *h5, *h6, *gc, *hx, *ta, *tb, *zw2, *zkinv, *divpart;
    float *tan_a, *tan_b;   // exact path [n*d, H]
    float* wt;              // transposed dense kernels: layer i at wt + F.w_off[i], stored [out][in] (K-major B operand)
    float *wx, *wxo;        // pre-split bf16 cross mirrors of wt / of the parameters themselves (null: layer sizes not multiples of 8)
};
namespace tc2p {
void register_cross(const float* base, size_t n_floats, const float* mirror);
void clear_cross();
}
// dst mirrors src (n8 groups of 8 floats): group g -> 8 bf16 of the values | 8 bf16 of their tf32 truncation rests
int presplit_weights(const float* src, float* dst, long long n_floats, cudaStream_t st);
// clears the mirror registry when an ABI call that registered mirrors returns
struct CrossScope { ~CrossScope() { tc2p::clear_cross(); } };
size_t field_bufs_bytes(const mfm_field_t& F, const mfm_target_t& T, int n, bool hutch);
bool field_bufs_take(FieldBufs& B, Workspace& w, const mfm_field_t& F, int n, bool hutch);
// B.wt <- transposes of the eight dense kernels (once per ABI call: the parameters may have changed)
int field_prepare_weights(const mfm_field_t& F, FieldBufs& B, cudaStream_t st);
void field_register_mirrors(const mfm_field_t& F, const FieldBufs& B);
// C[n,out] = relu?(A[n,in] W[in,out] + bias) gated by mask (relu' of another activation).
// WT = the kernel transposed, WT[o*ldwt + i]: both GEMM operands are K-major, which is what the
// persistent tcgen05 kernel's bf16 cross-term path needs.
int dense(int n, int in, int out, const float* A, long long lda, const float* WT, long long ldwt, const float* bias, int relu,
          float* C, long long ldc, const float* mask, long long ldm, int mask_div, cudaStream_t st,
          const int* n_rows_dev = nullptr);
// out_v = sgn * v(x, t); out_l = -sgn * div v (optional; z != null -> Hutchinson, else exact trace).
// Leaves the activations (ff, h0, cat=[s_x|s_t], h2, gt, h5, h6, gc) in B for a backward pass.
// n_rows_dev (optional): device count of leading rows that are live (compacted active chains);
// row_map (optional): compact row -> chain index used when writing out_v / out_l.
int field_eval(const mfm_field_t& F, const mfm_target_t& T, int n, const float* x, const float* tfield,
               const float* z, float sgn, float* out_v, float* out_l, FieldBufs& B, cudaStream_t st,
               const int* n_rows_dev = nullptr, const int* row_map = nullptr);

}  // namespace mfm
