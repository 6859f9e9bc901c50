// Error-compensated TF32 GEMM (3xTF32: hi*hi + hi*lo + lo*hi, fp32 accumulate) with functor
// epilogues.  C[M,N] = opA(A)[M,K] * opB(B)[K,N].
//
// Why 3xTF32: the adaptive Dopri5 controller runs at rtol=atol=1e-5 and the parity contract is
// 1e-4 relative on log-densities / divergences / losses, so the field evaluation needs ~fp32
// accuracy.  A single TF32/BF16 pass (1e-3) would thrash the step controller.
//
// This is the portable warp-level (mma.sync.m16n8k8) implementation: 128x128x16 CTA tile,
// 8 warps (2x4, 64x32 warp tile), 3-stage cp.async pipeline.  The tcgen05/TMEM variant of the
// same arithmetic lives in gemm_tcgen05.cuh and is selected for large aligned shapes.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace mfm {

constexpr int GBM = 128, GBN = 64, GBK = 16, GSTAGES = 3, GTHREADS = 256;
constexpr int G_MI = 2, G_NJ = 4;      // warp tile 32x32 = 2 (m16) x 4 (n8) MMA tiles; warps 4 (M) x 2 (N)
constexpr int G_FLUSH = 2;             // k-tiles (x16) accumulated inside the tensor core before an RN flush
// smem strides chosen so fragment loads are bank-conflict free (see DESIGN.md "GEMM").
constexpr int LDS_K = GBK + 4;         // operand stored [rows][K]   (K contiguous)
constexpr int LDS_AM = GBM + 8;        // A stored [K][M]
constexpr int LDS_BN = GBN + 8;        // B stored [K][N]
constexpr int A_STAGE = (GBM * LDS_K > GBK * LDS_AM) ? GBM * LDS_K : GBK * LDS_AM;
constexpr int B_STAGE = (GBN * LDS_K > GBK * LDS_BN) ? GBN * LDS_K : GBK * LDS_BN;
constexpr int GEMM_SMEM_BYTES = GSTAGES * (A_STAGE + B_STAGE) * 4 + GBM * 2 * 4;

struct GemmShape {
    int M, N, K;
    const float* A; long long lda;   // A_KMAJOR: A[m*lda + k]; else A[k*lda + m]
    const float* B; long long ldb;   // B_NMAJOR: B[k*ldb + n]; else B[n*ldb + k]
    const int* n_rows_dev;           // optional: device int, effective M (active rows); tiles beyond exit
    int k_split = 0;                 // >0: gridDim.z slices of k_split values of k; epilogue gets slice via at_z()
    // magnitudes of the operands for the scaled fp16 split of gemm_tcgen05_h16.cuh (ignored by the other kernels):
    const float* a_amax = nullptr;   // device float: max |A| (maintained by A's producer), or null
    const float* a_amax2 = nullptr;  // second slot when A spans two producers' outputs (the larger one counts)
    float a_bound = 0.0f;            // > 0: bound on |A| known on the host (takes precedence)
    const float* b_amax = nullptr;   // device float: max |B|, or null (weight mirrors carry their own)
    const float* b_mirror = nullptr; // pre-split (split16, scaled fp16) copy of B made with h16_scale(*b_amax), or null: look one up
    const float* a_split = nullptr;  // pre-split copy of A (same shape / leading dimension, written by A's producer), or null
    const float* a_scale_src = nullptr; // device float: the copy was scaled by h16_scale(*a_scale_src) (the producer's bound on |A|)
};

// functors that track what they store expose `amax_out` (slot or null) and a per-thread running maximum `vmax`
template <class E> __device__ __forceinline__ auto epi_stored_max(const E& e, int) -> decltype(e.vmax) { return e.vmax; }
template <class E> __device__ __forceinline__ float epi_stored_max(const E&, long) { return 0.0f; }
template <class E> __device__ __forceinline__ auto epi_amax_slot(const E& e, int) -> decltype(e.amax_out) { return e.amax_out; }
template <class E> __device__ __forceinline__ float* epi_amax_slot(const E&, long) { return nullptr; }
// call with whole warps (uniform branch on the slot)
// the bound a split-writing functor scaled its copy with, published once per launch for the consumer (call from whole warps)
template <class E> __device__ __forceinline__ auto epi_publish_bound(const E& e, int) -> decltype((void)e.bound_out) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0 && e.bound_out) *e.bound_out = e.out_bound();
}
template <class E> __device__ __forceinline__ void epi_publish_bound(const E&, long) {}
template <class E, class = void> struct epi_has_colsum { static constexpr bool value = false; };
template <class E> struct epi_has_colsum<E, decltype((void)E::kColSum)> { static constexpr bool value = E::kColSum; };
template <class E> __device__ __forceinline__ void epi_publish_amax(const E& e, float vmax) {
    float* slot = epi_amax_slot(e, 0);
    if (slot != nullptr) amax_publish_warp(slot, vmax);
    epi_publish_bound(e, 0);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x)); return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Load a [ROWS x GBK] operand tile.  KMAJ=true: global is [row][k] (k contiguous) -> smem [row][LDS_K].
// KMAJ=false: global is [k][row] (row contiguous) -> smem [k][LDS_R].
template <bool KMAJ, int ROWS, int LDS_R>
__device__ __forceinline__ void load_tile(float* s, const float* __restrict__ g, long long ld, int row0, int k0,
                                          int rows, int K, bool vec_ok) {
    const int tid = threadIdx.x;
    constexpr int CHUNKS = ROWS * GBK / 4;
#pragma unroll
    for (int it = 0; it < (CHUNKS + GTHREADS - 1) / GTHREADS; ++it) {
        const int c = tid + it * GTHREADS;
        if (CHUNKS % GTHREADS != 0 && c >= CHUNKS) break;
        int r, kk, nr, nk;          // chunk origin (row, k) and its extent along (row, k)
        float* dst;
        if (KMAJ) { r = c >> 2; kk = (c & 3) * 4; nr = 1; nk = 4; dst = s + r * LDS_K + kk; }
        else      { kk = c / (ROWS / 4); r = (c % (ROWS / 4)) * 4; nr = 4; nk = 1; dst = s + kk * LDS_R + r; }
        const int gr = row0 + r, gk = k0 + kk;
        const bool full = (gr + nr <= rows) && (gk + nk <= K);
        const bool none = (gr >= rows) || (gk >= K);
        if (vec_ok && (full || none)) {
            const float* src = full ? (KMAJ ? g + (long long)gr * ld + gk : g + (long long)gk * ld + gr) : g;
            cp_async16(dst, src, full);
            continue;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int er = KMAJ ? gr : gr + e, ek = KMAJ ? gk + e : gk;
            dst[e] = (er < rows && ek < K) ? (KMAJ ? g[(long long)er * ld + ek] : g[(long long)ek * ld + er]) : 0.0f;
        }
    }
}

// Epilogue functor contract:
//   struct Aux;  __device__ Aux load(int row, int col) const;          // every global read of the epilogue
//   __device__ float apply(int row, int col, float acc, const Aux&) const;   // stores; returns row-sum contribution
//   __device__ float operator()(int row, int col, float acc) const;   // = apply(row, col, acc, load(row, col))
//   (the tcgen05 kernels issue the loads of several rows before the first store: memory-level parallelism)
//   static constexpr bool kRowSum;                                     // if true: row_partial(row, ntile, sum)
//   __device__ void row_partial(int row, int ntile, float s) const;
//
// Accumulation: the tensor core adds into its fp32 accumulator with truncation, which biases long
// chains (measured ~K*2^-25 relative).  MMAs therefore accumulate only G_FLUSH k-tiles (32 values of
// k, 12 MMAs) from zero and the partial is then added into fp32 registers with round-to-nearest.
template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
__global__ void __launch_bounds__(GTHREADS, 2) gemm_tf32x3_kernel(GemmShape p, Epi epi) {
    extern __shared__ __align__(16) float gsm[];
    float* sA = gsm;
    float* sB = gsm + GSTAGES * A_STAGE;
    float* sRed = gsm + GSTAGES * (A_STAGE + B_STAGE);   // [GBM][2]

    const int M = p.n_rows_dev ? min(*p.n_rows_dev, p.M) : p.M;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    if (m0 >= M) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int g = lane >> 2, t = lane & 3;

    const bool a_vec = ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) && (p.lda % 4 == 0);
    const bool b_vec = ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0) && (p.ldb % 4 == 0);

    float acc[G_MI][G_NJ][4], part[G_MI][G_NJ][4];
#pragma unroll
    for (int i = 0; i < G_MI; ++i)
#pragma unroll
        for (int j = 0; j < G_NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) { acc[i][j][e] = 0.0f; part[i][j][e] = 0.0f; }

    const int kz0 = p.k_split > 0 ? blockIdx.z * p.k_split : 0;
    const int Kend = p.k_split > 0 ? min(p.K, kz0 + p.k_split) : p.K;
    const int KT = (Kend - kz0 + GBK - 1) / GBK;
    auto issue = [&](int kt) {
        if (kt < KT) {
            const int s = kt % GSTAGES;
            load_tile<A_KMAJOR, GBM, LDS_AM>(sA + s * A_STAGE, p.A, p.lda, m0, kz0 + kt * GBK, M, Kend, a_vec);
            load_tile<!B_NMAJOR, GBN, LDS_BN>(sB + s * B_STAGE, p.B, p.ldb, n0, kz0 + kt * GBK, p.N, Kend, b_vec);
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        issue(kt + GSTAGES - 1);
        const float* a = sA + (kt % GSTAGES) * A_STAGE;
        const float* b = sB + (kt % GSTAGES) * B_STAGE;
#pragma unroll
        for (int ks = 0; ks < GBK; ks += 8) {
            uint32_t ah[G_MI][4], al[G_MI][4], bh[G_NJ][2], bl[G_NJ][2];
#pragma unroll
            for (int i = 0; i < G_MI; ++i) {
                const int r = wm + i * 16 + g;
                float v0, v1, v2, v3;
                if (A_KMAJOR) {
                    v0 = a[r * LDS_K + ks + t];       v1 = a[(r + 8) * LDS_K + ks + t];
                    v2 = a[r * LDS_K + ks + t + 4];   v3 = a[(r + 8) * LDS_K + ks + t + 4];
                } else {
                    v0 = a[(ks + t) * LDS_AM + r];     v1 = a[(ks + t) * LDS_AM + r + 8];
                    v2 = a[(ks + t + 4) * LDS_AM + r]; v3 = a[(ks + t + 4) * LDS_AM + r + 8];
                }
                split_tf32(v0, ah[i][0], al[i][0]); split_tf32(v1, ah[i][1], al[i][1]);
                split_tf32(v2, ah[i][2], al[i][2]); split_tf32(v3, ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < G_NJ; ++j) {
                const int c = wn + j * 8 + g;
                float v0, v1;
                if (B_NMAJOR) { v0 = b[(ks + t) * LDS_BN + c]; v1 = b[(ks + t + 4) * LDS_BN + c]; }
                else          { v0 = b[c * LDS_K + ks + t];    v1 = b[c * LDS_K + ks + t + 4]; }
                split_tf32(v0, bh[j][0], bl[j][0]); split_tf32(v1, bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int i = 0; i < G_MI; ++i)
#pragma unroll
                for (int j = 0; j < G_NJ; ++j) {
                    mma_tf32(part[i][j], al[i], bh[j]);
                    mma_tf32(part[i][j], ah[i], bl[j]);
                    mma_tf32(part[i][j], ah[i], bh[j]);
                }
        }
        if ((kt % G_FLUSH) == G_FLUSH - 1 || kt == KT - 1) {
#pragma unroll
            for (int i = 0; i < G_MI; ++i)
#pragma unroll
                for (int j = 0; j < G_NJ; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) { acc[i][j][e] += part[i][j][e]; part[i][j][e] = 0.0f; }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ----
    if (p.k_split > 0) epi.at_z(blockIdx.z);
#pragma unroll
    for (int i = 0; i < G_MI; ++i) {
        float rs0 = 0.0f, rs1 = 0.0f;
        const int r0 = m0 + wm + i * 16 + g, r1 = r0 + 8;
#pragma unroll
        for (int j = 0; j < G_NJ; ++j) {
            const int c0 = n0 + wn + j * 8 + 2 * t;
            if (r0 < M) {
                if (c0 < p.N) rs0 += epi(r0, c0, acc[i][j][0]);
                if (c0 + 1 < p.N) rs0 += epi(r0, c0 + 1, acc[i][j][1]);
            }
            if (r1 < M) {
                if (c0 < p.N) rs1 += epi(r1, c0, acc[i][j][2]);
                if (c0 + 1 < p.N) rs1 += epi(r1, c0 + 1, acc[i][j][3]);
            }
        }
        if (Epi::kRowSum) {
            rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
            rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
            if (t == 0) {
                sRed[(wm + i * 16 + g) * 2 + (warp & 1)] = rs0;
                sRed[(wm + i * 16 + g + 8) * 2 + (warp & 1)] = rs1;
            }
        }
    }
    if (Epi::kRowSum) {
        __syncthreads();
        if (tid < GBM && m0 + tid < M) epi.row_partial(m0 + tid, blockIdx.x, sRed[tid * 2 + 0] + sRed[tid * 2 + 1]);
    }
    epi_publish_amax(epi, epi_stored_max(epi, 0));
}

template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
inline cudaError_t launch_gemm_mma(const GemmShape& p, const Epi& epi, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0) return cudaSuccess;
    auto kern = gemm_tf32x3_kernel<A_KMAJOR, B_NMAJOR, Epi>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((p.N + GBN - 1) / GBN, (p.M + GBM - 1) / GBM, p.k_split > 0 ? (p.K + p.k_split - 1) / p.k_split : 1);
    kern<<<grid, GTHREADS, GEMM_SMEM_BYTES, st>>>(p, epi);
    ++g_mfm_launches;
    return cudaGetLastError();
}

inline int gemm_n_tiles(int N) { return (N + GBN - 1) / GBN; }

// ---- helpers of the 4-column (float4) epilogue interface used by the CTA-pair kernel ------------
//   struct Col4 / Row4;  Col4 load_col4(col): column-only reads (hoisted out of the row loop);
//   Row4 load_row4(row, col): per-row reads;  float apply4(row, col, acc4, Col4, Row4): stores columns
//   col..col+3 of one row and returns their row-sum contribution;  bool vec_ok() (host): every
//   pointer 16-byte aligned and every leading dimension a multiple of 4.
__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(float v) { return make_float4(v, v, v, v); }

// ---- activations (exe_flow_matching.py:40-46 `non_lins`: jax.nn.tanh / elu / relu / gelu / swish) --------------------------
// code: 0 none, 1 relu, 2 tanh, 3 elu (alpha = 1), 4 gelu (jax.nn.gelu default: tanh approximation), 5 swish (x sigmoid(x)).
// Returns act(v) and its derivative at v (the backward pass and the forward-mode tangents multiply by it; relu keeps the
// cheaper sign-of-the-output gate: relu'(0) = 0 as in jax.nn.relu's custom JVP).
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_ELU = 3, ACT_GELU = 4, ACT_SWISH = 5 };
__device__ __forceinline__ float act_fwd(int act, float v, float& dv) {
    switch (act) {
        case ACT_RELU: dv = v > 0.0f ? 1.0f : 0.0f; return fmaxf(v, 0.0f);
        case ACT_TANH: { const float t = tanhf(v); dv = 1.0f - t * t; return t; }
        case ACT_ELU: { const float e = expm1f(v); dv = v > 0.0f ? 1.0f : e + 1.0f; return v > 0.0f ? v : e; }
        case ACT_GELU: {
            const float c = 0.7978845608028654f, a = 0.044715f;
            const float u = c * (v + a * v * v * v), t = tanhf(u);
            dv = 0.5f * (1.0f + t) + 0.5f * v * (1.0f - t * t) * c * (1.0f + 3.0f * a * v * v);
            return 0.5f * v * (1.0f + t);
        }
        case ACT_SWISH: { const float sg = 1.0f / (1.0f + expf(-v)); dv = sg + v * sg * (1.0f - sg); return v * sg; }
        default: dv = 1.0f; return v;
    }
}

// ---- standard epilogue: C = mask(act(alpha*acc + bias) ) (+ add) -----------------------------
// ACT = false: relu / no activation and sign-of-the-output gates only (the default configuration: nothing of the general
// activations' code - transcendental functions, derivative stores, multiplicative gates - reaches the kernels' epilogues, which
// sit at the register cap); ACT = true: every activation of act_fwd.
// CS = true: the functor also accumulates per-thread column sums of what it stores (`cs`); the scaled-fp16 kernel reduces them
// over each 32-row block and writes one partial row per block to `colsum_part` (deterministic; bias gradients without a pass
// over the tensor).
template <bool ACT, bool CS = false>
struct EpiStdT {
    static constexpr bool kRowSum = false;
    static constexpr bool kColSum = CS;
    float* C; long long ldc;
    const float* bias;               // [N] or null
    const float* mask; long long ldm; // out = mask[row/mask_div,col] > 0 ? out : 0  (relu' gate) or null
    const float* add; long long ldadd; // out += add[row,col] or null
    float alpha; int relu;
    int mask_div = 1;                  // rows of the mask are shared by mask_div consecutive output rows
    long long c_zstride = 0;           // split-K: slice z writes to C + z*c_zstride
    float* amax_out = nullptr;         // optional device slot: max |C| (C is the A operand of a later h16 GEMM)
    mutable float vmax = 0.0f;         // per-thread running maximum of what this copy of the functor stored
    float* dact = nullptr; long long lddact = 0;   // optional: act'(pre-activation) is written here (activations other than relu)
    int mask_mul = 0;                  // 1: out *= mask (the mask operand holds activation derivatives) instead of the > 0 gate
    float* colsum_part = nullptr; long long ldcs = 0;      // CS: [ceil(M / 32)][ldcs] partial column sums
    mutable float4 cs = {0.0f, 0.0f, 0.0f, 0.0f};
    __device__ __forceinline__ void at_z(int z) { C += (long long)z * c_zstride; }
    struct Aux { float bias, add, mask; };
    __device__ __forceinline__ Aux load(int row, int col) const {
        Aux a;
        a.bias = bias ? __ldg(bias + col) : 0.0f;
        a.add = add ? add[(long long)row * ldadd + col] : 0.0f;      // may alias C (in-place accumulate): plain load
        a.mask = mask ? __ldg(mask + (long long)(mask_div == 1 ? row : row / mask_div) * ldm + col) : 1.0f;
        return a;
    }
    __device__ __forceinline__ float apply(int row, int col, float acc, const Aux& a) const {
        float v = alpha * acc + a.bias + a.add;
        if constexpr (ACT) {
            if (relu == ACT_RELU) v = fmaxf(v, 0.0f);
            else if (relu) { float dv; v = act_fwd(relu, v, dv); if (dact) dact[(long long)row * lddact + col] = dv; }
            v = mask_mul ? v * a.mask : (a.mask > 0.0f ? v : 0.0f);
        } else {
            if (relu) v = fmaxf(v, 0.0f);
            v = a.mask > 0.0f ? v : 0.0f;
        }
        C[(long long)row * ldc + col] = v;
        vmax = fmaxf(vmax, fabsf(v));
        return 0.0f;
    }
    __device__ __forceinline__ float operator()(int row, int col, float acc) const { return apply(row, col, acc, load(row, col)); }
    __device__ __forceinline__ void row_partial(int, int, float) const {}
    // float4 interface
    struct Col4 { float4 bias; };
    struct Row4 { float4 add, mask; };
    bool vec_ok() const {
        return aligned16(C) && ldc % 4 == 0 && c_zstride % 4 == 0 && (!bias || aligned16(bias)) &&
               (!mask || (aligned16(mask) && ldm % 4 == 0)) && (!add || (aligned16(add) && ldadd % 4 == 0));
    }
    __device__ __forceinline__ Col4 load_col4(int col) const { Col4 c; c.bias = bias ? ldg4(bias + col) : f4(0.0f); return c; }
    __device__ __forceinline__ Row4 load_row4(int row, int col) const {
        Row4 r;
        r.add = add ? ld4(add + (long long)row * ldadd + col) : f4(0.0f);        // may alias C: plain load
        r.mask = mask ? ldg4(mask + (long long)(mask_div == 1 ? row : row / mask_div) * ldm + col) : f4(1.0f);
        return r;
    }
    __device__ __forceinline__ float apply4(int row, int col, const float4& acc, const Col4& c, const Row4& r) const {
        float4 v;
        v.x = alpha * acc.x + c.bias.x + r.add.x; v.y = alpha * acc.y + c.bias.y + r.add.y;
        v.z = alpha * acc.z + c.bias.z + r.add.z; v.w = alpha * acc.w + c.bias.w + r.add.w;
        if constexpr (ACT) {
            if (relu == ACT_RELU) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
            else if (relu) {
                float4 dv;
                v.x = act_fwd(relu, v.x, dv.x); v.y = act_fwd(relu, v.y, dv.y); v.z = act_fwd(relu, v.z, dv.z); v.w = act_fwd(relu, v.w, dv.w);
                if (dact) st4(dact + (long long)row * lddact + col, dv);
            }
        } else if (relu) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
        if (ACT && mask_mul) { v.x *= r.mask.x; v.y *= r.mask.y; v.z *= r.mask.z; v.w *= r.mask.w; }
        else {
            v.x = r.mask.x > 0.0f ? v.x : 0.0f; v.y = r.mask.y > 0.0f ? v.y : 0.0f;
            v.z = r.mask.z > 0.0f ? v.z : 0.0f; v.w = r.mask.w > 0.0f ? v.w : 0.0f;
        }
        st4(C + (long long)row * ldc + col, v);
        vmax = amax4(vmax, v);
        if constexpr (CS) { cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w; }
        return 0.0f;
    }
};
using EpiStd = EpiStdT<false>;
using EpiStdA = EpiStdT<true>;


// ---- EpiStd that ALSO writes C as the pre-split A operand of the next scaled-fp16 GEMM ----------------------------------------
// C is stored twice: as fp32 (masks, weight gradients and element-wise kernels read that) and, in `Cs`, in the split16 layout
// of gemm_tcgen05_h16.cuh (every 16 consecutive floats of a row -> 16 hi | 16 lo fp16 parts of the SCALED values, same leading
// dimension), so that the consuming GEMM's TMA delivers tensor-core format and its splitter warps have nothing to do - the
// shared-memory pipe, which bounds that kernel, then carries 88 KB instead of 120 KB per k-block.
// The scale must be known BEFORE the first element is written, i.e. before max |C| exists.  It comes from a bound instead:
//     |C[m][n]| <= max |A| * max_n sum_k |B[n][k]| + max |bias| + bound(add)          (Hoelder; relu and masks only shrink)
// with the EXACT maximum of A (tracked by A's producer), so the slack does not compound from layer to layer; the per-layer
// operator norms are computed once per parameter update (weight_norms_kernel).  For a dense layer with K = 1024 inputs the
// bound is ~2^7 above the true maximum, which leaves 11 octaves of full 22-bit precision below it (18 with an exact maximum)
// and the same absolute floor; overflow is impossible by construction.  The bound is published in `bound_out` for the
// consumer (GemmShape::a_scale_src) and the exact maximum of C still goes to `amax_out` for ITS bound.
__device__ __forceinline__ uint32_t pack_h2_rn(float first, float second) {
    uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first)); return r;
}
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hp, uint32_t& lp) {
    hp = pack_h2_rn(x, y);
    unsigned short h0, h1; asm("mov.b32 {%0, %1}, %2;" : "=h"(h0), "=h"(h1) : "r"(hp));
    const unsigned short m1 = 0xBC00; float d0, d1;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d0) : "h"(h0), "h"(m1), "f"(x));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d1) : "h"(h1), "h"(m1), "f"(y));
    lp = pack_h2_rn(d0, d1);
}
__host__ __device__ __forceinline__ uint32_t h16_scale_exp_(float amax) {     // == tc2h::h16_scale_exp (defined there for the kernel)
#ifdef __CUDA_ARCH__
    const uint32_t b = __float_as_uint(amax) & 0x7FFFFFFFu;
#else
    union { float f; uint32_t u; } c; c.f = amax; const uint32_t b = c.u & 0x7FFFFFFFu;
#endif
    int e = (int)(b >> 23);
    if (b == 0u || e == 255) return 127u;
    if (e == 0) e = 1;
    int se = 127 + 14 - (e - 127);
    se = se < 2 ? 2 : (se > 252 ? 252 : se);
    return (uint32_t)se;
}
template <bool ACT, bool CS = false>
struct EpiStdST {
    static constexpr bool kRowSum = false;
    static constexpr bool kColSum = CS;
    float* C; long long ldc;
    const float* bias; const float* mask; long long ldm; const float* add; long long ldadd; int relu; int mask_div;
    float* Cs;                         // split16 copy of C (leading dimension ldc)
    const float* in_amax; const float* in_amax2; float in_bound;   // exact maximum of the A operand (as GemmShape)
    const float* w_norm;               // device: max_n sum_k |B[n][k]|
    const float* bias_amax;            // device or null
    const float* add_bound;            // device or null: bound on |add|
    float* bound_out;                  // device: receives the bound on |C| the copy was scaled with
    // a second product whose result shares C's scale (the two halves of cat = [s_x | s_t] are ONE operand of Dense_5): the
    // bound is the larger of the two layers' bounds, so both layers compute the same scale
    const float* alt_amax = nullptr; const float* alt_w_norm = nullptr; const float* alt_bias = nullptr;
    float* amax_out = nullptr;         // device or null: exact max |C|
    mutable float vmax = 0.0f;
    float* dact = nullptr; long long lddact = 0;   // as EpiStd: activation derivative output / multiplicative mask
    int mask_mul = 0;
    float* colsum_part = nullptr; long long ldcs = 0;      // as EpiStdT
    mutable float4 cs = {0.0f, 0.0f, 0.0f, 0.0f};
    __device__ __forceinline__ void at_z(int) {}
    __device__ __forceinline__ float out_bound() const {
        float a = in_bound;
        if (!(a > 0.0f)) { a = *in_amax; if (in_amax2) a = fmaxf(a, *in_amax2); }
        float b = a * *w_norm;
        if (bias_amax) b += *bias_amax;
        if (add_bound) b += *add_bound;
        if (alt_amax) b = fmaxf(b, *alt_amax * *alt_w_norm + (alt_bias ? *alt_bias : 0.0f));
        return b;
    }
    __device__ __forceinline__ float out_scale() const { return __uint_as_float(h16_scale_exp_(out_bound()) << 23); }
    struct Aux { float bias, add, mask, scale; };
    __device__ __forceinline__ Aux load(int row, int col) const {
        Aux a;
        a.bias = bias ? __ldg(bias + col) : 0.0f;
        a.add = add ? add[(long long)row * ldadd + col] : 0.0f;
        a.mask = mask ? __ldg(mask + (long long)(mask_div == 1 ? row : row / mask_div) * ldm + col) : 1.0f;
        a.scale = out_scale();
        return a;
    }
    __device__ __forceinline__ float apply(int row, int col, float acc, const Aux& a) const {
        float v = acc + a.bias + a.add;
        if constexpr (ACT) {
            if (relu == ACT_RELU) v = fmaxf(v, 0.0f);
            else if (relu) { float dv; v = act_fwd(relu, v, dv); if (dact) dact[(long long)row * lddact + col] = dv; }
            v = mask_mul ? v * a.mask : (a.mask > 0.0f ? v : 0.0f);
        } else {
            if (relu) v = fmaxf(v, 0.0f);
            v = a.mask > 0.0f ? v : 0.0f;
        }
        C[(long long)row * ldc + col] = v;
        vmax = fmaxf(vmax, fabsf(v));
        const float x = v * a.scale;
        const __half h = __float2half_rn(x), l = __float2half_rn(x - __half2float(h));
        __half* g = reinterpret_cast<__half*>(Cs + (long long)row * ldc + (col & ~15));
        g[col & 15] = h; g[16 + (col & 15)] = l;
        return 0.0f;
    }
    __device__ __forceinline__ float operator()(int row, int col, float acc) const { return apply(row, col, acc, load(row, col)); }
    __device__ __forceinline__ void row_partial(int, int, float) const {}
    struct Col4 { float4 bias; float scale; };
    struct Row4 { float4 add, mask; };
    bool vec_ok() const {
        return aligned16(C) && ldc % 16 == 0 && (reinterpret_cast<uintptr_t>(Cs) & 63) == 0 && (!bias || aligned16(bias)) &&
               (!mask || (aligned16(mask) && ldm % 4 == 0)) && (!add || (aligned16(add) && ldadd % 4 == 0));
    }
    __device__ __forceinline__ Col4 load_col4(int col) const { Col4 c; c.bias = bias ? ldg4(bias + col) : f4(0.0f); c.scale = out_scale(); return c; }
    __device__ __forceinline__ Row4 load_row4(int row, int col) const {
        Row4 r;
        r.add = add ? ld4(add + (long long)row * ldadd + col) : f4(0.0f);        // may alias C: plain load
        r.mask = mask ? ldg4(mask + (long long)(mask_div == 1 ? row : row / mask_div) * ldm + col) : f4(1.0f);
        return r;
    }
    __device__ __forceinline__ float apply4(int row, int col, const float4& acc, const Col4& c, const Row4& r) const {
        float4 v;
        v.x = acc.x + c.bias.x + r.add.x; v.y = acc.y + c.bias.y + r.add.y;
        v.z = acc.z + c.bias.z + r.add.z; v.w = acc.w + c.bias.w + r.add.w;
        if constexpr (ACT) {
            if (relu == ACT_RELU) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
            else if (relu) {
                float4 dv;
                v.x = act_fwd(relu, v.x, dv.x); v.y = act_fwd(relu, v.y, dv.y); v.z = act_fwd(relu, v.z, dv.z); v.w = act_fwd(relu, v.w, dv.w);
                if (dact) st4(dact + (long long)row * lddact + col, dv);
            }
        } else if (relu) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
        if (ACT && mask_mul) { v.x *= r.mask.x; v.y *= r.mask.y; v.z *= r.mask.z; v.w *= r.mask.w; }
        else {
            v.x = r.mask.x > 0.0f ? v.x : 0.0f; v.y = r.mask.y > 0.0f ? v.y : 0.0f;
            v.z = r.mask.z > 0.0f ? v.z : 0.0f; v.w = r.mask.w > 0.0f ? v.w : 0.0f;
        }
        st4(C + (long long)row * ldc + col, v);
        vmax = amax4(vmax, v);
        if constexpr (CS) { cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w; }
        // columns col..col+3 of the 16-group: hi parts at byte 2 (col % 16), lo parts 32 bytes further
        uint2 hp, lp;
        split_pair(v.x * c.scale, v.y * c.scale, hp.x, lp.x);
        split_pair(v.z * c.scale, v.w * c.scale, hp.y, lp.y);
        char* g = reinterpret_cast<char*>(Cs + (long long)row * ldc + (col & ~15)) + 2 * (col & 15);
        *reinterpret_cast<uint2*>(g) = hp;
        *reinterpret_cast<uint2*>(g + 32) = lp;
        return 0.0f;
    }
};
using EpiStdS = EpiStdST<false>;
using EpiStdSA = EpiStdST<true>;

}  // namespace mfm

// ---- dispatcher: CTA-pair tcgen05 kernel for large aligned problems, single-CTA tcgen05 kernel for
// medium ones, warp-level kernel otherwise ----
#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_2sm.cuh"
#include "gemm_tcgen05_persist.cuh"
#include "gemm_tcgen05_h16.cuh"
namespace mfm {
// 0 = auto, 1 = force mma.sync (env MFM_GEMM=mma), 2 = tcgen05 single-CTA only (env MFM_GEMM=tc1),
// 3 = no persistent kernel: one-tile CTA-pair kernel with separate cross-term accumulators (env MFM_GEMM=tc2)
int gemm_backend();
template <bool A_KMAJOR, bool B_NMAJOR>
inline int gemm_path(const GemmShape& p) {       // 2 = CTA pair, 1 = single CTA, 0 = mma.sync
    const int be = gemm_backend();
    if (be == 1) return 0;
    if ((be == 0 || be == 3) && tc2::eligible<A_KMAJOR, B_NMAJOR>(p)) return 2;
    return tc::eligible<A_KMAJOR, B_NMAJOR>(p) ? 1 : 0;
}
template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
inline cudaError_t launch_gemm(const GemmShape& p, const Epi& epi, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0) return cudaSuccess;
    switch (gemm_path<A_KMAJOR, B_NMAJOR>(p)) {
        case 2:
            // persistent (epilogue overlapped with the next tile's MMAs) unless the reduction is split:
            // split-K tiles have long main loops (nothing to hide) and need the separate accumulators
            if constexpr (A_KMAJOR && !B_NMAJOR) {
                // scaled fp16 split, three tensor-core passes (gemm_tcgen05_h16.cuh): every K-major x K-major layer with 16-aligned K
                if (tc2h::gemm_h16() && gemm_backend() == 0 && tc2h::eligible(p, epi)) return tc2h::launch<Epi>(p, epi, st);
            }
            if (gemm_backend() == 0 && p.k_split == 0 && tc2p::eligible<A_KMAJOR, B_NMAJOR>(p, epi))
                return tc2p::launch<A_KMAJOR, B_NMAJOR, Epi>(p, epi, st);
            return tc2::launch<A_KMAJOR, B_NMAJOR, Epi>(p, epi, st);
        case 1: return tc::launch<A_KMAJOR, B_NMAJOR, Epi>(p, epi, st);
        default: return launch_gemm_mma<A_KMAJOR, B_NMAJOR, Epi>(p, epi, st);
    }
}
}  // namespace mfm
