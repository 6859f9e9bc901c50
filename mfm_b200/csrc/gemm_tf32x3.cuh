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
};

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

// ---- standard epilogue: C = mask(relu(alpha*acc + bias) ) (+ add) ----------------------------
struct EpiStd {
    static constexpr bool kRowSum = false;
    float* C; long long ldc;
    const float* bias;               // [N] or null
    const float* mask; long long ldm; // out = mask[row/mask_div,col] > 0 ? out : 0  (relu' gate) or null
    const float* add; long long ldadd; // out += add[row,col] or null
    float alpha; int relu;
    int mask_div = 1;                  // rows of the mask are shared by mask_div consecutive output rows
    long long c_zstride = 0;           // split-K: slice z writes to C + z*c_zstride
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
        if (relu) v = fmaxf(v, 0.0f);
        v = a.mask > 0.0f ? v : 0.0f;
        C[(long long)row * ldc + col] = v;
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
        if (relu) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
        v.x = r.mask.x > 0.0f ? v.x : 0.0f; v.y = r.mask.y > 0.0f ? v.y : 0.0f;
        v.z = r.mask.z > 0.0f ? v.z : 0.0f; v.w = r.mask.w > 0.0f ? v.w : 0.0f;
        st4(C + (long long)row * ldc + col, v);
        return 0.0f;
    }
};

}  // namespace mfm

// ---- dispatcher: CTA-pair tcgen05 kernel for large aligned problems, single-CTA tcgen05 kernel for
// medium ones, warp-level kernel otherwise ----
#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_2sm.cuh"
#include "gemm_tcgen05_persist.cuh"
#include "gemm_tcgen05_split16.cuh"
namespace mfm {
template <class Epi> struct Split16Epi { static constexpr bool value = false; };
template <> struct Split16Epi<EpiStd> { static constexpr bool value = true; };     // the FM / forward dense layers only
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
            if constexpr (A_KMAJOR && !B_NMAJOR && Split16Epi<Epi>::value) {
                // experimental 3-slot operand split (off by default): needs the weight operand's split16 mirror
                if (tc2s::gemm_split16() && gemm_backend() == 0 && p.k_split == 0 && p.K % 16 == 0 && p.ldb % 16 == 0 &&
                    tc2p::eligible<A_KMAJOR, B_NMAJOR>(p, epi)) {
                    const float* bx = tc2p::lookup_cross(p.B);
                    if (bx && (reinterpret_cast<uintptr_t>(bx) & 63) == 0) return tc2s::launch<Epi>(p, epi, bx, st);
                }
            }
            if (gemm_backend() == 0 && p.k_split == 0 && tc2p::eligible<A_KMAJOR, B_NMAJOR>(p, epi))
                return tc2p::launch<A_KMAJOR, B_NMAJOR, Epi>(p, epi, st);
            return tc2::launch<A_KMAJOR, B_NMAJOR, Epi>(p, epi, st);
        case 1: return tc::launch<A_KMAJOR, B_NMAJOR, Epi>(p, epi, st);
        default: return launch_gemm_mma<A_KMAJOR, B_NMAJOR, Epi>(p, epi, st);
    }
}
}  // namespace mfm
