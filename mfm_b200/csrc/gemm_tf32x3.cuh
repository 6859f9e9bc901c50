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

constexpr int GBM = 128, GBN = 128, GBK = 16, GSTAGES = 3, GTHREADS = 256;
// smem strides chosen so fragment loads are bank-conflict free (see DESIGN.md "GEMM").
constexpr int LDS_K = GBK + 4;    // operand stored [rows][K]   (K contiguous)
constexpr int LDS_MN = GBM + 8;   // operand stored [K][rows]   (M/N contiguous)
constexpr int A_STAGE = (GBM * LDS_K > GBK * LDS_MN) ? GBM * LDS_K : GBK * LDS_MN;
constexpr int B_STAGE = A_STAGE;
constexpr int GEMM_SMEM_BYTES = GSTAGES * (A_STAGE + B_STAGE) * 4 + GBM * 4 * 4;

struct GemmShape {
    int M, N, K;
    const float* A; long long lda;   // A_KMAJOR: A[m*lda + k]; else A[k*lda + m]
    const float* B; long long ldb;   // B_NMAJOR: B[k*ldb + n]; else B[n*ldb + k]
    const int* n_rows_dev;           // optional: device int, effective M (active rows); tiles beyond exit
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
// KMAJ=false: global is [k][row] (row contiguous) -> smem [k][LDS_MN].
template <bool KMAJ>
__device__ __forceinline__ void load_tile(float* s, const float* __restrict__ g, long long ld, int row0, int k0,
                                          int rows, int K, bool vec_ok) {
    const int tid = threadIdx.x;
    if (KMAJ) {
        // 128 rows x 16 k = 512 float4 chunks; 2 per thread
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int c = tid + it * GTHREADS;
            const int r = c >> 2, kk = (c & 3) * 4;
            float* dst = s + r * LDS_K + kk;
            const int gr = row0 + r, gk = k0 + kk;
            const bool full = (gr < rows) && (gk + 3 < K);
            if (vec_ok) {
                const float* src = g + (long long)(full ? gr : 0) * ld + (full ? gk : 0);
                if (full || gr >= rows || gk >= K) { cp_async16(dst, src, full); continue; }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
                dst[e] = (gr < rows && gk + e < K) ? g[(long long)gr * ld + gk + e] : 0.0f;
        }
    } else {
        // 16 k x 128 rows = 512 float4 chunks
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int c = tid + it * GTHREADS;
            const int kk = c >> 5, r = (c & 31) * 4;
            float* dst = s + kk * LDS_MN + r;
            const int gk = k0 + kk, gr = row0 + r;
            const bool full = (gk < K) && (gr + 3 < rows);
            if (vec_ok) {
                const float* src = g + (long long)(full ? gk : 0) * ld + (full ? gr : 0);
                if (full || gk >= K || gr >= rows) { cp_async16(dst, src, full); continue; }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
                dst[e] = (gk < K && gr + e < rows) ? g[(long long)gk * ld + gr + e] : 0.0f;
        }
    }
}

// Epilogue functor contract:
//   __device__ float operator()(int row, int col, float acc) const;   // stores; returns row-sum contribution
//   static constexpr bool kRowSum;                                     // if true: row_partial(row, ntile, sum)
//   __device__ void row_partial(int row, int ntile, float s) const;
template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
__global__ void __launch_bounds__(GTHREADS, 2) gemm_tf32x3_kernel(GemmShape p, Epi epi) {
    extern __shared__ __align__(16) float gsm[];
    float* sA = gsm;
    float* sB = gsm + GSTAGES * A_STAGE;
    float* sRed = gsm + GSTAGES * (A_STAGE + B_STAGE);   // [GBM][4]

    const int M = p.n_rows_dev ? min(*p.n_rows_dev, p.M) : p.M;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    if (m0 >= M) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
    const int g = lane >> 2, t = lane & 3;

    const bool a_vec = ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) && (p.lda % 4 == 0);
    const bool b_vec = ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0) && (p.ldb % 4 == 0);

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0f;

    const int KT = (p.K + GBK - 1) / GBK;
    auto issue = [&](int kt) {
        if (kt < KT) {
            const int s = kt % GSTAGES;
            load_tile<A_KMAJOR>(sA + s * A_STAGE, p.A, p.lda, m0, kt * GBK, M, p.K, a_vec);
            load_tile<!B_NMAJOR>(sB + s * B_STAGE, p.B, p.ldb, n0, kt * GBK, p.N, p.K, b_vec);
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
            uint32_t ah[4][4], al[4][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = wm + i * 16 + g;
                float v0, v1, v2, v3;
                if (A_KMAJOR) {
                    v0 = a[r * LDS_K + ks + t];       v1 = a[(r + 8) * LDS_K + ks + t];
                    v2 = a[r * LDS_K + ks + t + 4];   v3 = a[(r + 8) * LDS_K + ks + t + 4];
                } else {
                    v0 = a[(ks + t) * LDS_MN + r];     v1 = a[(ks + t) * LDS_MN + r + 8];
                    v2 = a[(ks + t + 4) * LDS_MN + r]; v3 = a[(ks + t + 4) * LDS_MN + r + 8];
                }
                split_tf32(v0, ah[i][0], al[i][0]); split_tf32(v1, ah[i][1], al[i][1]);
                split_tf32(v2, ah[i][2], al[i][2]); split_tf32(v3, ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wn + j * 8 + g;
                float v0, v1;
                if (B_NMAJOR) { v0 = b[(ks + t) * LDS_MN + c]; v1 = b[(ks + t + 4) * LDS_MN + c]; }
                else          { v0 = b[c * LDS_K + ks + t];    v1 = b[c * LDS_K + ks + t + 4]; }
                split_tf32(v0, bh[j][0], bl[j][0]); split_tf32(v1, bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma_tf32(acc[i][j], al[i], bh[j]);
                    mma_tf32(acc[i][j], ah[i], bl[j]);
                    mma_tf32(acc[i][j], ah[i], bh[j]);
                }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float rs0 = 0.0f, rs1 = 0.0f;
        const int r0 = m0 + wm + i * 16 + g, r1 = r0 + 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
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
                sRed[(wm + i * 16 + g) * 4 + (warp & 3)] = rs0;
                sRed[(wm + i * 16 + g + 8) * 4 + (warp & 3)] = rs1;
            }
        }
    }
    if (Epi::kRowSum) {
        __syncthreads();
        if (tid < GBM && m0 + tid < M) {
            const float s = ((sRed[tid * 4 + 0] + sRed[tid * 4 + 1]) + sRed[tid * 4 + 2]) + sRed[tid * 4 + 3];
            epi.row_partial(m0 + tid, blockIdx.x, s);
        }
    }
}

template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
inline cudaError_t launch_gemm(const GemmShape& p, const Epi& epi, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0) return cudaSuccess;
    auto kern = gemm_tf32x3_kernel<A_KMAJOR, B_NMAJOR, Epi>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((p.N + GBN - 1) / GBN, (p.M + GBM - 1) / GBM);
    kern<<<grid, GTHREADS, GEMM_SMEM_BYTES, st>>>(p, epi);
    return cudaGetLastError();
}

inline int gemm_n_tiles(int N) { return (N + GBN - 1) / GBN; }

// ---- standard epilogue: C = mask(relu(alpha*acc + bias) ) (+ add) ----------------------------
struct EpiStd {
    static constexpr bool kRowSum = false;
    float* C; long long ldc;
    const float* bias;               // [N] or null
    const float* mask; long long ldm; // out = mask[row,col] > 0 ? out : 0  (relu' gate) or null
    const float* add; long long ldadd; // out += add[row,col] or null
    float alpha; int relu;
    __device__ __forceinline__ float operator()(int row, int col, float acc) const {
        float v = alpha * acc;
        if (bias) v += bias[col];
        if (add) v += add[(long long)row * ldadd + col];
        if (relu) v = fmaxf(v, 0.0f);
        if (mask) v = mask[(long long)row * ldm + col] > 0.0f ? v : 0.0f;
        C[(long long)row * ldc + col] = v;
        return 0.0f;
    }
    __device__ __forceinline__ void row_partial(int, int, float) const {}
};

}  // namespace mfm
