// Weight-gradient GEMM on scaled-fp16 parts:  dW[in,out] = A^T[in,n] G[n,out]  with the batch (chains) as the reduction
// dimension, both operands read from the split16 copies their producers already wrote for the forward / backward-data layers
// (gemm_tcgen05_h16.cuh: every 16 floats of a row -> 16 hi | 16 lo fp16 parts of the values times a per-tensor power of two).
//
// Both operands are MN-major here (the contiguous index of a row is the OUTPUT index of the product).  A 4-D TMA map
//     {16 parts-of-a-group, 2 (hi | lo), cols / 16 groups, n rows},  box {16, 1, 1, 64},  SWIZZLE_32B
// picks the 16 hi (or lo) parts of ONE group for 64 consecutive rows: 64 rows x 32 bytes, dense - eight atoms of the canonical
// MN-major 32-byte-swizzle operand layout (16 MN-elements x 8 k-rows = 256 B; SBO = 256 B between 8-row groups, LBO = 2 KB
// between 16-element blocks).  [Measured with scripts/probes/tma4d_probe.cu: under SWIZZLE_128B a box whose inner row is 32 B is
// NOT packed - every inner row gets its own 128-byte line - so four groups per 128-byte line, the SWIZZLE_128B tile, cannot be
// produced from this layout; 32-byte rows under SWIZZLE_32B are dense.]  So the same bytes serve as K-major operand of the dense
// layers and as MN-major operand here, no transposed copy, no splitter warps: three kind::f16 MMAs (hi.hi' -> main
// accumulator, hi.lo' + lo.hi' -> cross accumulator, added in the epilogue as in the tf32 kernel) per 16 chains, against six
// tf32-rate slots before.  32 TMA boxes of 2 KB per stage and CTA, issued by two lanes (A: warp 0, G: warp 2).
//
// One 256 x 256 tile per CTA pair (cta_group::2), split-K over grid.z exactly like gemm_tcgen05_2sm.cuh (same epilogue, same
// functors; the accumulator is multiplied by 1 / (s_A s_G), exact powers of two, before the functor sees it).
// Stage = 64 chains: A hi + lo 2 x 16 KB, G half hi + lo 2 x 16 KB = 64 KB, 3 stages.
#pragma once
#include <cuda_fp16.h>
#include "gemm_tcgen05_h16.cuh"

namespace mfm {
namespace tc2w {

using tc::smem_u32; using tc::mbar_init; using tc::mbar_expect_tx; using tc::mbar_wait; using tc::tmem_ld32_nowait; using tc::make_desc;
using tc2::cluster_ctarank; using tc2::cluster_sync_all; using tc2::mbar_arrive_remote; using tc2::mma_commit_2sm;
using tc2p::mma_bf16_ss_2sm;

constexpr int BM = 128, BN = 256, BNH = 128;     // rows per CTA (pair tile 256 x 256), G columns staged per CTA
constexpr int BK = 64, STAGES = 3;               // chains per stage
constexpr int THREADS = 384;
constexpr int EPI_WARP0 = 4, EPI_WARPS = 8, FWD_WARP = 3;
constexpr int BLK_BYTES = BK * 32;               // one 16-element block (one split16 group) of one part: 64 rows x 32 B = 2 KB
constexpr int PART_BYTES = 8 * BLK_BYTES;        // 128 MN-elements of one part = 16 KB
constexpr int STAGE_BYTES = 4 * PART_BYTES;      // A hi | A lo | G hi | G lo = 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 512;
constexpr uint32_t EPI_LD = 132;
static_assert(8 * 32 * EPI_LD * 4 <= STAGES * STAGE_BYTES, "epilogue staging must fit in the idle pipeline buffers");

struct MapsW { CUtensorMap a, g; };

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc2w_kernel(const __grid_constant__ MapsW maps, GemmShape p, Epi epi, const float* __restrict__ a_scale_src,
                 const float* __restrict__ g_scale_src, int vec) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                  // own TMA landed
    uint64_t* both = bars + STAGES;         // the tiles of BOTH CTAs landed (leader only; one forwarded arrive per CTA)
    uint64_t* empty = bars + 2 * STAGES;    // MMAs done reading the stage (multicast commit)
    uint64_t* acc_full = bars + 3 * STAGES;
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int M = p.M;
    const int m0p = blockIdx.y * (2 * BM);
    const int m0 = m0p + (int)rank * BM;
    const int n0 = (blockIdx.x >> 1) * BN;
    const int nrem = p.N - n0;
    const int neff = nrem >= BN ? BN : ((nrem + 63) / 64) * 64;     // MMA N of this tile (each CTA's half: whole pairs of groups)
    const int nb0 = n0 + (int)rank * (neff / 2);
    const int kz0 = p.k_split > 0 ? blockIdx.z * p.k_split : 0;
    const int Kend = p.k_split > 0 ? min(p.K, kz0 + p.k_split) : p.K;
    const int KT = (Kend - kz0 + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.g) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 2); mbar_init(&both[s], 2); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 || warp == 2) {
        // ---------------- TMA producers (both CTAs): warp 0 the 8 groups x 2 parts of A, warp 2 those of this CTA's half of G ----------------
        if (lane == 0) {
            const bool isA = warp == 0;
            const int ngrp = isA ? 8 : neff / 32;          // groups of 16 columns this lane loads per part
            const int g0 = (isA ? m0 : nb0) / 16;
            const CUtensorMap* map = isA ? &maps.a : &maps.g;
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES + (isA ? 0 : 2 * PART_BYTES);
                mbar_expect_tx(&full[s], (uint32_t)(2 * ngrp * BLK_BYTES));
                const int k0 = kz0 + kt * BK;
#pragma unroll
                for (int part = 0; part < 2; ++part)
#pragma unroll
                    for (int b = 0; b < 8; ++b)
                        if (b < ngrp) tma_load_4d(st + part * PART_BYTES + b * BLK_BYTES, map, &full[s], 0, part, g0 + b, k0);
            }
        }
    } else if (warp == FWD_WARP) {
        // ---------------- forwards "own tiles landed" to the leader's MMA lane ----------------
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            mbar_wait(&full[s], (kt / STAGES) & 1);
            if (lane == 0) mbar_arrive_remote(&both[s], 0);
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (rank == 0 && lane == 0) {
            // kind::f16: D = f32, A = B = fp16, both MN-major (bits 15 / 16), M = 256 per pair, N = neff
            const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(neff >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                mbar_wait(&both[s], (kt / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + PART_BYTES;
                const uint32_t g_hi = a_hi + 2 * PART_BYTES, g_lo = a_hi + 3 * PART_BYTES;
#pragma unroll
                for (int ks = 0; ks < BK / 16; ++ks) {
                    const uint32_t o = ks * 512;           // 16 chains = two 8-row groups of 256 B
                    // MN-major, 32-byte swizzle (layout type 6): LBO = 2 KB between 16-element blocks, SBO = 256 B between 8-row groups
                    const uint64_t dah = make_desc(a_hi + o, BLK_BYTES, 256, 6), dal = make_desc(a_lo + o, BLK_BYTES, 256, 6);
                    const uint64_t dgh = make_desc(g_hi + o, BLK_BYTES, 256, 6), dgl = make_desc(g_lo + o, BLK_BYTES, 256, 6);
                    mma_bf16_ss_2sm(tmem_base + BN, dal, dgh, idesc, (kt | ks) != 0);
                    mma_bf16_ss_2sm(tmem_base + BN, dah, dgl, idesc, 1);
                    mma_bf16_ss_2sm(tmem_base, dah, dgh, idesc, (kt | ks) != 0);
                }
                mma_commit_2sm(&empty[s]);
            }
            mma_commit_2sm(acc_full);
        }
    } else if (warp >= EPI_WARP0) {
        // ---------------- epilogue (each CTA drains its own 128 rows); as gemm_tcgen05_2sm.cuh ----------------
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (p.k_split > 0) epi.at_z(blockIdx.z);
        const float inv = tc2h::h16_inv_scale(*a_scale_src) * tc2h::h16_inv_scale(*g_scale_src);
        const int ew = warp - EPI_WARP0;
        const int quad = warp & 3;
        const int chalf = ew >> 2;
        const int row_base = m0 + quad * 32;
        if (vec) {
            const uint32_t stg = smem_u32(smem) + (uint32_t)ew * (32u * EPI_LD * 4u);
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int col0 = chalf * 128 + cc * 32;
                if (n0 + col0 >= p.N) break;
                uint32_t r[32], r2[32];
                tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, r);
                tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(BN + col0), r2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const uint32_t dst = stg + ((uint32_t)lane * EPI_LD + (uint32_t)cc * 32u) * 4u;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + j * 16),
                                 "f"((__uint_as_float(r[4 * j]) + __uint_as_float(r2[4 * j])) * inv),
                                 "f"((__uint_as_float(r[4 * j + 1]) + __uint_as_float(r2[4 * j + 1])) * inv),
                                 "f"((__uint_as_float(r[4 * j + 2]) + __uint_as_float(r2[4 * j + 2])) * inv),
                                 "f"((__uint_as_float(r[4 * j + 3]) + __uint_as_float(r2[4 * j + 3])) * inv) : "memory");
            }
            __syncwarp();
            const int col = n0 + chalf * 128 + 4 * lane;
            const bool cvalid = col < p.N;
            typename Epi::Col4 ca;
            if (cvalid) ca = epi.load_col4(col);
            constexpr int RB = 4;
#pragma unroll 1
            for (int r0 = 0; r0 < 32; r0 += RB) {
                typename Epi::Row4 ra[RB];
                float4 acc[RB];
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    const int row = row_base + r0 + i;
                    if (row < M && cvalid) ra[i] = epi.load_row4(row, col);
                }
#pragma unroll
                for (int i = 0; i < RB; ++i)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(acc[i].x), "=f"(acc[i].y), "=f"(acc[i].z), "=f"(acc[i].w)
                                 : "r"(stg + ((uint32_t)(r0 + i) * EPI_LD + 4u * (uint32_t)lane) * 4u));
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    const int row = row_base + r0 + i;
                    if (row < M && cvalid) epi.apply4(row, col, acc[i], ca, ra[i]);
                }
            }
        } else {
            float* stgf = (float*)smem + ew * (32 * 33);
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int col0 = chalf * 128 + cc * 32;
                if (n0 + col0 >= p.N) break;
                uint32_t r[32], r2[32];
                tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, r);
                tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(BN + col0), r2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) stgf[lane * 33 + j] = (__uint_as_float(r[j]) + __uint_as_float(r2[j])) * inv;
                __syncwarp();
                const int col = n0 + col0 + lane;
#pragma unroll 4
                for (int rr = 0; rr < 32; ++rr) {
                    const int row = row_base + rr;
                    if (row < M && col < p.N) epi(row, col, stgf[rr * 33 + lane]);
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
// split16 copy X_s of X[rows][cols] (ld floats per row) as an MN-major fp16 operand: see the header comment
inline bool make_map_split16_mn(CUtensorMap* m, const float* base_split, long long ld, int cols, int rows) {
    cuuint64_t gdim[4] = {16, 2, (cuuint64_t)(cols / 16), (cuuint64_t)rows};
    cuuint64_t gstr[3] = {32, 64, (cuuint64_t)ld * 4};
    cuuint32_t box[4] = {16, 1, 1, (cuuint32_t)BK};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return tc::encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)base_split, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// p: M = in, N = out, K = rows (chains); p.A / p.B = the split16 copies of the activations [K][in] and of the back-propagated
// signal [K][out]; a_scale_src / g_scale_src: the device floats their scales were derived from
inline bool eligible(const GemmShape& p) {
    if (tc::encode_fn() == nullptr || !tc2h::gemm_h16() || gemm_backend() != 0) return false;
    if (p.M < 256 || p.N < 64 || p.K < 64) return false;
    if (p.M % 16 || p.N % 16 || p.lda % 16 || p.ldb % 16) return false;
    if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.B)) & 63) return false;
    if (p.k_split > 0 && (p.k_split % BK)) return false;
    return true;
}

template <class Epi>
inline cudaError_t launch(const GemmShape& p, const Epi& epi, const float* a_scale_src, const float* g_scale_src, cudaStream_t st) {
    MapsW maps;
    if (!make_map_split16_mn(&maps.a, p.A, p.lda, p.M, p.K) || !make_map_split16_mn(&maps.g, p.B, p.ldb, p.N, p.K)) return cudaErrorInvalidValue;
    auto kern = gemm_tc2w_kernel<Epi>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(2 * ((p.N + BN - 1) / BN), (p.M + 2 * BM - 1) / (2 * BM), p.k_split > 0 ? (p.K + p.k_split - 1) / p.k_split : 1);
    const int vec = (p.N % 4 == 0 && epi.vec_ok()) ? 1 : 0;
    kern<<<grid, THREADS, SMEM_BYTES, st>>>(maps, p, epi, a_scale_src, g_scale_src, vec);
    ++g_mfm_launches;
    return cudaGetLastError();
}

}  // namespace tc2w
}  // namespace mfm
