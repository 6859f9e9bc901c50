// 3xTF32 GEMM on 5th-gen tensor cores: TMA -> shared memory -> tcgen05.mma (kind::tf32) with the
// accumulator in TMEM, fp32 in / fp32 out.  Same arithmetic and epilogue functors as gemm_tf32x3.cuh.
//
// CTA = 128 x 256 output tile, BK = 32 (one 128-byte swizzle atom of fp32), 2-stage ring:
//   warp 0      TMA producer: raw fp32 A/B tiles (SWIZZLE_128B) -> stage buffers
//   warps 4-11  splitters: a -> (hi = tf32(a), lo = tf32(a - hi)) element-wise, in the swizzled layout
//               (hi overwrites the raw tile, lo goes to a twin buffer), then fence.proxy.async
//   warp 1      MMA issuer (one elected lane): per k-step of 8:  D += lo*hi, hi*lo, hi*hi
//               tcgen05.commit releases the stage back to the producer
//   warp 2      TMEM allocator (512 fp32 columns: main and cross-term accumulators)
//   warps 4-11  epilogue: tcgen05.ld 32x32b -> transpose through smem -> coalesced functor calls
//
// Operand layouts (all via TMA tensor maps built on the host):
//   K-major  X[rows][K]  : 2-D map, box {32 k, rows}, canonical K-major SW128 (SBO = 1024 B)
//   MN-major X[K][cols]  : 3-D map {32 cols, K, cols/32}, box {32, 32 k, blocks}: blocks of 32 columns,
//                          canonical MN-major SW128_32B (LBO = 4096 B between column blocks, SBO = 512 B)
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_tf32x3.cuh"

namespace mfm {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 2;
constexpr int THREADS = 384;                // 12 warps
constexpr int SPLIT_WARP0 = 4, SPLIT_WARPS = 8;
constexpr int A_BYTES = BM * BK * 4;        // 16 KB
constexpr int B_BYTES = BN * BK * 4;        // 32 KB
constexpr int HI_BYTES = A_BYTES + B_BYTES; // 48 KB (hi tiles, written by TMA)
constexpr int STAGE_BYTES = 2 * HI_BYTES;   // + lo twins
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;              // columns [0,256): hi*hi sums, [256,512): lo*hi + hi*lo sums

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (sm_100 "version 1").
// K-major operands: SWIZZLE_128B (layout type 2; 8-row x 128-byte atoms, SBO = 1024 B).
// MN-major tf32 operands must use SWIZZLE_128B_BASE32B (layout type 1; Swizzle<2,5,2>: 32-byte
// granules, 4-row x 128-byte atoms) -- the only MN-major layout the tensor core accepts for 32-bit
// types; TMA produces it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO = stride between 32-column
// blocks (4096 B), SBO = stride between 4-row groups (512 B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
    return ((uint64_t)hi << 32) | lo;
}

// instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Maps { CUtensorMap a, b; };

template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ Maps maps, GemmShape p, Epi epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                  // TMA landed
    uint64_t* split = bars + STAGES;        // hi/lo ready
    uint64_t* empty = bars + 2 * STAGES;    // MMAs done reading
    uint64_t* acc_full = bars + 3 * STAGES;
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = p.n_rows_dev ? min(*p.n_rows_dev, p.M) : p.M;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= M) return;                    // uniform per CTA: safe before any barrier use
    const int kz0 = p.k_split > 0 ? blockIdx.z * p.k_split : 0;
    const int Kend = p.k_split > 0 ? min(p.K, kz0 + p.k_split) : p.K;
    const int KT = (Kend - kz0 + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], SPLIT_WARPS * 32); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], HI_BYTES);
                const int k0 = kz0 + kt * BK;
                if (A_KMAJOR) tma_load_2d(st, &maps.a, &full[s], k0, m0);
                else          tma_load_3d(st, &maps.a, &full[s], 0, k0, m0 / 32);
                if (!B_NMAJOR) tma_load_2d(st + A_BYTES, &maps.b, &full[s], k0, n0);
                else           tma_load_3d(st + A_BYTES, &maps.b, &full[s], 0, k0, n0 / 32);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(!A_KMAJOR, B_NMAJOR);
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                mbar_wait(&split[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), b_hi = a_hi + A_BYTES;
                const uint32_t a_lo = a_hi + HI_BYTES, b_lo = b_hi + HI_BYTES;
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint32_t ao = A_KMAJOR ? ks * 32 : ks * 1024;
                    const uint32_t bo = !B_NMAJOR ? ks * 32 : ks * 1024;
                    const uint32_t albo = A_KMAJOR ? 16 : 4096, blbo = !B_NMAJOR ? 16 : 4096;
                    const uint32_t asbo = A_KMAJOR ? 1024 : 512, bsbo = !B_NMAJOR ? 1024 : 512;
                    const uint32_t alt = A_KMAJOR ? 2 : 1, blt = !B_NMAJOR ? 2 : 1;
                    const uint64_t dah = make_desc(a_hi + ao, albo, asbo, alt), dal = make_desc(a_lo + ao, albo, asbo, alt);
                    const uint64_t dbh = make_desc(b_hi + bo, blbo, bsbo, blt), dbl = make_desc(b_lo + bo, blbo, bsbo, blt);
                    // The TMEM accumulator adds with truncation (measured: bias ~ 2.3e-9 relative per
                    // accumulation into a large partial sum).  The two small cross terms therefore get
                    // their own accumulator, so the main chain sees one truncation per k-step, not three.
                    mma_tf32_ss(tmem_base + BN, dal, dbh, idesc, (kt | ks) != 0);
                    mma_tf32_ss(tmem_base + BN, dah, dbl, idesc, 1);
                    mma_tf32_ss(tmem_base, dah, dbh, idesc, (kt | ks) != 0);
                }
                mma_commit(&empty[s]);
            }
            mma_commit(acc_full);
        }
    } else if (warp >= SPLIT_WARP0) {
        // ---------------- splitters ----------------
        const int t = threadIdx.x - SPLIT_WARP0 * 32;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
            mbar_wait(&full[s], ph);
            float4* hi = (float4*)(smem + s * STAGE_BYTES);
            float4* lo = (float4*)(smem + s * STAGE_BYTES + HI_BYTES);
            constexpr int PER = HI_BYTES / 16 / (SPLIT_WARPS * 32);    // 12 float4 per thread
            float4 v[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) v[i] = hi[t + i * SPLIT_WARPS * 32];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                float4 h, l;
                uint32_t hh, ll;
                split_tf32(v[i].x, hh, ll); h.x = __uint_as_float(hh); l.x = __uint_as_float(ll);
                split_tf32(v[i].y, hh, ll); h.y = __uint_as_float(hh); l.y = __uint_as_float(ll);
                split_tf32(v[i].z, hh, ll); h.z = __uint_as_float(hh); l.z = __uint_as_float(ll);
                split_tf32(v[i].w, hh, ll); h.w = __uint_as_float(hh); l.w = __uint_as_float(ll);
                hi[t + i * SPLIT_WARPS * 32] = h;
                lo[t + i * SPLIT_WARPS * 32] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core
            mbar_arrive(&split[s]);
        }
        // ---------------- epilogue ----------------
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (p.k_split > 0) epi.at_z(blockIdx.z);
        const int ew = warp - SPLIT_WARP0;            // 0..7
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        const int chalf = ew >> 2;                    // column half (128 columns)
        float* stg = (float*)smem + ew * (32 * 33);   // pipeline buffers are idle now
        const int row_base = m0 + quad * 32;
        float rowacc[2] = {0.0f, 0.0f};
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
            const int col0 = chalf * 128 + cc * 32;
            if (n0 + col0 >= p.N) break;
            uint32_t r[32], r2[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, r);
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(BN + col0), r2);
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
            __syncwarp();
            const int col = n0 + col0 + lane;
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
                const int row = row_base + rr;
                float c = 0.0f;
                if (row < M && col < p.N) c = epi(row, col, stg[rr * 33 + lane]);
                if (Epi::kRowSum) {
                    c = warp_sum(c);
                    if (lane == rr) rowacc[cc >> 1] += c;
                }
            }
            __syncwarp();
        }
        if (Epi::kRowSum) {
            const int row = row_base + lane;
            if (row < M) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int tile = (n0 + chalf * 128 + g * 64) / GBN;     // 64-column groups, as the mma.sync path
                    if (n0 + chalf * 128 + g * 64 < p.N) epi.row_partial(row, tile, rowacc[g]);
                }
            }
        }
        epi_publish_amax(epi, epi_stored_max(epi, 0));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// K-major operand X[rows][K] (ld floats)
inline bool make_map_kmajor(CUtensorMap* m, const float* base, long long ld, int rows, int K, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// MN-major operand X[K][cols] (ld floats), cols % 32 == 0
inline bool make_map_mnmajor(CUtensorMap* m, const float* base, long long ld, int cols, int K, int box_blocks) {
    cuuint64_t gdim[3] = {32, (cuuint64_t)K, (cuuint64_t)(cols / 32)};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, 128};
    cuuint32_t box[3] = {32, 32, (cuuint32_t)box_blocks};
    cuuint32_t es[3] = {1, 1, 1};
    return encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool A_KMAJOR, bool B_NMAJOR>
inline bool eligible(const GemmShape& p) {
    if (encode_fn() == nullptr) return false;
    if (p.M < 128 || p.N < 64 || p.K < 32) return false;                          // tiny problems: warp-level path
    // ... and small ones: a single 128-row tile of a 128 x 128 x 256 layer took 51 us on this kernel (one CTA: TMEM allocation,
    // barrier set-up and a 2-stage pipeline for 8 k-blocks) against 13 us on the warp-level kernel (profiles/r02 4-mode launch list)
    if ((long long)p.M * p.N * p.K < (1ll << 25)) return false;
    if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.B) & 15)) return false;
    if (p.lda % 4 || p.ldb % 4) return false;
    if (!A_KMAJOR && (p.M % 32)) return false;
    if (B_NMAJOR && (p.N % 32)) return false;
    if (p.k_split > 0 && (p.k_split % BK)) return false;
    return true;
}

template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
inline cudaError_t launch(const GemmShape& p, const Epi& epi, cudaStream_t st) {
    Maps maps;
    bool ok = A_KMAJOR ? make_map_kmajor(&maps.a, p.A, p.lda, p.M, p.K, BM) : make_map_mnmajor(&maps.a, p.A, p.lda, p.M, p.K, BM / 32);
    ok = ok && (!B_NMAJOR ? make_map_kmajor(&maps.b, p.B, p.ldb, p.N, p.K, BN) : make_map_mnmajor(&maps.b, p.B, p.ldb, p.N, p.K, BN / 32));
    if (!ok) return cudaErrorInvalidValue;
    auto kern = gemm_tc_kernel<A_KMAJOR, B_NMAJOR, Epi>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.k_split > 0 ? (p.K + p.k_split - 1) / p.k_split : 1);
    kern<<<grid, THREADS, SMEM_BYTES, st>>>(maps, p, epi);
    ++g_mfm_launches;
    return cudaGetLastError();
}

}  // namespace tc
}  // namespace mfm
