// 3xTF32 GEMM on a CTA pair (tcgen05.mma.cta_group::2): the same arithmetic, operand layouts and
// epilogue functors as gemm_tcgen05.cuh, but two SMs of one TPC share a 256 x 256 output tile.
//
// Why a pair: with three tensor-core passes per k-step the single-CTA kernel is bound by shared-memory
// bandwidth, not by the tensor pipe (per 32-wide k-block a 128x256 tile moves 48 KB in by TMA, 144 KB
// through the hi/lo splitter and 144 KB into the MMAs = 336 KB against 128 B/clk x 1536 clk = 196 KB).
// In cta_group::2 every CTA stages its own 128 rows of A and only HALF of the B tile (128 of 256
// columns); the hardware exchanges the halves.  Per CTA and k-block: 32 KB TMA + 96 KB splitter + 96 KB
// MMA = 224 KB, and the L2->SM traffic per flop drops by a third.
//
// Cluster (2,1,1): rank 0 (leader) issues every MMA.  Per stage:
//   each CTA   TMA producer -> own full[s]   (raw fp32 A rows + B half, SWIZZLE_128B layouts)
//   each CTA   8 splitter warps: hi/lo in place, fence.proxy.async, one release.cluster arrive per warp
//              on the LEADER's split[s] (count 16)
//   leader     MMA lane: waits split[s], 12 x tcgen05.mma.cta_group::2 (lo*hi, hi*lo -> cross
//              accumulator, hi*hi -> main accumulator), tcgen05.commit multicast -> empty[s] of both CTAs
//   each CTA   epilogue warps drain their own 128 TMEM lanes (= their 128 rows of the pair tile)
// Edge column tiles issue the MMA with N trimmed to the next multiple of 64 (1600 = 6 x 256 + 64).
#pragma once
#include "gemm_tcgen05.cuh"

namespace mfm {
namespace tc2 {

using tc::smem_u32; using tc::mbar_init; using tc::mbar_expect_tx; using tc::mbar_wait; using tc::tma_load_2d;
using tc::tma_load_3d; using tc::tmem_ld32_nowait; using tc::make_desc; using tc::Maps;

constexpr int BM = 128;                     // rows per CTA (pair tile: 256 rows)
constexpr int BN = 256;                     // pair-tile columns
constexpr int BNH = BN / 2;                 // B columns staged per CTA
constexpr int BK = 32, STAGES = 3;
constexpr int THREADS = 384;                // 12 warps
constexpr int SPLIT_WARP0 = 4, SPLIT_WARPS = 8;
constexpr int A_BYTES = BM * BK * 4;        // 16 KB
constexpr int B_BYTES = BNH * BK * 4;       // 16 KB
constexpr int HI_BYTES = A_BYTES + B_BYTES; // 32 KB raw/hi tiles (written by TMA)
constexpr int STAGE_BYTES = 2 * HI_BYTES;   // + lo twins
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;              // [0,256): hi*hi sums, [256,512): lo*hi + hi*lo sums
constexpr uint32_t EPI_LD = 132;            // epilogue staging row pitch in floats (conflict-free float4 rows)
static_assert(8 * 32 * EPI_LD * 4 <= STAGES * STAGE_BYTES, "epilogue staging must fit in the idle pipeline buffers");

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.  Default
// (release.cta) semantics as CUTLASS's umma_arrive_2x1SM_sm0: the data this orders was written to the
// arriving CTA's own shared memory and made visible to its tensor core by fence.proxy.async; an
// explicit .release.cluster would add a GPU-scope MEMBAR per stage.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mma_tf32_ss_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
// completion of all prior MMAs of this thread -> mbarrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// optional timeline of one CTA pair (SM clock at 9 events), for tuning: flags bit 1
#ifdef MFM_TC2_TIMELINE
#define TC2_MARK(i) do { if (tl && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2 && blockIdx.z == 0 && lane == 0) tl[i] = clock64(); } while (0)
#else
#define TC2_MARK(i) do { } while (0)
#endif

// instruction descriptor: D=f32, A=B=tf32, majors, N>>3 (runtime), M = 256
__host__ __device__ constexpr uint32_t make_idesc_base(bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(256 >> 4) << 24);
}

template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ Maps maps, GemmShape p, Epi epi, int raw_hi, int vec, long long* tl) {
    extern __shared__ uint8_t smem_raw[];
    // identical carve-up in both CTAs (the dynamic window starts at the same offset in each)
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                  // own TMA landed               (local)
    uint64_t* split = bars + STAGES;        // hi/lo of BOTH CTAs ready     (used in the leader only)
    uint64_t* empty = bars + 2 * STAGES;    // MMAs done reading the stage  (multicast commit)
    uint64_t* acc_full = bars + 3 * STAGES; // accumulator complete         (multicast commit)
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();           // == blockIdx.x & 1
    const int M = p.n_rows_dev ? min(*p.n_rows_dev, p.M) : p.M;
    const int m0p = blockIdx.y * (2 * BM);             // pair tile origin
    if (m0p >= M) return;                              // uniform over the pair: safe before any barrier use
    const int m0 = m0p + (int)rank * BM;
    const int n0 = (blockIdx.x >> 1) * BN;
    const int nrem = p.N - n0;
    const int neff = nrem >= BN ? BN : ((nrem + 63) / 64) * 64;     // MMA N of this tile (multiple of 64)
    const int nb0 = n0 + (int)rank * (neff / 2);                   // first B column staged by this CTA
    const int kz0 = p.k_split > 0 ? blockIdx.z * p.k_split : 0;
    const int Kend = p.k_split > 0 ? min(p.K, kz0 + p.k_split) : p.K;
    const int KT = (Kend - kz0 + BK - 1) / BK;
    if (warp == 0) TC2_MARK(0);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 2 * SPLIT_WARPS); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                     // barriers of both CTAs initialised before any remote arrive / multicast
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) TC2_MARK(1);

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs) ----------------
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
                if (!B_NMAJOR) tma_load_2d(st + A_BYTES, &maps.b, &full[s], k0, nb0);
                else           tma_load_3d(st + A_BYTES, &maps.b, &full[s], 0, k0, nb0 / 32);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (rank == 0 && lane == 0) {
            const uint32_t idesc = make_idesc_base(!A_KMAJOR, B_NMAJOR) | ((uint32_t)(neff >> 3) << 17);
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const uint32_t ph = (kt / STAGES) & 1;
                mbar_wait(&split[s], ph);          // arrivals come from both CTAs (mbar_arrive_remote)
                if (kt == 0) TC2_MARK(4);
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
                    // cross terms in their own accumulator: the TMEM accumulator adds with truncation,
                    // so the main chain sees one truncation per k-step, not three (see gemm_tcgen05.cuh)
                    mma_tf32_ss_2sm(tmem_base + BN, dal, dbh, idesc, (kt | ks) != 0);
                    mma_tf32_ss_2sm(tmem_base + BN, dah, dbl, idesc, 1);
                    mma_tf32_ss_2sm(tmem_base, dah, dbh, idesc, (kt | ks) != 0);
                }
                mma_commit_2sm(&empty[s]);
            }
            mma_commit_2sm(acc_full);
            TC2_MARK(5);
        }
    } else if (warp >= SPLIT_WARP0) {
        // ---------------- splitters (both CTAs) ----------------
        const int t = threadIdx.x - SPLIT_WARP0 * 32;
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
            mbar_wait(&full[s], ph);
            if (kt == 0 && warp == SPLIT_WARP0) TC2_MARK(2);
            const uint32_t hi = smem_u32(smem + s * STAGE_BYTES) + (uint32_t)t * 16u;   // explicit shared-space accesses
            const uint32_t lo = hi + HI_BYTES;
            constexpr int PER = HI_BYTES / 16 / (SPLIT_WARPS * 32);    // 8 float4 per thread
            constexpr uint32_t STEP = SPLIT_WARPS * 32 * 16;
            float4 v[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(hi + i * STEP));
            if (raw_hi) {
                // the tensor core ignores the 13 low mantissa bits of a tf32 operand: leave the raw
                // fp32 tile in place (hi = trunc(x)) and store only lo = rn_tf32(x - trunc(x))
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    float4 l;
                    l.x = __uint_as_float(f2tf32(v[i].x - __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u)));
                    l.y = __uint_as_float(f2tf32(v[i].y - __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u)));
                    l.z = __uint_as_float(f2tf32(v[i].z - __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u)));
                    l.w = __uint_as_float(f2tf32(v[i].w - __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u)));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo + i * STEP), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                }
            } else {
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    float4 h, l;
                    uint32_t hh, ll;
                    split_tf32(v[i].x, hh, ll); h.x = __uint_as_float(hh); l.x = __uint_as_float(ll);
                    split_tf32(v[i].y, hh, ll); h.y = __uint_as_float(hh); l.y = __uint_as_float(ll);
                    split_tf32(v[i].z, hh, ll); h.z = __uint_as_float(hh); l.z = __uint_as_float(ll);
                    split_tf32(v[i].w, hh, ll); h.w = __uint_as_float(hh); l.w = __uint_as_float(ll);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hi + i * STEP), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo + i * STEP), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&split[s], 0);
            if (kt == 0 && warp == SPLIT_WARP0) TC2_MARK(3);
        }
        // ---------------- epilogue (each CTA drains its own 128 rows) ----------------
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == SPLIT_WARP0) TC2_MARK(6);
        if (p.k_split > 0) epi.at_z(blockIdx.z);
        const int ew = warp - SPLIT_WARP0;            // 0..7
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        const int chalf = ew >> 2;                    // column half (128 columns)
        const int row_base = m0 + quad * 32;
        if (vec) {
            // Warp block = 32 rows x 128 columns.  TMEM -> registers (main + cross) -> shared memory with a
            // 132-float row pitch (conflict-free 16-byte stores); then one ROW per warp instruction:
            // lane l owns columns 4l..4l+3, so every global access is a coalesced 512-byte float4 row
            // segment and the column-only operands are read once per warp.
            const uint32_t stg = smem_u32(smem) + (uint32_t)ew * (32u * EPI_LD * 4u);     // pipeline buffers are idle now
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
                                 "f"(__uint_as_float(r[4 * j]) + __uint_as_float(r2[4 * j])),
                                 "f"(__uint_as_float(r[4 * j + 1]) + __uint_as_float(r2[4 * j + 1])),
                                 "f"(__uint_as_float(r[4 * j + 2]) + __uint_as_float(r2[4 * j + 2])),
                                 "f"(__uint_as_float(r[4 * j + 3]) + __uint_as_float(r2[4 * j + 3])) : "memory");
            }
            __syncwarp();
            const int col = n0 + chalf * 128 + 4 * lane;          // N % 4 == 0: the four columns are valid together
            const bool cvalid = col < p.N;
            typename Epi::Col4 ca;
            if (cvalid) ca = epi.load_col4(col);
            constexpr int RB = 4;                                  // rows whose global reads are issued together
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
                    float c = 0.0f;
                    if (row < M && cvalid) c = epi.apply4(row, col, acc[i], ca, ra[i]);
                    if (Epi::kRowSum) {
                        // 64-column groups (lanes 0-15 / 16-31), as the mma.sync path's n-tiles
                        c += __shfl_xor_sync(0xffffffffu, c, 8); c += __shfl_xor_sync(0xffffffffu, c, 4);
                        c += __shfl_xor_sync(0xffffffffu, c, 2); c += __shfl_xor_sync(0xffffffffu, c, 1);
                        const int gcol = n0 + chalf * 128 + (lane >> 4) * 64;
                        if ((lane & 15) == 0 && row < M && gcol < p.N) epi.row_partial(row, gcol / GBN, c);
                    }
                }
            }
        } else {
            float* stgf = (float*)smem + ew * (32 * 33);
            float rowacc[2] = {0.0f, 0.0f};
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int col0 = chalf * 128 + cc * 32;
                if (n0 + col0 >= p.N) break;
                uint32_t r[32], r2[32];
                tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, r);
                tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(BN + col0), r2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) stgf[lane * 33 + j] = __uint_as_float(r[j]) + __uint_as_float(r2[j]);
                __syncwarp();
                const int col = n0 + col0 + lane;
#pragma unroll 4
                for (int rr = 0; rr < 32; ++rr) {
                    const int row = row_base + rr;
                    float c = 0.0f;
                    if (row < M && col < p.N) c = epi(row, col, stgf[rr * 33 + lane]);
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
        }
        epi_publish_amax(epi, epi_stored_max(epi, 0));
    }
    if (warp == SPLIT_WARP0) TC2_MARK(7);
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                     // the peer's shared memory / TMEM stay alive until both are done
    if (warp == 0) TC2_MARK(8);
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
template <bool A_KMAJOR, bool B_NMAJOR>
inline bool eligible(const GemmShape& p) {
    if (!tc::eligible<A_KMAJOR, B_NMAJOR>(p)) return false;
    return p.M >= 2 * BM;                   // smaller problems: single-CTA kernel
}

long long* gemm_timeline();                 // device buffer for the timeline of one CTA pair, or null (MFM_TC2_TIMELINE builds)
int gemm_raw_hi();                          // 1: rely on the tensor core truncating tf32 operands (env MFM_TC_RAWHI)

template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
inline cudaError_t launch(const GemmShape& p, const Epi& epi, cudaStream_t st) {
    Maps maps;
    bool ok = A_KMAJOR ? tc::make_map_kmajor(&maps.a, p.A, p.lda, p.M, p.K, BM) : tc::make_map_mnmajor(&maps.a, p.A, p.lda, p.M, p.K, BM / 32);
    ok = ok && (!B_NMAJOR ? tc::make_map_kmajor(&maps.b, p.B, p.ldb, p.N, p.K, BNH) : tc::make_map_mnmajor(&maps.b, p.B, p.ldb, p.N, p.K, BNH / 32));
    if (!ok) return cudaErrorInvalidValue;
    auto kern = gemm_tc2_kernel<A_KMAJOR, B_NMAJOR, Epi>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(2 * ((p.N + BN - 1) / BN), (p.M + 2 * BM - 1) / (2 * BM), p.k_split > 0 ? (p.K + p.k_split - 1) / p.k_split : 1);
    const int vec = (p.N % 4 == 0 && epi.vec_ok()) ? 1 : 0;         // float4 epilogue when everything is 16-byte aligned
    kern<<<grid, THREADS, SMEM_BYTES, st>>>(maps, p, epi, gemm_raw_hi(), vec, gemm_timeline());
    ++g_mfm_launches;
    return cudaGetLastError();
}

}  // namespace tc2
}  // namespace mfm
