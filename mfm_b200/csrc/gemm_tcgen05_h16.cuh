// "h16" dense-layer kernel: fp32-accurate products from THREE fp16 tensor-core passes per 16 k-values.
//
// Arithmetic.  Both operands are scaled by a per-tensor power of two and split into two fp16 parts,
//     a * s_a = hi + lo,   hi = rn_fp16(a * s_a),   lo = rn_fp16(a * s_a - hi)        (22 significant bits)
// and  a.b  ~  (hi.hi' + hi.lo' + lo.hi') / (s_a s_b):  three kind::f16 K = 16 MMAs per 16 k-values (the dropped lo.lo'
// term is <= 2^-22 relative).  Against the tf32 + bf16-cross scheme of gemm_tcgen05_persist.cuh (one tf32 K = 8 MMA and one
// bf16 K = 16 MMA per 8 k-values = 4 bf16-rate slots per 16 k) this needs 3 slots, i.e. the ceiling of the arithmetic moves
// from 1/4 to 1/3 of the bf16 peak, and it is MORE accurate (operand rounding 2^-22 instead of bf16's 2^-8 on the cross
// terms: measured max error / max |C| 8e-8 against 1.3e-6, profiles/r01_emulation_error_study.txt).
//
// Why the scale.  fp16 has 5 exponent bits: unscaled, hi is sub-normal below 6e-5 and lo below 0.1, which ruined
// back-propagated signals of size 1e-6 (2e-2 relative error in the study).  With s chosen so that max |a| s lies in
// [2^14, 2^15) every element down to 2^-18 of the tensor's maximum keeps all 22 bits and smaller ones an ABSOLUTE error of
// 2^-25 / s (2^-40 of the maximum) - the normwise accuracy of fp32 for the whole tensor.  Powers of two make the scaling and
// the un-scaling of the accumulator exact.  The maximum is exact, not estimated: every producer of a GEMM operand
// (EpiStd epilogues, the element-wise kernels) folds max |value| into a device slot with one atomicMax per warp, the
// consumer reads the slot (GemmShape::a_amax); tensors without a slot get one reduction pass (absmax_kernel).
//
// Layout ("split16"): every 16 consecutive fp32 of a K-major row are replaced, in the same 64 bytes, by 16 hi parts followed
// by 16 lo parts.  A 32-wide k-block is still a 128-byte SWIZZLE_128B row [hi 0..15 | lo 0..15 | hi 16..31 | lo 16..31], so
// TMA maps, row pitch and the +32-byte descriptor stepping of the persistent kernel apply unchanged.  The weight operand B
// arrives pre-split from global memory when the caller registered a mirror (presplit_h16_kernel, once per parameter
// update; BPRE) and is split in shared memory otherwise; A is loaded raw and split IN PLACE by the splitter warps (one
// 64-byte group per thread per k-block, no cross region).  A stage is therefore 32 KB (A 16 KB + half of B 16 KB) and the
// ring has 6 stages in the same 192 KB.  Shared-memory pipe per k-block and CTA: TMA 32 KB in, splitters 16 KB out + 16 KB in
// (32 + 32 without a mirror), six MMAs x (4 KB of A + 4 KB of B, the pair's other half arrives by the cta_group::2
// broadcast) = 112 KB against 768 clk x 128 B = 96 KB at the full tensor rate: the kernel is co-limited by the
// shared-memory pipe at ~88 % of the 3-slot ceiling, which is what it measures.
// Everything else - work list incl. stream-K, TMEM double buffering, 8 + 8 epilogue warps - is the persistent kernel's.
#pragma once
#include <cuda_fp16.h>
#include "gemm_tcgen05_persist.cuh"

namespace mfm {
namespace tc2h {

using namespace tc2p;       // tile constants, Sched / PairWork, Maps3, barrier + MMA helpers
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { mma_bf16_ss_2sm(d, a, b, idesc, acc); }   // kind::f16: the descriptor says fp16

constexpr int S_STAGES = 6, S_STAGE_BYTES = HI_BYTES;          // 6 x 32 KB
constexpr int S_NBARS = 3 * S_STAGES + 4;
constexpr int S_SMEM_BYTES = S_STAGES * S_STAGE_BYTES + EPI_STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static_assert(S_NBARS * 8 + 8 + 64 <= 256, "barrier area: barriers, TMEM slot, PairWork");
static_assert(S_STAGE_BYTES == EPI_WARPS * 4096, "the helpers' staging is exactly ring stage 0");

// power of two s with amax * s in [2^14, 2^15) (fp16 max is 65504); 1 for a zero or non-finite maximum
__host__ __device__ __forceinline__ uint32_t h16_scale_exp(float amax) {
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
__device__ __forceinline__ float h16_scale(float amax) { return __uint_as_float(h16_scale_exp(amax) << 23); }
__device__ __forceinline__ float h16_inv_scale(float amax) { return __uint_as_float((254u - h16_scale_exp(amax)) << 23); }

// what the kernel needs to know about the operands' magnitudes
struct H16Scales {
    const float* a_amax; const float* a_amax2;   // device: max |A| (the larger of the two when A spans two producers' outputs)
    float a_bound;                               // > 0: a bound known on the host (Fourier features <= 1, normal draws <= 8)
    const float* b_amax;                         // device: max |B| (the mirror was split with h16_scale of this value)
    const float* a_scale_src;                    // APRE: device float the producer of the pre-split A derived its scale from
    int cvt_unpack;                              // A/B switch of the splitter arithmetic (see split16)
    int split_groups;                            // 1, 2 or 4: groups of splitter warps that take the stages in turn
};
__device__ __forceinline__ float h16_a_amax(const H16Scales& s) {
    if (s.a_bound > 0.0f) return s.a_bound;
    float m = *s.a_amax;
    if (s.a_amax2) m = fmaxf(m, *s.a_amax2);
    return m;
}

__device__ __forceinline__ uint32_t pack_h2(float first, float second) {   // `first` lands first in memory
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first));
    return r;
}
__device__ __forceinline__ float h2_lo(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p & 0xFFFFu))); }
__device__ __forceinline__ float h2_hi(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p >> 16))); }
// x - h for the two fp16 halves h of `hp`, in ONE instruction each: sm_100's mixed-precision FMA (fma.rn.f32.f16 -> FHFMA)
// multiplies the fp16 half by -1 and adds the fp32 value, so the hi parts never go back through the conversion unit.
// The difference is exact (|x - h| is at most half an fp16 ulp of x and both share x's exponent range).
__device__ __forceinline__ void sub_h2(uint32_t hp, float x0, float x1, float& d0, float& d1) {
    unsigned short h0, h1;
    asm("mov.b32 {%0, %1}, %2;" : "=h"(h0), "=h"(h1) : "r"(hp));
    const unsigned short m1 = 0xBC00;       // -1.0
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d0) : "h"(h0), "h"(m1), "f"(x0));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d1) : "h"(h1), "h"(m1), "f"(x1));
}
// 16 floats (4 float4, scaled by s) -> 8 words of hi parts, 8 words of lo parts: hi = rn_fp16(x s), lo = rn_fp16(x s - hi).
// Per pair of values: 2 FMUL, 1 F2FP (pack hi), 2 FHFMA, 1 F2FP (pack lo).  Measured against two alternatives at
// 65536 x 1024 x 1024 (scripts/h16_bench.py, sustained): converting hi back with cvt.f32.f16 (CVT_UNPACK, 2 more conversions and
// 2 FADD per pair) 0.377 ms; rounding to 11 bits on the integer pipe instead (4 more integer ops per value) 0.396 ms - the
// splitter warps are bound by issue slots, not by the conversion unit.
template <bool CVT_UNPACK = false>
__device__ __forceinline__ void split16(const float4 (&v)[4], float s, uint32_t (&hp)[8], uint32_t (&lp)[8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float x = v[j].x * s, y = v[j].y * s, z = v[j].z * s, w = v[j].w * s;
        hp[2 * j] = pack_h2(x, y); hp[2 * j + 1] = pack_h2(z, w);
        float dx, dy, dz, dw;
        if (CVT_UNPACK) {      // A/B variant (MFM_GEMM_H16=2)
            dx = x - h2_lo(hp[2 * j]); dy = y - h2_hi(hp[2 * j]); dz = z - h2_lo(hp[2 * j + 1]); dw = w - h2_hi(hp[2 * j + 1]);
        } else {
            sub_h2(hp[2 * j], x, y, dx, dy); sub_h2(hp[2 * j + 1], z, w, dz, dw);
        }
        lp[2 * j] = pack_h2(dx, dy); lp[2 * j + 1] = pack_h2(dz, dw);
    }
}

template <bool APRE, bool BPRE, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc2h_kernel(const __grid_constant__ Maps3 maps, GemmShape p, Epi epi, H16Scales sc, long long* tl, float* sk_ws, unsigned* sk_flags, unsigned sk_epoch) {
#ifdef MFM_TC2_TIMELINE
    // tuning aid: SM clock at 4 events of the first 16 tiles of pair 0's leader (MMA start / accumulator
    // committed / epilogue start / epilogue end)
#ifndef MFM_TL_PAIR
#define MFM_TL_PAIR 0
#endif
#define TC2P_MARK(tile, ev) do { if (tl && blockIdx.x == 2 * MFM_TL_PAIR && (tile) < 12 && lane == 0) tl[(tile) * 4 + (ev)] = clock64(); } while (0)
#define TC2P_MARKX(slot) do { if (tl && blockIdx.x == 2 * MFM_TL_PAIR && lane == 0) tl[slot] = clock64(); } while (0)
#else
#define TC2P_MARK(tile, ev) do { } while (0)
#define TC2P_MARKX(slot) do { } while (0)
#endif
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stg_base = smem + S_STAGES * S_STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(stg_base + EPI_STG_BYTES);
    uint64_t* full = bars;                      // own TMA landed                      (local)
    uint64_t* split = bars + S_STAGES;            // lo tiles of BOTH CTAs ready          (leader's copy is used)
    uint64_t* empty = bars + 2 * S_STAGES;        // MMAs done reading the stage          (multicast commit)
    uint64_t* acc_full = bars + 3 * S_STAGES;     // [2] accumulator buffer complete      (multicast commit)
    uint64_t* acc_empty = bars + 3 * S_STAGES + 2;// [2] buffer drained by BOTH CTAs      (leader's copy is used)
    uint32_t* tmem_slot = (uint32_t*)(bars + S_NBARS);
    volatile PairWork* work = (volatile PairWork*)(bars + S_NBARS + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int M = p.n_rows_dev ? min(*p.n_rows_dev, p.M) : p.M;
    const Tiles T = make_tiles(M, p.N, p.K, p.k_split);
    if (T.total == 0) return;                   // uniform over the grid
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
#ifdef MFM_TC2_TIMELINE
    if (tl && blockIdx.x == 2 * MFM_TL_PAIR && threadIdx.x == 0) tl[62] = clock64();
#endif

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b) : "memory");
        // split[s]: the stage's operands are in tensor-core format in BOTH CTAs.  With everything pre-split (APRE and BPRE) one warp
        // per CTA just forwards its own `full` to the leader; otherwise the splitter warps of the stage's group arrive
        const uint32_t split_count = (APRE && BPRE) ? 2u : (uint32_t)(2 * SPLIT_WARPS / sc.split_groups);
        for (int s = 0; s < S_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], split_count); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 2 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (warp == 3 && lane == 0) {
        const Sched S = make_sched(T.total, n_pairs, pair, p.K, p.k_split, sk_ws != nullptr);
        Item head, last;
        const int have = S.tail(head, last);
        if (have & 1) { volatile Item* d = &work->head; d->tile = head.tile; d->kb0 = head.kb0; d->kb1 = head.kb1; d->kind = head.kind; d->c_first = 0; d->c_count = 0; }
        if (have & 2) { volatile Item* d = &work->last; d->tile = last.tile; d->kb0 = last.kb0; d->kb1 = last.kb1; d->kind = last.kind; d->c_first = last.c_first; d->c_count = last.c_count; }
        work->n_head = have & 1; work->n_tail = (have & 1) + ((have >> 1) & 1); work->n_items = (have & 1) + ((have >> 1) & 1) + S.whole();
        work->fix = ((have & 2) && last.kind == ITEM_FINISH) ? 1 : 0;
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // item i of this pair -> tile and k-block range (the list stays in shared memory: the epilogue has no registers to spare)
    auto item_at = [&](int i, int& tile, int& kb0, int& kb1) {
        const int j = i - work->n_tail;
        if (j >= 0) { tile = pair + j * n_pairs; kb0 = 0; kb1 = (p.K + BK - 1) / BK; }
        else { volatile const Item* w = i < work->n_head ? &work->head : &work->last; tile = w->tile; kb0 = w->kb0; kb1 = w->kb1; }
    };

    // tile t -> (z, row tile, column tile); column tiles vary fastest so concurrent pairs share A rows
    auto tile_origin = [&](int t, int& m0p, int& n0, int& kz0, int& KT, int& neff, int& z) {
        const int per_z = T.m_tiles * T.n_tiles;
        z = t / per_z;
        const int r = t - z * per_z;
        m0p = (r / T.n_tiles) * (2 * BM);
        n0 = (r % T.n_tiles) * BN;
        kz0 = p.k_split > 0 ? z * p.k_split : 0;
        const int Kend = p.k_split > 0 ? min(p.K, kz0 + p.k_split) : p.K;
        KT = (Kend - kz0 + BK - 1) / BK;
        const int nrem = p.N - n0;
        neff = nrem >= BN ? BN : ((nrem + 63) / 64) * 64;
    };

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs) ----------------
        if (lane == 0) {
            uint32_t g = 0;                     // k-blocks issued so far (ring position)
            for (int idx = 0; idx < work->n_items; ++idx) {
                int tile, kt0, kt1, m0p, n0, kz0, KT, neff, z;
                item_at(idx, tile, kt0, kt1);
                tile_origin(tile, m0p, n0, kz0, KT, neff, z);
                const int m0 = m0p + (int)rank * BM;
                const int nb0 = n0 + (int)rank * (neff / 2);
                if (p.k_split > 0) { kt0 = 0; kt1 = KT; }
                for (int kt = kt0; kt < kt1; ++kt, ++g) {
                    const uint32_t s = g % S_STAGES, ph = (g / S_STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * S_STAGE_BYTES;
                    mbar_expect_tx(&full[s], HI_BYTES);
                    const int k0 = kz0 + kt * BK;
                    tma_load_2d(st, &maps.a, &full[s], k0, m0);                    // rows of A: raw fp32 (split in place below) or pre-split (APRE)
                    tma_load_2d(st + A_BYTES, &maps.b, &full[s], k0, nb0);         // B half-tile: pre-split [16 hi | 16 lo] per 16 k (BPRE) or raw
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (rank == 0 && lane == 0) {
            uint32_t g = 0, i = 0;
            for (; (int)i < work->n_items; ++i) {
                int tile, kt0, kt1, m0p, n0, kz0, KT, neff, z;
                item_at((int)i, tile, kt0, kt1);
                tile_origin(tile, m0p, n0, kz0, KT, neff, z);
                if (p.k_split > 0) { kt0 = 0; kt1 = KT; }
                const uint32_t b = i & 1, u = i >> 1;
                // kind::f16: D = f32 (bit 4), A = B = fp16 (format 0 in bits 7-9 / 10-12), both K-major, M = 256 per pair, N = neff
                const uint32_t idesc16 = (1u << 4) | ((uint32_t)(neff >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
                const uint32_t acc = tmem_base + b * BN;
                mbar_wait(&acc_empty[b], (u & 1) ^ 1);          // both CTAs have drained this buffer (tile i-2)
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                TC2P_MARK(i, 0);
                for (int kt = kt0; kt < kt1; ++kt, ++g) {
                    const uint32_t s = g % S_STAGES, ph = (g / S_STAGES) & 1;
                    mbar_wait(&split[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_t = smem_u32(smem + s * S_STAGE_BYTES), b_t = a_t + A_BYTES;
                    // a 128-byte row of either tile = [hi(k 0..15) | lo(k 0..15) | hi(k 16..31) | lo(k 16..31)], 32 bytes each:
                    // per 16 k-values three K = 16 MMAs  hi.hi' + hi.lo' + lo.hi'
#pragma unroll
                    for (int gq = 0; gq < 2; ++gq) {
                        const uint64_t dah = make_desc(a_t + gq * 64, 16, 1024, 2), dal = make_desc(a_t + gq * 64 + 32, 16, 1024, 2);
                        const uint64_t dbh = make_desc(b_t + gq * 64, 16, 1024, 2), dbl = make_desc(b_t + gq * 64 + 32, 16, 1024, 2);
                        mma_f16_ss_2sm(acc, dah, dbh, idesc16, ((kt - kt0) | gq) != 0);
                        mma_f16_ss_2sm(acc, dah, dbl, idesc16, 1);
                        mma_f16_ss_2sm(acc, dal, dbh, idesc16, 1);
                    }
                    mma_commit_2sm(&empty[s]);
                }
                mma_commit_2sm(&acc_full[b]);
                TC2P_MARK(i, 1);
            }
        }
    } else if (warp >= SPLIT_WARP0 && warp < EPI_WARP0) {
        // ---------------- splitters (both CTAs) ----------------
        // The 8 warps work as `split_groups` groups that take the stages in turn (group = stage number mod groups): a group's
        // threads each convert `split_groups` 64-byte units of the tile.  One stage costs a fixed latency (barrier wake-up,
        // shared-memory round trip, fence.proxy.async, the remote arrive) plus the conversion work; two groups keep two stages in
        // flight, so the fixed part overlaps and the splitters keep up with the 768-clock MMA time of a k-block.
        const int G = (APRE && BPRE) ? 1 : sc.split_groups, per_group = SPLIT_WARPS * 32 / G;
        const int tix = threadIdx.x - SPLIT_WARP0 * 32;
        const int grp = tix / per_group, tg = tix - grp * per_group;
        // raw operands are scaled while they are split: A by the power of two of its maximum, B (no mirror) likewise
        const float s_a = APRE ? 1.0f : h16_scale(h16_a_amax(sc)), s_b = BPRE ? 1.0f : h16_scale(*sc.b_amax);
        uint32_t g = 0;
        for (int idx = 0; idx < work->n_items; ++idx) {
            int tile, kb0, kb1;
            item_at(idx, tile, kb0, kb1);
            int n_kb = kb1 - kb0;
            if (p.k_split > 0) { int m0p, n0, kz0, KT, neff, z; tile_origin(tile, m0p, n0, kz0, KT, neff, z); n_kb = KT; }
            for (int kt = 0; kt < n_kb; ++kt, ++g) {
                if ((int)(g % (uint32_t)G) != grp) continue;
                const uint32_t s = g % S_STAGES, ph = (g / S_STAGES) & 1;
                mbar_wait(&full[s], ph);
                if (APRE && BPRE) {
                    // nothing to convert: TMA delivered tensor-core format; one warp tells the leader that this CTA's half is in.
                    // The other warps still WALK the stages: they become epilogue helpers of the pair's last tile afterwards, and
                    // their parity wait on that tile's accumulator barrier is only unambiguous once the tile's own stages have
                    // landed (the buffer's previous use is then complete) - starting early would read the wrong phase.
                    if (warp == SPLIT_WARP0 && lane == 0) mbar_arrive_remote(&split[s], 0);
                    continue;
                }
                // in-place split of the raw tile(s): unit (row r, half hq) = the 16 floats k = 16 hq .. 16 hq + 15 of row r, i.e.
                // logical 16-byte chunks 4 hq .. 4 hq + 3 (physical chunk = logical ^ (r & 7), SWIZZLE_128B), replaced by
                // 16 hi | 16 lo fp16 parts of the SCALED values.  Units 0..255: the A tile, 256..511: the B half-tile.
#pragma unroll 1
                for (int un = (APRE ? 256 : 0) + tg; un < (BPRE ? 256 : 512) && !(APRE && BPRE); un += per_group) {
                    const int op = un >> 8;
                    const uint32_t r = ((uint32_t)un & 255u) >> 1, hq = (uint32_t)un & 1u;
                    const uint32_t rowa = smem_u32(smem + s * S_STAGE_BYTES) + (op ? A_BYTES : 0) + r * 128u;
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w)
                                     : "r"(rowa + (((4u * hq + (uint32_t)j) ^ (r & 7u)) << 4)));
                    uint32_t hp[8], lp[8];
                    if (sc.cvt_unpack) split16<true>(v, op ? s_b : s_a, hp, lp); else split16<false>(v, op ? s_b : s_a, hp, lp);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((4u * hq + (uint32_t)j) ^ (r & 7u)) << 4)),
                                     "r"(hp[4 * j]), "r"(hp[4 * j + 1]), "r"(hp[4 * j + 2]), "r"(hp[4 * j + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((4u * hq + 2u + (uint32_t)j) ^ (r & 7u)) << 4)),
                                     "r"(lp[4 * j]), "r"(lp[4 * j + 1]), "r"(lp[4 * j + 2]), "r"(lp[4 * j + 3]) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&split[s], 0);
            }
        }
    }
    if (warp >= SPLIT_WARP0) {
        // ---------------- epilogue (each CTA drains its own 128 TMEM lanes) ----------------
        // Epilogue warps 12-19: (TMEM lane quadrant, column half) = 4 chunks of 32 columns per tile.  The
        // splitter warps 4-11 have nothing left to do once the pair's LAST tile is split, so they take
        // half of that tile's chunks: its epilogue is the only one no main loop hides.
        const bool helper = warp < EPI_WARP0;
        const int ew = helper ? warp - SPLIT_WARP0 : warp - EPI_WARP0;   // 0..7
        const int quad = ew & 3;                      // == warp & 3: the TMEM lane quadrant this warp may access
        const int chalf = ew >> 2;                    // column half of the tile
        // staging: the epilogue warps' own buffers; helpers use ring stage 0 (free: every MMA has completed)
        const uint32_t stg = helper ? smem_u32(smem) + (uint32_t)ew * 4096u : smem_u32(stg_base) + (uint32_t)ew * 4096u;
        const int rsub = lane >> 3, cpiece = (lane & 7) * 4;
        constexpr int RB = sizeof(typename Epi::Row4) > 36 ? 2 : 4;   // steps whose global reads are issued together (register budget: 102)
        // exact un-scaling: the accumulator holds (s_a s_b) A.B; a pre-split A was scaled by its producer with the power of two of *a_scale_src
        const float inv_s = h16_inv_scale(APRE ? *sc.a_scale_src : h16_a_amax(sc)) * h16_inv_scale(*sc.b_amax);
        float vmax = 0.0f;                                            // max |value stored| by this thread (EpiStd::amax_out)
        if (work->n_head != 0 && (!helper || work->n_items == 1)) {
            // ---- stream-K contribution (item 0, accumulator buffer 0): dump to this pair's slot, flag per chunk ----
            int tile, kb0_, kb1_, m0p, n0, kz0, KT, neff, z;
            item_at(0, tile, kb0_, kb1_);
            tile_origin(tile, m0p, n0, kz0, KT, neff, z);
            const int n_chunks = min(BN / 32, (p.N - n0 + 31) / 32);
            int cc_begin = chalf * 4, cc_end = min(n_chunks, chalf * 4 + 4);
            if (work->n_items == 1) { if (helper) cc_begin = min(cc_begin + 2, cc_end); else cc_end = min(cc_end, cc_begin + 2); }
            mbar_wait(&acc_full[0], 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ew == 0 && !helper) TC2P_MARKX(48);
            if (cc_begin >= cc_end && !helper) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&acc_empty[0], 0);
            }
#pragma unroll 1
            for (int cc = cc_begin; cc < cc_end; ++cc) {
                uint32_t r[32];
                tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cc * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc == cc_end - 1 && !helper) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&acc_empty[0], 0);
                }
                const uint32_t dst = stg + (uint32_t)lane * 128u;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)((j ^ (lane & 7)) << 4)), "f"(__uint_as_float(r[4 * j])),
                                 "f"(__uint_as_float(r[4 * j + 1])), "f"(__uint_as_float(r[4 * j + 2])), "f"(__uint_as_float(r[4 * j + 3])) : "memory");
                __syncwarp();
                // this warp's 32 x 32 chunk goes to rows rank*128 + quad*32 .., columns cc*32 .. of the slot, moved like C
                // itself (4 rows x 128 contiguous bytes per instruction)
                float* dstp = sk_ws + (size_t)pair * SK_SLOT_FLOATS + (size_t)((int)rank * BM + quad * 32 + rsub) * BN + cc * 32 + cpiece;
                const uint32_t sa = stg + (uint32_t)rsub * 128u;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                                 : "r"(sa + (uint32_t)k * 512u + (uint32_t)(((lane & 7) ^ ((k * 4 + rsub) & 7)) << 4)));
                    __stcg(reinterpret_cast<float4*>(dstp + (size_t)k * 4 * BN), v);
                }
            }
            // one release for the warp's chunks: every lane's stores -> __syncwarp -> lane 0's fence -> the flags
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                for (int cc = cc_begin; cc < cc_end; ++cc)
                    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(sk_flags + (size_t)pair * SK_SLOT_FLAGS + rank * 32u + (uint32_t)quad * 8u + (uint32_t)cc),
                                 "r"(sk_epoch) : "memory");
            }
        }
        if (ew == 0 && !helper) TC2P_MARKX(49);
        // the finishing part and the whole tiles (helpers: the last item only, unless that was the contribution)
        uint32_t i = (uint32_t)(helper ? max(work->n_items - 1, work->n_head) : work->n_head);
        for (; (int)i < work->n_items; ++i) {
            int tile, kb0_, kb1_, m0p, n0, kz0, KT, neff, z;
            item_at((int)i, tile, kb0_, kb1_);
            tile_origin(tile, m0p, n0, kz0, KT, neff, z);
            const uint32_t b = i & 1, u = i >> 1;
            const bool last = (int)i == work->n_items - 1;
            const bool fix = work->fix != 0 && (int)i == work->n_head;      // stream-K FINISH item: add the earlier k-ranges
            Epi e = epi;
            if (p.k_split > 0) e.at_z(z);
            const int row_base = m0p + (int)rank * BM + quad * 32;
            const int n_chunks = min(BN / 32, (p.N - n0 + 31) / 32);     // chunks of the whole tile
            int cc_begin = chalf * 4, cc_end = min(n_chunks, chalf * 4 + 4);
            if (last) { if (helper) cc_begin = min(cc_begin + 2, cc_end); else cc_end = min(cc_end, cc_begin + 2); }
            mbar_wait(&acc_full[b], u & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ew == 0 && !helper) TC2P_MARK(i, 2);
            float rs[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) rs[k] = 0.0f;
            if (cc_begin >= cc_end && !helper) {
                // nothing to drain in this column half (edge tile): still release the buffer
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&acc_empty[b], 0);
            }
#pragma unroll 1
            for (int cc = cc_begin; cc < cc_end; ++cc) {
                const int col0 = cc * 32;
                uint32_t r[32];
                tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + b * BN + (uint32_t)col0, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc == cc_end - 1 && !helper) {
                    // this warp's last chunk has left TMEM: hand the buffer back to the MMA issuer (on the
                    // pair's last tile nobody waits for it any more, so the helpers' chunks need no arrive)
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&acc_empty[b], 0);
                }
                const uint32_t dst = stg + (uint32_t)lane * 128u;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)((j ^ (lane & 7)) << 4)), "f"(__uint_as_float(r[4 * j])),
                                 "f"(__uint_as_float(r[4 * j + 1])), "f"(__uint_as_float(r[4 * j + 2])), "f"(__uint_as_float(r[4 * j + 3])) : "memory");
                __syncwarp();
                if (fix) {
                    // stream-K: add this chunk of the contributing pairs' slots into the staged accumulator, moved like C itself
                    // (4 rows x 128 contiguous bytes per instruction; each lane updates exactly the pieces it reads back below)
                    const int c_first = work->last.c_first, c_count = work->last.c_count;
                    const uint32_t fidx = rank * 32u + (uint32_t)quad * 8u + (uint32_t)cc;
                    if (lane < c_count) {                     // one lane per contributing pair
                        const unsigned* fl = sk_flags + (size_t)(c_first + lane) * SK_SLOT_FLAGS + fidx;
                        unsigned seen;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(fl) : "memory");
                        } while (seen != sk_epoch);
                    }
                    __syncwarp();
                    const size_t toff = (size_t)((int)rank * BM + quad * 32 + rsub) * BN + col0 + cpiece;
                    const uint32_t sa = stg + (uint32_t)rsub * 128u;
#pragma unroll 1
                    for (int c = 0; c < c_count; ++c) {
                        const float* srcp = sk_ws + (size_t)(c_first + c) * SK_SLOT_FLOATS + toff;
                        float4 v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = __ldcg(reinterpret_cast<const float4*>(srcp + (size_t)k * 4 * BN));
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t ad = sa + (uint32_t)k * 512u + (uint32_t)(((lane & 7) ^ ((k * 4 + rsub) & 7)) << 4);
                            float4 w;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "r"(ad));
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ad), "f"(w.x + v[k].x), "f"(w.y + v[k].y), "f"(w.z + v[k].z), "f"(w.w + v[k].w) : "memory");
                        }
                    }
                }
                const int col = n0 + col0 + cpiece;           // N % 4 == 0: the four columns are valid together
                const bool cvalid = col < p.N;
                typename Epi::Col4 ca;
                if (cvalid) ca = e.load_col4(col);
#pragma unroll
                for (int it0 = 0; it0 < 8; it0 += RB) {
                    typename Epi::Row4 ra[RB];
                    float4 acc[RB];
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
                        const int row = row_base + (it0 + k) * 4 + rsub;
                        if (row < M && cvalid) ra[k] = e.load_row4(row, col);
                    }
#pragma unroll
                    for (int k = 0; k < RB; ++k)
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(acc[k].x), "=f"(acc[k].y), "=f"(acc[k].z), "=f"(acc[k].w)
                                     : "r"(stg + (uint32_t)((it0 + k) * 4 + rsub) * 128u + (uint32_t)((((lane & 7) ^ (((it0 + k) * 4 + rsub) & 7))) << 4)));
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
                        const int row = row_base + (it0 + k) * 4 + rsub;
                        float c = 0.0f;
                        acc[k].x *= inv_s; acc[k].y *= inv_s; acc[k].z *= inv_s; acc[k].w *= inv_s;
                        if (row < M && cvalid) c = e.apply4(row, col, acc[k], ca, ra[k]);
                        if (Epi::kRowSum) {
                            c += __shfl_xor_sync(0xffffffffu, c, 4); c += __shfl_xor_sync(0xffffffffu, c, 2);
                            c += __shfl_xor_sync(0xffffffffu, c, 1);
                            rs[it0 + k] += c;               // 32-column sums; pairs of chunks make the 64-column groups
                        }
                    }
                }
                if constexpr (epi_has_colsum<Epi>::value) {
                    // column sums of this 32-row x 32-column chunk: the four row groups of a column piece live 8 lanes apart
                    float4 s4 = e.cs;
                    s4.x += __shfl_xor_sync(0xffffffffu, s4.x, 8); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, 8);
                    s4.z += __shfl_xor_sync(0xffffffffu, s4.z, 8); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, 8);
                    s4.x += __shfl_xor_sync(0xffffffffu, s4.x, 16); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, 16);
                    s4.z += __shfl_xor_sync(0xffffffffu, s4.z, 16); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, 16);
                    if (rsub == 0 && cvalid && row_base < M && e.colsum_part)
                        *reinterpret_cast<float4*>(e.colsum_part + (long long)(row_base >> 5) * e.ldcs + col) = s4;
                    e.cs = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
                if (Epi::kRowSum && ((cc & 1) || cc == cc_end - 1)) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int row = row_base + k * 4 + rsub;
                        if ((lane & 7) == 0 && row < M) e.row_partial(row, (n0 + (cc & ~1) * 32) / GBN, rs[k]);
                        rs[k] = 0.0f;
                    }
                }
                __syncwarp();                                 // staging is reused by the next chunk
            }
            if (ew == 0 && !helper) TC2P_MARK(i, 3);
            vmax = fmaxf(vmax, epi_stored_max(e, 0));
        }
        epi_publish_amax(epi, vmax);
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                     // the peer's shared memory / TMEM / barriers stay alive until both are done
#ifdef MFM_TC2_TIMELINE
    if (tl && blockIdx.x == 2 * MFM_TL_PAIR && threadIdx.x == 0) tl[63] = clock64();
#endif
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- host side ----------------------------------------------------------------------------------
int split_groups();                         // tuning switch (env MFM_H16_GROUPS=1|2|4, default 2)
int gemm_h16();                             // 1 (default): K-major x K-major dense layers with 16-aligned K run here (env MFM_GEMM_H16=0|1)
// per-stream scratch of 64 device floats for the maxima of operands nobody tracked (rng.cu)
float* amax_scratch(cudaStream_t st);
// pre-split mirror registered for the weight buffer that contains `p` (and the device slot holding its max |w|), or null
const float* lookup_mirror_h16(const float* p, const float** amax);
cudaError_t launch_absmax(const float* x, long long ld, int rows, int cols, const int* n_rows_dev, float* out, cudaStream_t st);

template <class Epi>
inline bool eligible(const GemmShape& p, const Epi& epi) {
    // operands nobody tracked a maximum for cost one reduction pass: only worth it for layers that are tensor-bound
    const float* bam = nullptr;
    const bool a_tracked = p.a_amax != nullptr || p.a_bound > 0.0f || (p.a_split != nullptr && p.a_scale_src != nullptr);
    const bool b_ready = p.b_mirror != nullptr || lookup_mirror_h16(p.B, &bam) != nullptr;      // else B needs its own reduction + in-kernel split
    if (!(a_tracked && b_ready) && (long long)p.N * p.K < 512ll * 512ll) return false;
    return tc2p::eligible<true, false>(p, epi) && p.k_split == 0 && p.K % 16 == 0 && p.lda % 16 == 0 && p.ldb % 16 == 0 &&
           ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.B)) & 63) == 0;
}

template <class Epi>
inline cudaError_t launch(const GemmShape& p, const Epi& epi, cudaStream_t st) {
    const bool apre = p.a_split != nullptr && p.a_scale_src != nullptr && (reinterpret_cast<uintptr_t>(p.a_split) & 63) == 0;
    H16Scales sc{p.a_amax, p.a_amax2, p.a_bound, p.b_amax, apre ? p.a_scale_src : nullptr, gemm_h16() == 2 ? 1 : 0, split_groups()};
    float* scratch = nullptr;
    if (!apre && sc.a_bound <= 0.0f && !sc.a_amax) {
        // nobody tracked this operand's maximum: one reduction pass over it (HBM-bound, ~10 % of the GEMM at 65 536 rows)
        if (!(scratch = amax_scratch(st))) return cudaErrorMemoryAllocation;
        cudaError_t e = launch_absmax(p.A, p.lda, p.M, p.K, p.n_rows_dev, scratch, st);
        if (e != cudaSuccess) return e;
        sc.a_amax = scratch; sc.a_amax2 = nullptr;
    }
    const float* b_amax = sc.b_amax;
    const float* bx = (p.b_mirror && p.b_amax) ? p.b_mirror : lookup_mirror_h16(p.B, &b_amax);
    if (!bx && !sc.b_amax) {
        if (!scratch && !(scratch = amax_scratch(st))) return cudaErrorMemoryAllocation;
        cudaError_t e = launch_absmax(p.B, p.ldb, p.N, p.K, nullptr, scratch + 1, st);
        if (e != cudaSuccess) return e;
        b_amax = scratch + 1;
    }
    sc.b_amax = b_amax;
    Maps3 maps;
    bool ok = tc::make_map_kmajor(&maps.a, apre ? p.a_split : p.A, p.lda, p.M, p.K, BM) && tc::make_map_kmajor(&maps.b, bx ? bx : p.B, p.ldb, p.N, p.K, BNH);
    if (!ok) return cudaErrorInvalidValue;
    maps.bx = maps.b;
    auto kern = apre ? (bx ? gemm_tc2h_kernel<true, true, Epi> : gemm_tc2h_kernel<true, false, Epi>)
                     : (bx ? gemm_tc2h_kernel<false, true, Epi> : gemm_tc2h_kernel<false, false, Epi>);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc2h_kernel<true, true, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc2h_kernel<true, false, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc2h_kernel<false, true, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc2h_kernel<false, false, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const Tiles T = make_tiles(p.M, p.N, p.K, 0);
    float* sk_ws = nullptr; unsigned* sk_flags = nullptr; unsigned sk_epoch = 0;
    const int rem = T.total % sm_pairs();
    const bool sk = (p.K + BK - 1) / BK >= SK_MAX_SPLIT * SK_MIN_KB && (p.n_rows_dev || (rem != 0 && rem * SK_MAX_SPLIT <= sm_pairs())) &&
                    streamk_workspace(st, &sk_ws, &sk_flags, &sk_epoch);
    const int pairs = (sk || T.total >= sm_pairs()) ? sm_pairs() : T.total;
    kern<<<dim3(2 * pairs), THREADS, S_SMEM_BYTES, st>>>(maps, p, epi, sc, tc2::gemm_timeline(), sk_ws, sk_flags, sk_epoch);
    ++g_mfm_launches;
    return cudaGetLastError();
}

}  // namespace tc2h
}  // namespace mfm
