// Persistent CTA-pair 3xTF32 GEMM: gemm_tcgen05_2sm.cuh's pair tile (256 x 256, cta_group::2, in-kernel
// hi/lo split) with the epilogue of tile i overlapped with the main loop of tile i+1.
//
// Measured on B200 (scripts/gemm_bench.py --timeline) for the one-tile-per-pair kernel at K = 1024:
// main loop 48.3 k clocks (1 510 clk per 32-wide k-block = 98 % of the kind::tf32 issue rate), but
// 4.1 k prologue + pipeline fill, 1.3 k drain, 11.7 k epilogue (all SMs write C in lock-step, so the
// phase is HBM-write bound) and 1.6 k teardown per tile: 28 % of the tile is not MMA.  Here one pair
// per TPC stays resident and walks a static tile list:
//   warp 0        TMA producer          (runs ahead across tiles; 3-stage ring, 64 KB per stage)
//   warp 1        MMA issuer (leader)   accumulator buffer b = tile & 1 (TMEM columns b*256 .. +255)
//   warp 2        TMEM allocator
//   warps 4-11    splitters             hi (raw, truncated by the tensor core) / lo = rn_tf32(x - trunc x)
//   warps 12-19   epilogue              drain buffer b while the MMAs fill buffer b^1 (quadrant x column half)
// TMEM holds 512 columns, so two 256-column buffers leave no room for a separate cross-term
// accumulator: lo*hi, hi*lo and hi*hi of a k-step accumulate into the same buffer.  The tensor core's
// fp32 accumulator truncates on every add, so this costs accuracy: measured bias ~ 7.5e-9 * K relative
// (separate accumulators: 2.5e-9 * K).  MFM_GEMM=tc2 selects the one-tile kernel when that matters.
//
// Epilogue: a warp owns 32 rows (its TMEM lane quadrant) x 128 columns, drained in 32-column chunks:
// tcgen05.ld -> shared memory (XOR-swizzled) -> 8 steps of 4 rows x 128 bytes, float4 per lane, through
// the functors' Col4/Row4 interface (gemm_tf32x3.cuh).  Buffer b is released (remote arrive on the
// leader's acc_empty[b]) as soon as a warp's last chunk has left TMEM.
// Measured (scripts/epi_bench.py --timeline): the shared-memory / L1 data pipe is ~90 % busy with MMA
// operand reads and the splitters, so a global load issued by an epilogue warp returns after ~2.5 k clocks
// whether it hits L2 or not (an L2 prefetch one tile ahead changed nothing and was removed); the epilogue
// is bounded by loads in flight, hence 8 epilogue warps.  Per 256x256 tile at K = 1024: MMAs 41 k clocks,
// epilogue 17 k (bias only) / 40 k (with a mask or residual operand).
#pragma once
#include "gemm_tcgen05_2sm.cuh"

namespace mfm {
namespace tc2p {

using tc::smem_u32; using tc::mbar_init; using tc::mbar_expect_tx; using tc::mbar_wait; using tc::tma_load_2d;
using tc::tma_load_3d; using tc::tmem_ld32_nowait; using tc::make_desc; using tc::Maps;
using tc2::cluster_ctarank; using tc2::cluster_sync_all; using tc2::mbar_arrive_remote; using tc2::mma_tf32_ss_2sm;
using tc2::mma_commit_2sm; using tc2::make_idesc_base;

constexpr int BM = 128, BN = 256, BNH = BN / 2, BK = 32, STAGES = 3;
constexpr int THREADS = 640;                // 20 warps
constexpr int SPLIT_WARP0 = 4, SPLIT_WARPS = 8, EPI_WARP0 = 12, EPI_WARPS = 8;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BNH * BK * 4;
constexpr int HI_BYTES = A_BYTES + B_BYTES; // 32 KB
constexpr int STAGE_BYTES = 2 * HI_BYTES;   // 64 KB
// epilogue staging: one 32x32 fp32 chunk per warp, 128-byte rows, 16-byte pieces XOR-swizzled by the row
// (piece j of row r at r*128 + ((j ^ (r & 7)) << 4)): conflict-free for the row-per-lane float4 writes and
// for the 4-rows-per-instruction float4 reads, with no padding (8 warps x 4 KB must fit beside the ring)
constexpr int EPI_STG_BYTES = EPI_WARPS * 32 * 32 * 4;          // 32 KB
constexpr int NBARS = 3 * STAGES + 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;
static_assert(NBARS * 8 + 8 + 64 <= 256, "barrier area: barriers, TMEM slot, PairWork");

struct Tiles {
    int m_tiles, n_tiles, z_tiles, total;
};
__host__ __device__ inline Tiles make_tiles(int M, int N, int K, int k_split) {
    Tiles t;
    t.m_tiles = (M + 2 * BM - 1) / (2 * BM);
    t.n_tiles = (N + BN - 1) / BN;
    t.z_tiles = k_split > 0 ? (K + k_split - 1) / k_split : 1;
    t.total = t.m_tiles * t.n_tiles * t.z_tiles;
    return t;
}

// ---- work list of one CTA pair ---------------------------------------------------------------------
// Rounds of whole tiles (pair p takes tiles p, p + P, ...) plus the REMAINDER round of rem < P tiles.
// When the remainder is small (rem * 4 <= P: e.g. 8 192 chains x 1 600 columns = 224 tiles = 3 rounds + 2
// tiles, or the few active chains at the end of an ODE solve) its rem * KT k-blocks are cut into
// Pp = 4 * rem equal contiguous ranges (stream-K): the round then lasts a quarter of a tile time instead of
// a whole one.  A pair's range covers at most the END of one tile and the BEGINNING of the next one:
//   * the beginning (CONTRIB) is the pair's FIRST item: its accumulator is dumped to the pair's workspace
//     slot (256 x 256 fp32, L2 resident) and flagged per 32 x 32 chunk.  Contributions wait for nothing,
//     so there is no dependency chain between pairs (every pair of the grid is co-resident);
//   * the end (FINISH) comes second: its epilogue adds, chunk by chunk and in pair order, the slots of the
//     pairs that hold the tile's earlier k-ranges before the functor runs;
//   * the whole tiles follow, so the fix-up epilogue (global loads at ~2.5 k clocks each) hides under
//     their main loops like any other epilogue.
// Summation order is fixed => results are deterministic.
// Measured on B200 (scripts/streamk_bench.py, scripts/streamk_timeline.py): 1.04-1.17x on the layers it
// applies to.  NOT applied to large remainders (54 tiles on 74 pairs): there the round is bounded by the
// epilogue chain (dump 13 k + fix-up + 18 k epilogue clocks against 41 k of MMAs), the short items pay a
// pipeline fill each and the dump / fix-up traffic slows the concurrent MMAs by 15 % - three orderings
// (fix-up in the epilogue, remainder first; fix-up pre-loaded into TMEM with tcgen05.st, remainder last)
// all came out 0.7-0.96x of whole tiles.
enum { ITEM_FULL = 0, ITEM_CONTRIB = 1, ITEM_FINISH = 2 };
struct Item { int tile, kb0, kb1, kind, c_first, c_count; };
struct Sched {
    int P, KT, n_full, rem, Pp, U, pair;
    __host__ __device__ __forceinline__ int lo(int q) const { return (int)(((long long)q * U) / Pp); }
    // this pair's share of the remainder: `head` (a CONTRIB item, runs first) and / or `last` (the part that
    // reaches a tile's final k-block, runs last); returns bit 0 = head present, bit 1 = last present
    __host__ __device__ __forceinline__ int tail(Item& head, Item& last) const {
        if (Pp == 0 || pair >= Pp) return 0;
        const int u0 = lo(pair), u1 = lo(pair + 1), base = n_full * P;
        if (u1 <= u0) return 0;
        const int ta = u0 / KT, tb = (u1 - 1) / KT;
        int have = 0;
        if (ta != tb || u1 - ta * KT < KT) {         // the range ends inside tile tb: contribution
            head.tile = base + tb; head.kb0 = ta == tb ? u0 - ta * KT : 0; head.kb1 = u1 - tb * KT;
            head.kind = ITEM_CONTRIB; head.c_first = 0; head.c_count = 0;
            have |= 1;
        }
        if (ta != tb || u1 - ta * KT == KT) {        // the range reaches the end of tile ta
            last.tile = base + ta; last.kb0 = u0 - ta * KT; last.kb1 = KT; last.kind = ITEM_FULL; last.c_first = 0; last.c_count = 0;
            if (last.kb0 > 0) {                      // pairs q < pair whose ranges reach into this tile
                last.kind = ITEM_FINISH;
                int q = pair - 1;
                while (q > 0 && lo(q) > ta * KT) --q;
                last.c_first = q; last.c_count = pair - q;
            }
            have |= 2;
        }
        return have;
    }
    // number of whole tiles of this pair (without stream-K the remainder round is one more whole tile)
    __host__ __device__ __forceinline__ int whole() const { return n_full + ((Pp == 0 && pair < rem) ? 1 : 0); }
    // item `idx` of this pair; false past the end
    __host__ __device__ __forceinline__ bool get(int idx, Item& it) const {
        Item head, last;
        const int have = tail(head, last), nh = have & 1, nl = (have >> 1) & 1, nw = whole();
        if (idx < nh) { it = head; return true; }
        if (idx < nh + nl) { it = last; return true; }
        if (idx - nh - nl < nw) { it.tile = pair + (idx - nh - nl) * P; it.kb0 = 0; it.kb1 = KT; it.kind = ITEM_FULL; it.c_first = 0; it.c_count = 0; return true; }
        return false;
    }
};
// what the roles read (shared memory, written once in the prologue)
struct PairWork { int n_head, n_tail, n_items, fix; Item head, last; };   // n_tail = n_head + (last present); fix: `last` needs a fix-up
constexpr int SK_MAX_SPLIT = 4, SK_MIN_KB = 4;
__host__ __device__ __forceinline__ Sched make_sched(int total, int n_pairs, int pair, int K, int k_split, bool streamk) {
    Sched s;
    s.P = n_pairs; s.pair = pair; s.KT = (K + BK - 1) / BK;
    s.n_full = total / n_pairs; s.rem = total - s.n_full * n_pairs; s.Pp = 0; s.U = 0;
    // small remainders only (see above): every remainder tile is cut SK_MAX_SPLIT ways and every range keeps >= SK_MIN_KB k-blocks
    if (streamk && k_split == 0 && s.rem > 0 && s.rem * SK_MAX_SPLIT <= n_pairs && s.KT >= SK_MAX_SPLIT * SK_MIN_KB) {
        s.Pp = s.rem * SK_MAX_SPLIT; s.U = s.rem * s.KT;
    }
    return s;
}
constexpr int SK_SLOT_FLOATS = 2 * BM * BN;     // one pair tile
constexpr int SK_SLOT_FLAGS = 2 * 4 * (BN / 32);// (CTA, TMEM lane quadrant, 32-column chunk)
static_assert(SK_SLOT_FLOATS == 256 * 256 && SK_SLOT_FLAGS == 64, "rng.cu::streamk_workspace sizes its buffers with these");

__device__ __forceinline__ void mma_bf16_ss_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
// two floats -> packed bf16x2 (round to nearest even); `lo16` lands in the low half (first in memory)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo16, float hi16) {
    uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi16), "f"(lo16)); return r;
}
__device__ __forceinline__ float tf32_trunc_rest(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// XBF16 (K-major x K-major operands only): the two cross terms of a k-step are ONE kind::f16 MMA with
// K = 16 over a bf16 tile the splitters build next to the raw fp32 tile:
//     A' = [ bf16(a_lo[0..7]) | bf16(a[0..7]) ],   B' = [ bf16(b[0..7]) | bf16(b_lo[0..7]) ]
//     A'.B' = sum_k a_lo b + a b_lo            (a_lo = a - trunc_tf32(a): what the hi*hi MMA dropped)
// 16 bf16 are 32 bytes = 8 tf32, so the cross tile has the raw tile's byte geometry (128 rows x 128 B,
// SWIZZLE_128B, +32 B per k-step): same descriptors, half the tensor-pipe time of two tf32 MMAs and a
// third less MMA operand traffic.  bf16 keeps 8 bits of factors that are 2^-11 relative, so the cross
// terms carry ~2^-19 relative error (unbiased, round-to-nearest): fp32-class like the rest.
//
// BPRE: the B operand's cross tile [ bf16(b) | bf16(b_lo) ] comes PRE-SPLIT from global memory (maps.bx: a byte-congruent
// mirror of B in which every 8 floats are replaced by 8 + 8 bf16, written once per parameter update by presplit_kernel)
// and is loaded by TMA straight into the cross region; the splitter warps then only convert the A rows.  The dense layers'
// B operand is a weight matrix that every CTA pair of every tile of every layer call would otherwise re-split: this
// halves the splitters' shared-memory traffic (measured: LSU wavefronts 52 % + tensor-core reads 35 % of the pipe).
struct Maps3 { CUtensorMap a, b, bx; };

template <bool A_KMAJOR, bool B_NMAJOR, bool XBF16, bool BPRE, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc2p_kernel(const __grid_constant__ Maps3 maps, GemmShape p, Epi epi, long long* tl, float* sk_ws, unsigned* sk_flags, unsigned sk_epoch) {
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
    static_assert(!XBF16 || (A_KMAJOR && !B_NMAJOR), "bf16 cross terms need K-major operands");
    static_assert(!BPRE || XBF16, "a pre-split B operand is a bf16 cross tile");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stg_base = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(stg_base + EPI_STG_BYTES);
    uint64_t* full = bars;                      // own TMA landed                      (local)
    uint64_t* split = bars + STAGES;            // lo tiles of BOTH CTAs ready          (leader's copy is used)
    uint64_t* empty = bars + 2 * STAGES;        // MMAs done reading the stage          (multicast commit)
    uint64_t* acc_full = bars + 3 * STAGES;     // [2] accumulator buffer complete      (multicast commit)
    uint64_t* acc_empty = bars + 3 * STAGES + 2;// [2] buffer drained by BOTH CTAs      (leader's copy is used)
    uint32_t* tmem_slot = (uint32_t*)(bars + NBARS);
    volatile PairWork* work = (volatile PairWork*)(bars + NBARS + 1);

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
        if (BPRE) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.bx) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&split[s], 2 * SPLIT_WARPS); mbar_init(&empty[s], 1); }
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
                    const uint32_t s = g % STAGES, ph = (g / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_expect_tx(&full[s], HI_BYTES + (BPRE ? B_BYTES : 0));
                    const int k0 = kz0 + kt * BK;
                    if (A_KMAJOR) tma_load_2d(st, &maps.a, &full[s], k0, m0);
                    else          tma_load_3d(st, &maps.a, &full[s], 0, k0, m0 / 32);
                    if (!B_NMAJOR) tma_load_2d(st + A_BYTES, &maps.b, &full[s], k0, nb0);
                    else           tma_load_3d(st + A_BYTES, &maps.b, &full[s], 0, k0, nb0 / 32);
                    if (BPRE) tma_load_2d(st + HI_BYTES + A_BYTES, &maps.bx, &full[s], k0, nb0);   // B cross half-tile, ready made
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
                const uint32_t idesc = make_idesc_base(!A_KMAJOR, B_NMAJOR) | ((uint32_t)(neff >> 3) << 17);
                // kind::f16: D = f32, A = B = bf16, both K-major, same M and N
                const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(neff >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
                const uint32_t acc = tmem_base + b * BN;
                mbar_wait(&acc_empty[b], (u & 1) ^ 1);          // both CTAs have drained this buffer (tile i-2)
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                TC2P_MARK(i, 0);
                for (int kt = kt0; kt < kt1; ++kt, ++g) {
                    const uint32_t s = g % STAGES, ph = (g / STAGES) & 1;
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
                        if (XBF16) {
                            mma_tf32_ss_2sm(acc, dah, dbh, idesc, ((kt - kt0) | ks) != 0);
                            mma_bf16_ss_2sm(acc, dal, dbl, idesc16, 1);      // "lo" region = the bf16 cross tiles
                        } else {
                            mma_tf32_ss_2sm(acc, dal, dbh, idesc, ((kt - kt0) | ks) != 0);
                            mma_tf32_ss_2sm(acc, dah, dbl, idesc, 1);
                            mma_tf32_ss_2sm(acc, dah, dbh, idesc, 1);
                        }
                    }
                    mma_commit_2sm(&empty[s]);
                }
                mma_commit_2sm(&acc_full[b]);
                TC2P_MARK(i, 1);
            }
        }
    } else if (warp >= SPLIT_WARP0 && warp < EPI_WARP0) {
        // ---------------- splitters (both CTAs) ----------------
        const int tix = threadIdx.x - SPLIT_WARP0 * 32;
        uint32_t g = 0;
        for (int idx = 0; idx < work->n_items; ++idx) {
            int tile, kb0, kb1;
            item_at(idx, tile, kb0, kb1);
            int n_kb = kb1 - kb0;
            if (p.k_split > 0) { int m0p, n0, kz0, KT, neff, z; tile_origin(tile, m0p, n0, kz0, KT, neff, z); n_kb = KT; }
            for (int kt = 0; kt < n_kb; ++kt, ++g) {
                const uint32_t s = g % STAGES, ph = (g / STAGES) & 1;
                mbar_wait(&full[s], ph);
                const uint32_t hi = smem_u32(smem + s * STAGE_BYTES) + (uint32_t)tix * 16u;
                const uint32_t lo = hi + HI_BYTES;
                constexpr int PER = (BPRE ? A_BYTES : HI_BYTES) / 16 / (SPLIT_WARPS * 32);    // 8 float4 per thread (4: A rows only)
                constexpr uint32_t STEP = SPLIT_WARPS * 32 * 16;
                float4 v[PER];
#pragma unroll
                for (int k = 0; k < PER; ++k)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "r"(hi + k * STEP));
                if (XBF16) {
                    // chunk P = tix + 256 k of the stage: k < 4 -> A rows, k >= 4 -> B rows; row r = (P % 1024) / 8,
                    // physical 16-byte chunk pc = P % 8 holds logical chunk c = pc ^ (r % 8) (SWIZZLE_128B), i.e.
                    // k-group g = c / 2, half h = c % 2.  Its 4 lo and 4 full values go to bytes 8h.. of logical
                    // chunks 2g and 2g + 1 of the cross tile (A: lo first; B: full first) = physical pc^h, pc^h^1.
#pragma unroll
                    for (int k = 0; k < PER; ++k) {
                        const uint32_t q = (uint32_t)tix + (uint32_t)(k & 3) * 256u;      // chunk within the A or B tile
                        const uint32_t r = q >> 3, pc = q & 7u, h = (pc ^ r) & 1u;
                        const uint32_t l0 = pack_bf16x2(tf32_trunc_rest(v[k].x), tf32_trunc_rest(v[k].y));
                        const uint32_t l1 = pack_bf16x2(tf32_trunc_rest(v[k].z), tf32_trunc_rest(v[k].w));
                        const uint32_t f0 = pack_bf16x2(v[k].x, v[k].y), f1 = pack_bf16x2(v[k].z, v[k].w);
                        const bool is_b = k >= 4;
                        const uint32_t row = smem_u32(smem + s * STAGE_BYTES) + HI_BYTES + (is_b ? A_BYTES : 0) + r * 128u + (h << 3);
                        const uint32_t d0 = row + ((pc ^ h) << 4), d1 = row + ((pc ^ h ^ 1u) << 4);
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d0), "r"(is_b ? f0 : l0), "r"(is_b ? f1 : l1) : "memory");
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d1), "r"(is_b ? l0 : f0), "r"(is_b ? l1 : f1) : "memory");
                    }
                } else
                // hi = the raw fp32 tile (the tensor core ignores the 13 low mantissa bits); lo = rn_tf32(x - trunc(x))
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    float4 l;
                    l.x = __uint_as_float(f2tf32(v[k].x - __uint_as_float(__float_as_uint(v[k].x) & 0xFFFFE000u)));
                    l.y = __uint_as_float(f2tf32(v[k].y - __uint_as_float(__float_as_uint(v[k].y) & 0xFFFFE000u)));
                    l.z = __uint_as_float(f2tf32(v[k].z - __uint_as_float(__float_as_uint(v[k].z) & 0xFFFFE000u)));
                    l.w = __uint_as_float(f2tf32(v[k].w - __uint_as_float(__float_as_uint(v[k].w) & 0xFFFFE000u)));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo + k * STEP), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
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
                        if (row < M && cvalid) c = e.apply4(row, col, acc[k], ca, ra[k]);
                        if (Epi::kRowSum) {
                            c += __shfl_xor_sync(0xffffffffu, c, 4); c += __shfl_xor_sync(0xffffffffu, c, 2);
                            c += __shfl_xor_sync(0xffffffffu, c, 1);
                            rs[it0 + k] += c;               // 32-column sums; pairs of chunks make the 64-column groups
                        }
                    }
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
template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
inline bool eligible(const GemmShape& p, const Epi& epi) {
    return tc2::eligible<A_KMAJOR, B_NMAJOR>(p) && p.N % 4 == 0 && epi.vec_ok();
}

int sm_pairs();                             // number of TPC pairs to keep resident (74 on B200)
// per-stream stream-K scratch (SK slots of 256 KB + flags, allocated on first use) and the launch epoch the
// flags are compared with; false when stream-K is switched off (env MFM_STREAMK=0 / mfm_set_gemm_streamk)
bool streamk_workspace(cudaStream_t st, float** ws, unsigned** flags, unsigned* epoch);
int gemm_cross_bf16();                      // 1: K-major x K-major GEMMs take their cross terms from one bf16 MMA (env MFM_GEMM_CROSS=tf32|bf16)

// pre-split mirrors of weight buffers registered for the duration of an ABI call (rng.cu): mirror pointer for `p`, or null
const float* lookup_cross(const float* p);

template <bool A_KMAJOR, bool B_NMAJOR, class Epi>
inline cudaError_t launch(const GemmShape& p, const Epi& epi, cudaStream_t st) {
    Maps3 maps;
    bool ok = A_KMAJOR ? tc::make_map_kmajor(&maps.a, p.A, p.lda, p.M, p.K, BM) : tc::make_map_mnmajor(&maps.a, p.A, p.lda, p.M, p.K, BM / 32);
    ok = ok && (!B_NMAJOR ? tc::make_map_kmajor(&maps.b, p.B, p.ldb, p.N, p.K, BNH) : tc::make_map_mnmajor(&maps.b, p.B, p.ldb, p.N, p.K, BNH / 32));
    if (!ok) return cudaErrorInvalidValue;
    constexpr bool KK = A_KMAJOR && !B_NMAJOR;
    const bool xb = KK && gemm_cross_bf16();
    const float* bx = (xb && p.K % 8 == 0 && p.ldb % 8 == 0) ? lookup_cross(p.B) : nullptr;
    if (bx && !tc::make_map_kmajor(&maps.bx, bx, p.ldb, p.N, p.K, BNH)) return cudaErrorInvalidValue;
    if (!bx) maps.bx = maps.b;
    auto kern = bx ? gemm_tc2p_kernel<A_KMAJOR, B_NMAJOR, KK, KK, Epi>
                   : (xb ? gemm_tc2p_kernel<A_KMAJOR, B_NMAJOR, KK, false, Epi> : gemm_tc2p_kernel<A_KMAJOR, B_NMAJOR, false, false, Epi>);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc2p_kernel<A_KMAJOR, B_NMAJOR, false, false, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess && KK) e = cudaFuncSetAttribute(gemm_tc2p_kernel<A_KMAJOR, B_NMAJOR, KK, false, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess && KK) e = cudaFuncSetAttribute(gemm_tc2p_kernel<A_KMAJOR, B_NMAJOR, KK, KK, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const Tiles T = make_tiles(p.M, p.N, p.K, p.k_split);
    // stream-K (small remainder round cut along K) needs the per-stream workspace and the whole grid resident; with a
    // device-side row count the kernel decides by itself
    float* sk_ws = nullptr; unsigned* sk_flags = nullptr; unsigned sk_epoch = 0;
    const int rem = T.total % sm_pairs();
    const bool sk = p.k_split == 0 && (p.K + BK - 1) / BK >= SK_MAX_SPLIT * SK_MIN_KB && (p.n_rows_dev || (rem != 0 && rem * SK_MAX_SPLIT <= sm_pairs())) &&
                    streamk_workspace(st, &sk_ws, &sk_flags, &sk_epoch);
    const int pairs = (sk || T.total >= sm_pairs()) ? sm_pairs() : T.total;
    kern<<<dim3(2 * pairs), THREADS, SMEM_BYTES, st>>>(maps, p, epi, tc2::gemm_timeline(), sk_ws, sk_flags, sk_epoch);
    ++g_mfm_launches;
    return cudaGetLastError();
}

}  // namespace tc2p
}  // namespace mfm
