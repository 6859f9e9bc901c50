// Shared device helpers for the MFM hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define MFM_OK 0
#define MFM_ERR_ARG (-1)
#define MFM_ERR_CUDA (-2)
#define MFM_ERR_WORKSPACE (-3)
#define MFM_ERR_UNSUPPORTED (-4)

#define MFM_CUDA_CHECK(expr)                                  \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) { mfm_set_last_error(_e, __FILE__, __LINE__); return MFM_ERR_CUDA; } \
    } while (0)
// every kernel launch site is followed by exactly one MFM_LAUNCH_CHECK (or goes through launch_gemm),
// which also feeds the launch counter exposed as mfm_launch_count().
extern unsigned long long g_mfm_launches;
#define MFM_LAUNCH_CHECK() do { ++g_mfm_launches; MFM_CUDA_CHECK(cudaGetLastError()); } while (0)

void mfm_set_last_error(cudaError_t e, const char* file, int line);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// threefry2x32-20 with JAX's key schedule (jax/_src/prng.py threefry2x32_p; Random123 KATs).
// ---------------------------------------------------------------------------------------------
struct u32x2 { uint32_t a, b; };

__host__ __device__ __forceinline__ uint32_t mfm_rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ u32x2 threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
    x0 += k0; x1 += k1;
#define MFM_TF_R(r) { x0 += x1; x1 = mfm_rotl(x1, r); x1 ^= x0; }
    MFM_TF_R(13) MFM_TF_R(15) MFM_TF_R(26) MFM_TF_R(6)
    x0 += k1; x1 += k2 + 1u;
    MFM_TF_R(17) MFM_TF_R(29) MFM_TF_R(16) MFM_TF_R(24)
    x0 += k2; x1 += k0 + 2u;
    MFM_TF_R(13) MFM_TF_R(15) MFM_TF_R(26) MFM_TF_R(6)
    x0 += k0; x1 += k1 + 3u;
    MFM_TF_R(17) MFM_TF_R(29) MFM_TF_R(16) MFM_TF_R(24)
    x0 += k1; x1 += k2 + 4u;
    MFM_TF_R(13) MFM_TF_R(15) MFM_TF_R(26) MFM_TF_R(6)
    x0 += k2; x1 += k0 + 5u;
#undef MFM_TF_R
    u32x2 o; o.a = x0; o.b = x1; return o;
}

// Word `i` of jax's random_bits(key, 32, n) stream ("halves" layout, n padded to even):
// counts split in halves (lo, hi); word i < half is o0 of block (i, i+half), else o1 of
// block (i-half, i).  The padded count (odd n) is 0, not n.
__host__ __device__ __forceinline__ uint32_t threefry_stream_word(uint32_t k0, uint32_t k1, uint32_t i, uint32_t n) {
    const uint32_t half = (n + 1u) >> 1;
    const bool first = i < half;
    const uint32_t lo = first ? i : i - half;
    uint32_t hi = lo + half;
    if (hi >= n) hi = 0u;                      // odd-size pad element
    const u32x2 o = threefry2x32(k0, k1, lo, hi);
    return first ? o.a : o.b;
}

// key `j` of jax.random.split(key, num): words (2j, 2j+1) of the 2*num stream.
__host__ __device__ __forceinline__ u32x2 threefry_split_key(uint32_t k0, uint32_t k1, uint32_t j, uint32_t num) {
    u32x2 r;
    r.a = threefry_stream_word(k0, k1, 2u * j, 2u * num);
    r.b = threefry_stream_word(k0, k1, 2u * j + 1u, 2u * num);
    return r;
}

// jax.random.uniform float32 in [0,1): mantissa fill then subtract one.
__host__ __device__ __forceinline__ float bits_to_unit_float(uint32_t bits) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
#else
    union { uint32_t u; float f; } c; c.u = (bits >> 9) | 0x3F800000u; return c.f - 1.0f;
#endif
}

// XLA ErfInv (f32): Giles' single-precision polynomial in w = -log1p(-x*x), restated with XLA CPU's operation order
// (xla/client/lib/math.cc ErfInv32 + elemental_ir_emitter EmitLog1p; third-party, restated from the published sources - unpinned):
//   * every multiply and add rounds separately (XLA CPU does not contract to FMA without fast-math): __fmul_rn / __fadd_rn;
//   * log1p(t) = (-0.5 t + 1) t for |t| < 1e-4, else log(1 + t) with the sum rounded to float32 first;
//   * the logarithm itself is taken in double and rounded, i.e. correctly rounded logf - XLA's vectorised polynomial is within
//     an ulp of that; oracle/threefry.py does the same, so device and oracle agree bit for bit.
__device__ __forceinline__ float xla_log1p_f32(float t) {
    if (fabsf(t) < 1e-4f) return __fmul_rn(__fadd_rn(__fmul_rn(-0.5f, t), 1.0f), t);
    return (float)log((double)__fadd_rn(1.0f, t));
}
__device__ __forceinline__ float xla_erfinv_f32(float x) {
    float w = -xla_log1p_f32(-__fmul_rn(x, x));
    const bool lt = w < 5.0f;
    w = lt ? __fadd_rn(w, -2.5f) : __fadd_rn(__fsqrt_rn(w), -3.0f);
    float p = lt ? 2.81022636e-08f : -0.000200214257f;
    p = __fadd_rn(lt ? 3.43273939e-07f : 0.000100950558f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? -3.5233877e-06f : 0.00134934322f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? -4.39150654e-06f : -0.00367342844f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? 0.00021858087f : 0.00573950773f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? -0.00125372503f : -0.0076224613f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? -0.00417768164f : 0.00943887047f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? 0.246640727f : 1.00167406f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? 1.50140941f : 2.83297682f, __fmul_rn(p, w));
    const float r = __fmul_rn(p, x);
    return fabsf(x) == 1.0f ? x * INFINITY : r;
}

// ---- the same normal, cheaper: correctly rounded logf from a 16-entry table ------------------------------------------------
// The double-precision log above is ~40 % of the cost of a normal draw, and the two kernels that draw one or two normals per
// state element (pines_propose_kernel, fm_batch_kernel) are ALU-bound on it.  log_rn_f32 computes log(z) for a float32 z in
// double with glibc-logf's range reduction (z = 2^k z', z' in [0.699, 1.398), 16 sub-intervals with centre c_i, r = z' / c_i - 1,
// |r| < 0.03; the interval around 1 has c = 1 so that small logarithms keep their relative accuracy) and a degree-9 series for
// log1p(r): error < 2^-46 relative, 12 FP64 operations instead of ~35.  A Ziv test makes the result EXACTLY the rounding of
// the true logarithm: if y (1 - 2^-44) and y (1 + 2^-44) round to different floats (probability 2^-19) the generic path
// runs instead.  mfm_debug_normal_fast_check compares the two paths over the whole 32-bit input range (tests/test_gpu_rng.py).
// {1 / c_i, log c_i}: generated with 50-digit arithmetic (scripts in DESIGN.md section 5), hex floats are exact
static __constant__ double2 c_log16[16] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2}, {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2},
    {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3}, {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4}, {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5},
    {0x1.0000000000000p+0, 0x0.0p+0},              {0x1.e573ae5c66190p-1, 0x1.b42db8ba5c447p-5},  {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},  {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},
    {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
// copy the table to shared memory (divergent indices into constant memory serialise); block-wide, blockDim.x >= 16
__device__ __forceinline__ void log_tab_load(double2* s) {
    if (threadIdx.x < 16) s[threadIdx.x] = c_log16[threadIdx.x];
    __syncthreads();
}
__device__ __forceinline__ float log_rn_f32(float z, const double2* __restrict__ tab) {
    if (!(z > 1e-30f && z < 1e30f)) return (float)log((double)z);
    const uint32_t ix = __float_as_uint(z);
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double2 T = tab[i];
    const double r = fma((double)__uint_as_float(iz), T.x, -1.0);
    double p = 1.0 / 9.0;
    p = fma(p, r, -1.0 / 8.0); p = fma(p, r, 1.0 / 7.0); p = fma(p, r, -1.0 / 6.0); p = fma(p, r, 1.0 / 5.0);
    p = fma(p, r, -1.0 / 4.0); p = fma(p, r, 1.0 / 3.0); p = fma(p, r, -0.5);
    const double y = fma((double)k, 0x1.62e42fefa39efp-1, T.y) + fma(r * r, p, r);
    const float a = (float)(y * (1.0 - 0x1p-44)), b = (float)(y * (1.0 + 0x1p-44));
    return a == b ? a : (float)log((double)z);
}
__device__ __forceinline__ float xla_log1p_f32_t(float t, const double2* tab) {
    if (fabsf(t) < 1e-4f) return __fmul_rn(__fadd_rn(__fmul_rn(-0.5f, t), 1.0f), t);
    return log_rn_f32(__fadd_rn(1.0f, t), tab);
}
__device__ __forceinline__ float xla_erfinv_tail_f32(float x, float w) {      // the polynomial part of xla_erfinv_f32
    const bool lt = w < 5.0f;
    w = lt ? __fadd_rn(w, -2.5f) : __fadd_rn(__fsqrt_rn(w), -3.0f);
    float p = lt ? 2.81022636e-08f : -0.000200214257f;
    p = __fadd_rn(lt ? 3.43273939e-07f : 0.000100950558f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? -3.5233877e-06f : 0.00134934322f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? -4.39150654e-06f : -0.00367342844f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? 0.00021858087f : 0.00573950773f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? -0.00125372503f : -0.0076224613f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? -0.00417768164f : 0.00943887047f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? 0.246640727f : 1.00167406f, __fmul_rn(p, w));
    p = __fadd_rn(lt ? 1.50140941f : 2.83297682f, __fmul_rn(p, w));
    const float r = __fmul_rn(p, x);
    return fabsf(x) == 1.0f ? x * INFINITY : r;
}
__device__ __forceinline__ float bits_to_normal_t(uint32_t bits, const double2* tab) {
    const float lo = -0.99999994f;
    const float u = fmaxf(lo, __fadd_rn(__fmul_rn(bits_to_unit_float(bits), 2.0f), lo));
    const float w = -xla_log1p_f32_t(-__fmul_rn(u, u), tab);
    return __fmul_rn(1.41421354f, xla_erfinv_tail_f32(u, w));
}

// jax.random.normal float32 from 32 random bits:
// u = max(lo, f*(1-lo)+lo) with lo = nextafter(-1,0); (1-lo) rounds to 2.0f in f32.
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
    const float lo = -0.99999994f;
    float u = fmaxf(lo, __fadd_rn(__fmul_rn(bits_to_unit_float(bits), 2.0f), lo));
    return __fmul_rn(1.41421354f, xla_erfinv_f32(u));
}

// ---- jax_enable_x64 (the reference as shipped, multi_modal.py:14): float64 draws ---------------------------------------------
// random_bits(key, 64, (n,)) generates 2n words in the halves layout and pairs word i (high) with word n + i (low): exactly the
// two outputs of ONE threefry block (i, n + i).  uniform: mantissa fill of the top 52 bits, minus one.  normal:
// sqrt(2) erfinv(uniform(nextafter(-1, 0), 1)) in float64 (1 - nextafter(-1, 0) rounds to 2.0).  The library computes in
// float32, so the draw is rounded once at the end; CUDA's double erfinv and XLA's differ by a few float64 ulps at most,
// invisible after that rounding except on ties.
__host__ __device__ __forceinline__ double bits64_to_unit_double(uint32_t hi, uint32_t lo) {
    const unsigned long long b = ((((unsigned long long)hi << 32) | lo) >> 12) | 0x3FF0000000000000ull;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b) - 1.0;
#else
    union { unsigned long long u; double f; } c; c.u = b; return c.f - 1.0;
#endif
}
__device__ __forceinline__ float rng_uniform_at(uint32_t k0, uint32_t k1, uint32_t i, uint32_t n, int x64) {
    if (!x64) return bits_to_unit_float(threefry_stream_word(k0, k1, i, n));
    const u32x2 o = threefry2x32(k0, k1, i, n + i);
    return (float)bits64_to_unit_double(o.a, o.b);
}
__device__ __forceinline__ float bits64_to_normal(uint32_t hi, uint32_t lo) {
    const double lo_ = -0x1.fffffffffffffp-1;
    const double u = fmax(lo_, bits64_to_unit_double(hi, lo) * 2.0 + lo_);
    return (float)(0x1.6a09e667f3bcdp+0 * erfinv(u));
}
__device__ __forceinline__ float rng_normal_at_t(uint32_t k0, uint32_t k1, uint32_t i, uint32_t n, const double2* tab) {   // float32 draws
    return bits_to_normal_t(threefry_stream_word(k0, k1, i, n), tab);
}
__device__ __forceinline__ float rng_normal_at(uint32_t k0, uint32_t k1, uint32_t i, uint32_t n, int x64) {
    if (!x64) return bits_to_normal(threefry_stream_word(k0, k1, i, n));
    const u32x2 o = threefry2x32(k0, k1, i, n + i);
    return bits64_to_normal(o.a, o.b);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; `red` is >= 32 floats of shared memory.  Result valid in all threads.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.0f;
    r = warp_sum(r);
    return r;
}

// ---- max |value| of a tensor a later GEMM will split (h16): producers fold it into a device slot ----------------------
// Non-negative floats order like their bit patterns, so one integer atomicMax per warp does it; NaNs are skipped by fmaxf
// (they still poison the data itself, which is what apply_if_finite looks at).
__device__ __forceinline__ void amax_publish_warp(float* slot, float vmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((threadIdx.x & 31) == 0 && vmax > 0.0f) atomicMax(reinterpret_cast<int*>(slot), __float_as_int(vmax));
}
// Block-wide variant for element-wise kernels with ~10^5 blocks (every thread of the block must call it): one candidate per
// block, and the atomic is skipped when the slot already holds a value at least as large - after the first few blocks nearly
// every block.  (Measured: one same-address atomic per WARP cost 0.5 ms per launch at 65 536 x 1 600 elements.)  The plain
// read may be stale (it can come from this SM's L1): stale means smaller, which only costs an unnecessary atomic.
__device__ __forceinline__ void amax_publish_block(float* slot, float vmax, float* red32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) red32[wid] = vmax;
    __syncthreads();
    if (wid == 0) {
        float m = lane < nw ? red32[lane] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m > 0.0f && !(m <= *slot)) atomicMax(reinterpret_cast<int*>(slot), __float_as_int(m));
    }
}
__device__ __forceinline__ float amax4(float m, const float4& v) { return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w))); }
