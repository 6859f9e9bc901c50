// jax.random-compatible generators (threefry2x32, legacy uint32[2] keys, x64 off).
// Replaces the XLA threefry2x32 / erf_inv HLO behind every jax.random.* call on the hot path
// (bblackjax/util.py:81, proposal.py:179, exe_flow_matching.py:142-166,212,265-275,303).
#include "common.cuh"
#include "../../include/mfm_b200.h"
#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";
void mfm_set_last_error(cudaError_t e, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
}
void mfm_set_last_error_msg(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }
extern "C" const char* mfm_last_error(void) { return g_err; }
extern "C" int mfm_version(void) { return 100; }
unsigned long long g_mfm_launches = 0;
#include <stdlib.h>
namespace mfm {
static int g_backend = -1;
int gemm_backend() {
    if (g_backend < 0) {
        const char* e = getenv("MFM_GEMM");
        g_backend = (e && strcmp(e, "mma") == 0) ? 1 : ((e && strcmp(e, "tc1") == 0) ? 2 : ((e && strcmp(e, "tc2") == 0) ? 3 : 0));
    }
    return g_backend;
}
namespace tc2p {
static int g_cross_bf16 = -1;
int gemm_cross_bf16() {
    if (g_cross_bf16 < 0) { const char* e = getenv("MFM_GEMM_CROSS"); g_cross_bf16 = (e && strcmp(e, "tf32") == 0) ? 0 : 1; }
    return g_cross_bf16;
}
int sm_pairs();
// ---- pre-split weight mirrors (gemm_tcgen05_persist.cuh, BPRE) -------------------------------------------------------
// [base, base + n_floats) -> mirror (same byte layout, every 8 floats replaced by 8 + 8 bf16).  Registered by the ABI call
// that built the mirrors in ITS workspace and cleared when it returns (CrossScope), so no stale range survives a call.
struct CrossRange { const float* base; size_t n; const float* mirror; };
static thread_local CrossRange g_cross[8];
static thread_local int g_n_cross = 0;
void register_cross(const float* base, size_t n_floats, const float* mirror) {
    if (g_n_cross < 8) g_cross[g_n_cross++] = CrossRange{base, n_floats, mirror};
}
void clear_cross() { g_n_cross = 0; }
const float* lookup_cross(const float* p) {
    for (int i = 0; i < g_n_cross; ++i)
        if (p >= g_cross[i].base && p < g_cross[i].base + g_cross[i].n && ((p - g_cross[i].base) % 8) == 0) return g_cross[i].mirror + (p - g_cross[i].base);
    return nullptr;
}
static int g_split16 = -1;
static int g_streamk = -1;
constexpr int SK_SLOT_FLOATS = 256 * 256, SK_SLOT_FLAGS = 64;   // = gemm_tcgen05_persist.cuh (static_assert there)
struct SkWs { cudaStream_t st; int dev; float* ws; unsigned* flags; unsigned epoch; };
static SkWs g_skws[8];
static int g_n_skws = 0;
// One scratch area per (device, stream): GEMMs on one stream are ordered, so a slot is never rewritten
// while an earlier launch still reads it.  Allocated on the first stream-K launch of the stream (the only
// allocation the library makes; 19.4 MB + 19 KB).  Streams beyond the 8th run without stream-K.
bool streamk_workspace(cudaStream_t st, float** ws, unsigned** flags, unsigned* epoch) {
    if (g_streamk < 0) { const char* e = getenv("MFM_STREAMK"); g_streamk = (e && e[0] == '0') ? 0 : 1; }
    if (!g_streamk) return false;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    // never inside a CUDA-graph capture: the launch epoch is a kernel argument, a replay would meet its own stale flags
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return false; }
    SkWs* w = nullptr;
    for (int i = 0; i < g_n_skws; ++i) if (g_skws[i].st == st && g_skws[i].dev == dev) { w = &g_skws[i]; break; }
    if (!w) {
        if (g_n_skws == 8) return false;
        SkWs n{st, dev, nullptr, nullptr, 0};
        const size_t fbytes = (size_t)sm_pairs() * SK_SLOT_FLAGS * sizeof(unsigned);
        if (cudaMalloc(&n.ws, (size_t)sm_pairs() * SK_SLOT_FLOATS * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return false; }
        if (cudaMalloc(&n.flags, fbytes) != cudaSuccess || cudaMemset(n.flags, 0, fbytes) != cudaSuccess) { cudaGetLastError(); cudaFree(n.ws); return false; }
        g_skws[g_n_skws] = n;
        w = &g_skws[g_n_skws++];
    }
    *ws = w->ws; *flags = w->flags; *epoch = ++w->epoch;
    return true;
}
}   // namespace tc2p
namespace tc2s {
// EXPERIMENTAL split16 dense-layer kernel (gemm_tcgen05_split16.cuh): off unless MFM_GEMM_SPLIT=bf16x3 / mfm_set_gemm_split16(1)
int gemm_split16() {
    if (tc2p::g_split16 < 0) { const char* e = getenv("MFM_GEMM_SPLIT"); tc2p::g_split16 = (e && strcmp(e, "bf16x3") == 0) ? 1 : 0; }
    return tc2p::g_split16;
}
}
namespace tc2p {
int sm_pairs() {
    static int pairs = 0;
    if (pairs == 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 2) sms = 148;
        pairs = sms / 2;
    }
    return pairs;
}
}
namespace tc2 {
static int g_raw_hi = -1;
int gemm_raw_hi() {
    if (g_raw_hi < 0) { const char* e = getenv("MFM_TC_RAWHI"); g_raw_hi = (e && e[0] == '0') ? 0 : 1; }
    return g_raw_hi;
}
}
}
extern "C" void mfm_set_gemm_raw_hi(int v) { mfm::tc2::g_raw_hi = v ? 1 : 0; }
extern "C" void mfm_set_gemm_cross_bf16(int v) { mfm::tc2p::g_cross_bf16 = v ? 1 : 0; }
extern "C" void mfm_set_gemm_streamk(int v) { mfm::tc2p::g_streamk = v ? 1 : 0; }
extern "C" void mfm_set_gemm_split16(int v) { mfm::tc2p::g_split16 = v ? 1 : 0; }
extern "C" void mfm_gemm_register_mirror(const float* base, long long n_floats, const float* mirror) {
    if (base && mirror && n_floats > 0) mfm::tc2p::register_cross(base, (size_t)n_floats, mirror); else mfm::tc2p::clear_cross();
}
// tuning aid (not part of the ABI header): SM-clock timeline of one CTA pair of the last tc2 GEMM
namespace mfm { namespace tc2 {
static long long* g_timeline_buf = nullptr;
static int g_timeline = 0;
long long* gemm_timeline() { return g_timeline ? g_timeline_buf : nullptr; }
} }
extern "C" int mfm_debug_gemm_timeline(int enable, long long* out16) {
    using namespace mfm::tc2;
    if (!g_timeline_buf && cudaMalloc(&g_timeline_buf, 64 * sizeof(long long)) != cudaSuccess) return -1;
    g_timeline = enable ? 1 : 0;
    if (out16) return cudaMemcpy(out16, g_timeline_buf, 64 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
    return 0;
}
extern "C" void mfm_set_gemm_backend(int b) { mfm::g_backend = b; }
extern "C" unsigned long long mfm_launch_count(void) { return g_mfm_launches; }

namespace {

__global__ void split_kernel(const uint32_t* __restrict__ keys, int n, int num, uint32_t* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)n * num) return;
    const int c = (int)(i / num), j = (int)(i % num);
    const u32x2 k = threefry_split_key(keys[2 * c], keys[2 * c + 1], (uint32_t)j, (uint32_t)num);
    out[2 * i] = k.a; out[2 * i + 1] = k.b;
}

// MODE 0 bits, 1 uniform, 2 normal.  One thread per threefry block -> two stream words.
template <int MODE>
__global__ void stream_kernel(const uint32_t* __restrict__ keys, int nkeys, long long n, float minval, float maxval,
                              void* __restrict__ out) {
    const long long half = (n + 1) >> 1;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= half * nkeys) return;
    const int c = (int)(i / half);
    const long long lo = i % half;
    long long hi = lo + half;
    const bool has_hi = hi < n;
    const u32x2 o = threefry2x32(keys[2 * c], keys[2 * c + 1], (uint32_t)lo, has_hi ? (uint32_t)hi : 0u);
    const long long base = (long long)c * n;
    if (MODE == 0) {
        uint32_t* p = (uint32_t*)out;
        p[base + lo] = o.a; if (has_hi) p[base + hi] = o.b;
    } else if (MODE == 1) {
        float* p = (float*)out;
        const float sc = maxval - minval;
        p[base + lo] = fmaxf(minval, __fadd_rn(__fmul_rn(bits_to_unit_float(o.a), sc), minval));  // no FMA: mul then add as XLA
        if (has_hi) p[base + hi] = fmaxf(minval, __fadd_rn(__fmul_rn(bits_to_unit_float(o.b), sc), minval));
    } else {
        float* p = (float*)out;
        p[base + lo] = bits_to_normal(o.a);
        if (has_hi) p[base + hi] = bits_to_normal(o.b);
    }
}

template <int MODE>
int launch_stream(const uint32_t* keys, int nkeys, long long n, float lo, float hi, void* out, cudaStream_t st) {
    if (n <= 0 || nkeys <= 0) return MFM_OK;
    if (n > 0xFFFFFFFFll) return MFM_ERR_UNSUPPORTED;
    const long long work = ((n + 1) >> 1) * nkeys;
    stream_kernel<MODE><<<ceil_div(work, 256), 256, 0, st>>>(keys, nkeys, n, lo, hi, out);
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

}  // namespace

extern "C" {

int mfm_threefry_split(const uint32_t* key, int num, uint32_t* out, mfm_stream_t stream) {
    return mfm_threefry_split_batched(key, 1, num, out, stream);
}

int mfm_threefry_split_batched(const uint32_t* keys, int n, int num, uint32_t* out, mfm_stream_t stream) {
    if (n <= 0 || num <= 0) return MFM_OK;
    split_kernel<<<ceil_div((long long)n * num, 256), 256, 0, stream>>>(keys, n, num, out);
    MFM_LAUNCH_CHECK();
    return MFM_OK;
}

int mfm_threefry_bits(const uint32_t* key, long long n, uint32_t* out, mfm_stream_t stream) {
    return launch_stream<0>(key, 1, n, 0.f, 1.f, out, stream);
}
int mfm_threefry_uniform(const uint32_t* key, long long n, float minval, float maxval, float* out, mfm_stream_t stream) {
    return launch_stream<1>(key, 1, n, minval, maxval, out, stream);
}
int mfm_threefry_normal(const uint32_t* key, long long n, float* out, mfm_stream_t stream) {
    return launch_stream<2>(key, 1, n, 0.f, 1.f, out, stream);
}
int mfm_threefry_uniform_batched(const uint32_t* keys, int n, int d, float minval, float maxval, float* out, mfm_stream_t stream) {
    return launch_stream<1>(keys, n, d, minval, maxval, out, stream);
}
int mfm_threefry_normal_batched(const uint32_t* keys, int n, int d, float* out, mfm_stream_t stream) {
    return launch_stream<2>(keys, n, d, 0.f, 1.f, out, stream);
}

void mfm_host_threefry_split(const uint32_t key[2], int num, uint32_t* out) {
    for (int j = 0; j < num; ++j) {
        const u32x2 k = threefry_split_key(key[0], key[1], (uint32_t)j, (uint32_t)num);
        out[2 * j] = k.a; out[2 * j + 1] = k.b;
    }
}

}  // extern "C"
