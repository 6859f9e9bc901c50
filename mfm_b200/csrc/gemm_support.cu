// Host-side state of the dense-layer kernels (backend switches, pre-split weight mirrors, per-stream scratch) and the
// small reduction kernel the scaled-fp16 GEMM falls back to for operands whose maximum nobody tracked.
#include "internal.h"
#include "gemm_tf32x3.cuh"
#include <stdio.h>
#include <string.h>
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
    // (never both kinds of mirror: a workspace holds the bf16-cross mirrors only while the h16 kernel is switched off)
    for (int i = 0; i < g_n_cross; ++i)
        if (p >= g_cross[i].base && p < g_cross[i].base + g_cross[i].n && ((p - g_cross[i].base) % 8) == 0) return g_cross[i].mirror + (p - g_cross[i].base);
    return nullptr;
}
static int g_streamk = -1;
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
namespace tc2h {
// scaled-fp16 three-pass dense-layer kernel (gemm_tcgen05_h16.cuh): the default for K-major x K-major layers with 16-aligned K;
// MFM_GEMM_H16=0 / mfm_set_gemm_h16(0) falls back to the tf32 + bf16-cross kernel
static int g_h16 = -1, g_groups = -1;
int split_groups() {
    if (g_groups < 0) { const char* e = getenv("MFM_H16_GROUPS"); g_groups = (e && (e[0] == '1' || e[0] == '4')) ? e[0] - '0' : 2; }
    return g_groups;
}
int gemm_h16() {
    if (g_h16 < 0) { const char* e = getenv("MFM_GEMM_H16"); g_h16 = (e && e[0] == '0') ? 0 : ((e && e[0] == '2') ? 2 : 1); }
    return g_h16;
}
// ---- pre-split fp16 weight mirrors: [base, base + n) -> mirror in the split16 layout, split with h16_scale(*amax) ------------
struct MirrorRange { const float* base; size_t n; const float* mirror; const float* amax; };
static thread_local MirrorRange g_mirror[8];
static thread_local int g_n_mirror = 0;
void register_mirror_h16(const float* base, size_t n_floats, const float* mirror, const float* amax) {
    if (g_n_mirror < 8) g_mirror[g_n_mirror++] = MirrorRange{base, n_floats, mirror, amax};
}
void clear_mirrors_h16() { g_n_mirror = 0; }
const float* lookup_mirror_h16(const float* p, const float** amax) {
    for (int i = 0; i < g_n_mirror; ++i)
        if (p >= g_mirror[i].base && p < g_mirror[i].base + g_mirror[i].n && ((p - g_mirror[i].base) % 16) == 0) {
            *amax = g_mirror[i].amax;
            return g_mirror[i].mirror + (p - g_mirror[i].base);
        }
    return nullptr;
}
// per-(device, stream) scratch of 64 floats for the maxima of operands nobody tracked (allocated on first use, like the
// stream-K scratch; stable pointers, so launches that use it can be captured into CUDA graphs)
struct AmaxWs { cudaStream_t st; int dev; float* p; };
static AmaxWs g_amax_ws[16];
static int g_n_amax_ws = 0;
float* amax_scratch(cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (int i = 0; i < g_n_amax_ws; ++i) if (g_amax_ws[i].st == st && g_amax_ws[i].dev == dev) return g_amax_ws[i].p;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone || g_n_amax_ws == 16) { cudaGetLastError(); return nullptr; }
    float* p = nullptr;
    if (cudaMalloc(&p, 64 * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_amax_ws[g_n_amax_ws++] = AmaxWs{st, dev, p};
    return p;
}
float* amax_scratch_for(cudaStream_t st) { return amax_scratch(st); }     // make sure a stream has its scratch BEFORE a capture on it begins
// out[0] = max |x[r, c]| over the first rows x cols (rows clipped by the device-side count): grid-stride float4 reads,
// one atomicMax per warp.  HBM-bound.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long ld, int rows, int cols,
                                                     const int* __restrict__ n_rows_dev, float* __restrict__ out) {
    if (n_rows_dev) rows = min(rows, *n_rows_dev);
    const int c4 = cols >> 2;
    const long long total = (long long)rows * c4;
    float m = 0.0f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c4; const int c = (int)(i - r * c4);
        m = amax4(m, __ldg(reinterpret_cast<const float4*>(x + r * ld) + c));
    }
    amax_publish_warp(out, m);
}
cudaError_t launch_absmax(const float* x, long long ld, int rows, int cols, const int* n_rows_dev, float* out, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
    if (e != cudaSuccess) return e;
    const long long total = (long long)rows * (cols >> 2);
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    absmax_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(x, ld, rows, cols, n_rows_dev, out);
    ++g_mfm_launches;
    return cudaGetLastError();
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
extern "C" void mfm_set_gemm_h16(int v) { mfm::tc2h::g_h16 = v < 0 ? 0 : (v > 2 ? 1 : v); }
extern "C" int mfm_gemm_h16_enabled(void) { return mfm::tc2h::gemm_h16(); }
// narrowest hidden width for which the MLP runs on the scaled-fp16 path (maxima, operator norms, mirrors, pre-split copies): below
// it the layers are memory- / latency-bound and the extra copies cost more than the faster arithmetic gains (phi-four, H = 128:
// 1.40 M chain-steps/s without, 1.24 M with).  Tests force 0 to cover the hand-over logic at small widths.
namespace mfm { namespace tc2h { static int g_min_hidden = 512; int h16_min_hidden() { return g_min_hidden; } } }
extern "C" void mfm_debug_set_h16_min_hidden(int h) { mfm::tc2h::g_min_hidden = h < 0 ? 512 : h; }
extern "C" void mfm_debug_set_h16_groups(int g) { mfm::tc2h::g_groups = (g == 1 || g == 4) ? g : 2; }   // tuning aid, not in the ABI header
extern "C" void mfm_gemm_register_mirror(const float* base, long long n_floats, const float* mirror) {
    if (!(base && mirror && n_floats > 0)) { mfm::tc2p::clear_cross(); mfm::tc2h::clear_mirrors_h16(); return; }
    if (mfm::tc2h::gemm_h16()) mfm::tc2h::register_mirror_h16(base, (size_t)n_floats, mirror, mirror + n_floats);   // max |base| behind the parts (mfm_gemm_presplit)
    else mfm::tc2p::register_cross(base, (size_t)n_floats, mirror);
}
extern "C" const char* mfm_gemm_describe(void) {
    if (mfm::tc2h::gemm_h16())
        return "K-major layers: operands scaled by a per-tensor power of two and split into 2 fp16 parts, 3 kind::f16 tcgen05 MMAs per 16 k "
               "(ceiling 1/3 of the bf16 peak); weight gradients: the same three passes on MN-major tiles of the operands' pre-split copies; "
               "unaligned / tiny shapes: 3xTF32 (tcgen05 or mma.sync)";
    return "K-major layers: tf32 hi*hi + bf16 cross terms, 2 tcgen05 MMAs per 8 k (ceiling 1/4 of the bf16 peak); weight gradients: 3xTF32";
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
