// probe: what does a 4-D TMA box {16, 1, G, R} of 16-bit elements leave in shared memory under SWIZZLE_128B / SWIZZLE_32B?
// (finding: SWIZZLE_128B pads every 32-byte inner row to its own 128-byte line; SWIZZLE_32B packs them densely)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma4d_probe scripts/probes/tma4d_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, uint16_t* out, int part, int grp, int row, int bytes_expected, int* done_bytes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* tile = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(tile + 65536);
    for (int i = threadIdx.x; i < 65536 / 2; i += blockDim.x) ((uint16_t*)tile)[i] = 0xFFFF;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes_expected) : "memory");
        uint32_t d = (uint32_t)__cvta_generic_to_shared(tile);
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(d), "l"(&map), "r"(b), "r"(0), "r"(part), "r"(grp), "r"(row) : "memory");
        uint32_t ok = 0; int spins = 0;
        while (!ok && spins < 2000000) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
            ++spins;
        }
        *done_bytes = ok ? 1 : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 65536 / 2; i += blockDim.x) out[i] = ((uint16_t*)tile)[i];
}
int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    const int rows = 256, cols = 256;          // "floats" per row; halves per row = 2 * cols
    std::vector<uint16_t> h((size_t)rows * cols * 2);
    // value encodes (row, group, part, e): row * 1024 + grp * 32 + part * 16 + e   (< 65536 for rows < 64)
    for (int r = 0; r < rows; ++r) for (int g = 0; g < cols / 16; ++g) for (int p = 0; p < 2; ++p) for (int e = 0; e < 16; ++e)
        h[(size_t)r * cols * 2 + g * 32 + p * 16 + e] = (uint16_t)((r % 64) * 1024 + g * 32 + p * 16 + e);
    uint16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    uint16_t* out; cudaMalloc(&out, 65536); int* done; cudaMalloc(&done, 4);
    for (int mode = 0; mode < 2; ++mode) {
        const int R = 64, G = mode == 0 ? 4 : 1;
        CUtensorMap m;
        cuuint64_t gdim[4] = {16, 2, (cuuint64_t)(cols / 16), (cuuint64_t)rows};
        cuuint64_t gstr[3] = {32, 64, (cuuint64_t)cols * 4};
        cuuint32_t box[4] = {16, 1, (cuuint32_t)G, (cuuint32_t)R};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("mode %d (%s, %d groups per box) encode rc=%d\n", mode, mode == 0 ? "SWIZZLE_128B" : "SWIZZLE_32B", G, (int)rc);
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2048);
        probe<<<1, 256, 65536 + 2048>>>(m, out, 1, 4, 0, 16 * G * R * 2, done);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  sync: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint16_t> o(32768); int dn; cudaMemcpy(o.data(), out, 65536, cudaMemcpyDeviceToHost); cudaMemcpy(&dn, done, 4, cudaMemcpyDeviceToHost);
        int last = -1; long written = 0;
        for (int i = 0; i < 32768; ++i) if (o[i] != 0xFFFF) { last = i; ++written; }
        printf("  barrier completed with %d expected bytes: %d; halves written %ld, last written half index %d (byte %d)\n", 16 * G * R * 2, dn, written, last, last * 2);
        // print the first 4 rows of 128 B (64 halves) decoded
        for (int r = 0; r < (mode == 0 ? 4 : 3); ++r) {
            printf("  smem row %d:", r);
            for (int c = 0; c < 64; c += 8) { uint16_t v = o[r * 64 + c]; if (v == 0xFFFF) printf(" [----]"); else printf(" [r%d g%d p%d e%d]", v / 1024, (v % 1024) / 32, (v % 32) / 16, v % 16); }
            printf("\n");
        }
    }
    return 0;
}
