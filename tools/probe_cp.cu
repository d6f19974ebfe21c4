// Probe (B200, sm_100a): tcgen05.cp.128x256b from a K-major SWIZZLE_128B shared-memory slab into TMEM - does the copy see the slab
// the way an MMA A operand does (k-step s = descriptor start + 32 bytes), and does lane m / column 8 s + j receive the bf16 pair
// (m, 16 s + 2 j), (m, 16 s + 2 j + 1)?  That is the layout tcgen05.mma reads a TMEM A operand in (8 columns per k-step).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dct_pruning_b200/csrc -o tools/probe_cp_bin tools/probe_cp.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "umma.cuh"
using namespace dctp::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__host__ __device__ inline uint32_t kmajor_off(uint32_t row, uint32_t k) {
    return (row >> 3) * 1024u + (row & 7) * 128u + ((((k >> 3) ^ row) & 7) << 4) + ((k & 7) << 1);
}

__global__ void __launch_bounds__(128) probe(uint32_t* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 128 * 64; i += 128) {
        const uint32_t m = i >> 6, k = i & 63;
        *reinterpret_cast<uint16_t*>(smem + kmajor_off(m, k)) = static_cast<uint16_t>((m << 6) | k);
    }
    if (warp == 0) tmem_alloc<64>(&slot);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint64_t desc = make_smem_desc(smem_u32(smem), 16, 1024, SWIZZLE_128B);
        for (int s = 0; s < 4; ++s) tmem_cp_128x256b(tmem + 8 * s, desc_advance(desc, 32 * s));
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    uint32_t v[32];
    tmem_ld32(tmem + ((warp * 32u) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[tid * 32 + j] = v[j];
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
    uint32_t* d;
    CK(cudaMalloc(&d, 128 * 32 * 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    probe<<<1, 128, 16384 + 1024>>>(d);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> h(128 * 32);
    CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int j = 0; j < 32; ++j) {
            const uint32_t want = (uint32_t)((m << 6) | (2 * j)) | ((uint32_t)((m << 6) | (2 * j + 1)) << 16);
            if (h[m * 32 + j] != want && bad++ < 12) {
                const uint32_t g = h[m * 32 + j];
                printf("lane %d col %d: got (m %u k %u | m %u k %u), want k %d, %d\n", m, j, (g & 0xFFFF) >> 6, g & 63, (g >> 16) >> 6, (g >> 16) & 63, 2 * j, 2 * j + 1);
            }
        }
    printf("tcgen05.cp.128x256b from a SWIZZLE_128B K-major slab: %d of %d words differ from the TMEM A-operand layout\n", bad, 128 * 32);
    return 0;
}
