// tcgen05.mma issue-rate probe: cycles per MMA (M=128, K=16, bf16 -> f32) for N in {32,64,128,256}, A from TMEM (.ts) or smem (.ss),
// accumulating into one D tile or alternating between two.  One CTA per SM, one issuing thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dct_pruning_b200/csrc -o /tmp/mma_rate tools/mma_rate.cu
#include <cstdio>
#include "umma.cuh"
using namespace dctp::umma;
template <int N, bool TS, int NACC, bool FRESH = false>
__global__ void __launch_bounds__(128) probe(long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x * 16; i < 160 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    fence_async_smem(); tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = slot;
    const uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024, SWIZZLE_128B);
    const uint64_t db = make_smem_desc(smem_u32(smem + 65536), 16, 1024, SWIZZLE_128B);
    if (threadIdx.x < 32) {
        if (elect_one()) {
            long long t0 = clock64();
            uint32_t phase = 0;
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t d = tmem + 256 + (NACC == 2 ? (i & 1) * 128 : 0);
                    if (FRESH) {     // operands walk through 2 x 16 KB (A) and 3 x 16 KB (B) slabs, 4 k-steps each: nothing is re-read back to back
                        const uint64_t fa = desc_with_lo(da, static_cast<uint32_t>(da) + ((i >> 2) & 1) * 1024 + (i & 3) * 2);
                        const uint64_t fb = desc_with_lo(db, static_cast<uint32_t>(db) + ((i >> 2) % 3) * 1024 + (i & 3) * 2);
                        if (TS) mma_bf16_ts(d, tmem + (i & 7) * 8, fb, idesc, 1);
                        else mma_bf16_ss(d, fa, fb, idesc, 1);
                    } else if (TS) mma_bf16_ts(d, tmem + (i & 3) * 8, db, idesc, 1);
                    else mma_bf16_ss(d, da, db, idesc, 1);
                }
                mma_commit(&bar);
                mbar_wait(&bar, phase); phase ^= 1;
            }
            long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
template <int N, bool TS, int NACC, bool FRESH = false>
void run(long long* d_out, const char* name) {
    cudaFuncSetAttribute(probe<N, TS, NACC, FRESH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int reps = 256;
    probe<N, TS, NACC, FRESH><<<148, 128, 160 * 1024>>>(d_out, reps);
    probe<N, TS, NACC, FRESH><<<148, 128, 160 * 1024>>>(d_out, reps);
    long long h = 0; cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %.1f cycles per MMA (incl. commit+wait per 16)  %s\n", name, double(h) / (reps * 16), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    run<64, true, 1>(d, "ts N=64 one accumulator");
    run<64, true, 2>(d, "ts N=64 two accumulators");
    run<64, false, 1>(d, "ss N=64 one accumulator");
    run<32, true, 1>(d, "ts N=32 one accumulator");
    run<128, true, 1>(d, "ts N=128 one accumulator");
    run<128, false, 1>(d, "ss N=128 one accumulator");
    run<256, true, 1>(d, "ts N=256 one accumulator");
    run<256, false, 1>(d, "ss N=256 one accumulator");
    run<128, false, 1, true>(d, "ss N=128 fresh operands");
    run<128, true, 1, true>(d, "ts N=128 fresh B operand");
    run<112, true, 1, true>(d, "ts N=112 fresh B operand");
    run<64, false, 1, true>(d, "ss N=64 fresh operands");
    run<64, true, 1, true>(d, "ts N=64 fresh B operand");
    run<256, false, 1, true>(d, "ss N=256 fresh operands");
    return 0;
}
